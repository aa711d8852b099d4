"""A/B of the generic dense kernel's shared-memory residency (FBSTAB_DENSE_RESIDENT = 0:
global workspace, 2: LDL' workspace in shared memory, 3: data too): single-instance
latency through the C-ABI with host buffers (the reference's own case, BASELINE config 1),
batch throughput with device-resident inputs, and a hash of every output (the modes must
return identical bytes).  Usage: python tools/ab_dense_generic.py"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fbstab_b200 as fb

dev = torch.device("cuda:0")
for nz, nl, nv, B in [(50, 10, 100, 8192), (40, 0, 80, 8192), (96, 16, 180, 4096),
                      (120, 20, 240, 2048)]:
    d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=1)
    one = {k: np.ascontiguousarray(a[:a.size // B]) for k, a in d.items()}
    dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
    for mode in ("0", "2", "3"):
        os.environ["FBSTAB_DENSE_RESIDENT"] = mode
        s1 = fb.FBstabDense(nz, nl, nv, max_batch=1)
        lat = []
        for _ in range(40):
            z, l, v = np.zeros(nz), np.zeros(nl), np.zeros(nv)
            t0 = time.perf_counter()
            o1, y1 = s1.solve_batch(one, z, l, v)
            lat.append(time.perf_counter() - t0)
        s = fb.FBstabDense(nz, nl, nv, max_batch=B)
        ts = []
        for _ in range(3):
            zt, lt, vt = (torch.zeros(B * n, dtype=torch.float64, device=dev) for n in (nz, nl, nv))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            out, y = s.solve_batch(dd, zt, lt, vt)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
        h = hashlib.sha256()
        for t in (zt, lt, vt, y):
            h.update(t.cpu().numpy().tobytes())
        for f in ("eflag", "newton_iters", "prox_iters", "status", "ls_backtracks"):
            h.update(np.ascontiguousarray(o[f]).tobytes())
        print(f"{nz}/{nl}/{nv} resident={mode} path={s.path[:60]!r} single-instance latency "
              f"{1e3 * np.median(lat):.3f} ms (newton {int(o1['newton_iters'][0])}) | batch {B}: "
              f"{min(ts):.2f} ms = {B / min(ts) * 1e3:.0f} solves/s | flags "
              f"{np.bincount(o['eflag'], minlength=6).tolist()} sha {h.hexdigest()[:16]}", flush=True)
