"""Where the single-instance latency of the generic dense kernel goes (BASELINE config 1,
50/10/100): kernel alone (device-resident inputs, CUDA events) against the whole C-ABI call
with host buffers, per residency mode and block size.  Usage: python tools/probe_single_latency.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fbstab_b200 as fb

dev = torch.device("cuda:0")
nz, nl, nv = 50, 10, 100
d = fb.problems.random_dense_qp(nz, nl, nv, count=1, config=1)
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
for mode in ("0", "3"):
    for block in ("64", "128", "256"):
        os.environ["FBSTAB_DENSE_RESIDENT"] = mode
        os.environ["FBSTAB_BLOCK"] = block
        s = fb.FBstabDense(nz, nl, nv, max_batch=1)
        host, kern = [], []
        for _ in range(30):
            z, l, v = np.zeros(nz), np.zeros(nl), np.zeros(nv)
            t0 = time.perf_counter()
            o1, y1 = s.solve_batch(d, z, l, v)
            host.append(time.perf_counter() - t0)
        for _ in range(30):
            zt, lt, vt = (torch.zeros(n, dtype=torch.float64, device=dev) for n in (nz, nl, nv))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            out, y = s.solve_batch(dd, zt, lt, vt)
            e1.record()
            torch.cuda.synchronize()
            kern.append(e0.elapsed_time(e1))
        print(f"resident={mode} block={block}: host call {1e3 * np.median(host):.3f} ms, kernel alone "
              f"{np.median(kern):.3f} ms, newton {int(o1['newton_iters'][0])}, solve_time field "
              f"{float(o1['solve_time'][0]) * 1e3:.3f} ms", flush=True)
