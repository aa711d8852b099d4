"""A/B timing + output hash of one BASELINE config (device-resident, CUDA events).
Usage: FBSTAB_B200_LIB=build/variants/X.so python tools/ab_config.py <cfg> [reps]"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import fbstab_b200 as fb

wl = bench.Workload(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = wl.batch
dev = torch.device("cuda:0")
d = wl.generate(fb.problems, B, 0, 8)
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
s = wl.solver(fb, B, 0)
ts = []
for _ in range(reps):
    z, l, v = (torch.zeros(B * n, dtype=torch.float64, device=dev) for n in (wl.nz, wl.nl, wl.nv))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out, y = s.solve_batch(dd, z, l, v)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
h = hashlib.sha256()
for t in (z, l, v, y):
    h.update(t.cpu().numpy().tobytes())
for f in ("eflag", "newton_iters", "prox_iters", "status", "residual", "ls_backtracks"):
    h.update(np.ascontiguousarray(o[f]).tobytes())
print(os.path.basename(os.environ.get("FBSTAB_B200_LIB", "tree")), "cfg", wl.name, s.path[:28], "ms",
      " ".join(f"{t:.2f}" for t in ts), "| best %.0f solves/s" % (B / min(ts) * 1e3), "| flags",
      np.bincount(o["eflag"], minlength=6).tolist(), "newton %.3f" % o["newton_iters"].mean(),
      "sha", h.hexdigest()[:16])
