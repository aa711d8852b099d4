"""A/B timing of dense-small kernel variants on BASELINE config 2.
Usage: FBSTAB_B200_LIB=build/variants/X.so python tools/ab_dense_small.py [batch] [reps]
Prints ms per batched solve (CUDA events, device-resident inputs), iteration
statistics and a SHA-256 of every output byte (bit-identity across variants)."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fbstab_b200 as fb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nz, nl, nv = 32, 8, 64
d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=2)
s = fb.FBstabDense(nz, nl, nv, max_batch=B)
dev = torch.device("cuda:0")
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
ts = []
for it in range(reps):
    zt = torch.zeros(B * nz, dtype=torch.float64, device=dev)
    lt = torch.zeros(B * nl, dtype=torch.float64, device=dev)
    vt = torch.zeros(B * nv, dtype=torch.float64, device=dev)
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
for f in ("eflag", "newton_iters", "prox_iters", "status", "residual", "ls_backtracks"):
    h.update(np.ascontiguousarray(o[f]).tobytes())
print(os.environ.get("FBSTAB_B200_LIB", "default"), s.path[:24], "ms", " ".join(f"{t:.2f}" for t in ts),
      "| best %.0f solves/s" % (B / min(ts) * 1e3), "| flags", np.bincount(o["eflag"], minlength=6).tolist(),
      "newton %.3f" % o["newton_iters"].mean(), "sha", h.hexdigest()[:16])
