"""One batched MPC solve for profiling.  Usage: python tools/prof_mpc.py kind N batch [rho]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb  # noqa: E402

kind, N, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rho = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
dims, d = fb.problems.ocp_batch(kind, N, count=B, config=3, rho=rho)
s = fb.FBstabMpc(*dims, max_batch=B)
dev = torch.device("cuda:0")
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
for it in range(2):
    z = torch.zeros(B * s.nz, dtype=torch.float64, device=dev)
    l = torch.zeros(B * s.nl, dtype=torch.float64, device=dev)
    v = torch.zeros(B * s.nv, dtype=torch.float64, device=dev)
    out, y = s.solve_batch(dd, z, l, v)
    torch.cuda.synchronize()
print(s.path)
