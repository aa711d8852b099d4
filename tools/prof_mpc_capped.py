"""One batched MPC solve with a Newton cap, for profiling long-running shapes.
Usage: python tools/prof_mpc_capped.py kind N batch newton_cap [rho]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb  # noqa: E402

kind, N, B, cap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rho = float(sys.argv[5]) if len(sys.argv) > 5 else 0.05
dims, d = fb.problems.ocp_batch(kind, N, count=B, config=4, rho=rho)
s = fb.FBstabMpc(*dims, max_batch=B)
s.update_options(fb.FBstabMpc.default_options(max_newton_iters=cap))
dev = torch.device("cuda:0")
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
for it in range(2):
    z = torch.zeros(B * s.nz, dtype=torch.float64, device=dev)
    l = torch.zeros(B * s.nl, dtype=torch.float64, device=dev)
    v = torch.zeros(B * s.nv, dtype=torch.float64, device=dev)
    out, y = s.solve_batch(dd, z, l, v)
    torch.cuda.synchronize()
print(s.path)
