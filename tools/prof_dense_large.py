"""One batched dense-large solve for profiling.  Usage: python tools/prof_dense_large.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
nz, nl, nv = 512, 128, 1024
d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=5, nthreads=16)
s = fb.FBstabDense(nz, nl, nv, max_batch=B)
dev = torch.device("cuda:0")
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
for it in range(2):
    z = torch.zeros(B * nz, dtype=torch.float64, device=dev)
    l = torch.zeros(B * nl, dtype=torch.float64, device=dev)
    v = torch.zeros(B * nv, dtype=torch.float64, device=dev)
    out, y = s.solve_batch(dd, z, l, v)
    torch.cuda.synchronize()
print(s.path)
