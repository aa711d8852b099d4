"""One single-instance solve of BASELINE config 1 (dense 50/10/100) on the generic CTA kernel,
for `ncu -k regex:dense_generic_kernel`.  Usage: python tools/prof_generic_single.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb

nz, nl, nv = 50, 10, 100
d = fb.problems.random_dense_qp(nz, nl, nv, count=1, config=1)
s = fb.FBstabDense(nz, nl, nv, max_batch=1)
for _ in range(2):
    z, l, v = np.zeros(nz), np.zeros(nl), np.zeros(nv)
    out, y = s.solve_batch(d, z, l, v)
print(s.path, out["newton_iters"], out["eflag"])
