"""A small sparse batch for compute-sanitizer (sparse_lane_kernel: 70 instances, the last
warp partly idle; mixed feasible / infeasible instances)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fbstab_b200 as fb
from test_sparse import random_sparse_qp

nz, nl, nv, B = 14, 3, 20, 70
pat, vals, _ = random_sparse_qp(np.random.default_rng(2), nz, nl, nv, count=B)
vals["b"][5 * nv:6 * nv] -= 50.0  # one instance with an empty feasible set
s = fb.FBstabSparse(nz, nl, nv, pat, max_batch=B)
z, l, v = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)
out, y = s.solve_batch(vals, z, l, v)
print(s.path[:50], np.bincount(out["eflag"], minlength=6).tolist(), out["newton_iters"][:6])
