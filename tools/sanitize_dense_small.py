import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb
for nz, nl, nv in ((32, 8, 64), (5, 0, 7), (9, 3, 4), (31, 7, 63)):
    B = 20
    d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=12)
    s = fb.FBstabDense(nz, nl, nv, max_batch=B)
    s.update_options(fb.FBstabDense.default_options(max_newton_iters=3))
    z, l, v = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)
    out, y = s.solve_batch(d, z, l, v)
    print(s.path[:30], out["eflag"][:4], out["newton_iters"][:4], out["status"][:4])
