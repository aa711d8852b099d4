import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb
for nz, nl, nv in ((136, 24, 200), (129, 0, 70)):
    B = 2
    d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=12)
    s = fb.FBstabDense(nz, nl, nv, max_batch=B)
    s.update_options(fb.FBstabDense.default_options(max_newton_iters=2))
    z, l, v = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)
    out, y = s.solve_batch(d, z, l, v)
    print(s.path, out["eflag"], out["newton_iters"], out["status"])
