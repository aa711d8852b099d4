"""Small MPC solves for compute-sanitizer: the CTA kernel (TMA stage-data ring, factor
ring) on a shape with an odd nu x nu block (read in place) and, with
FBSTAB_MPC_LANE_MIN=256, the lane kernel with shared and per-instance data."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb

lane = os.environ.get("SANITIZE_LANE") == "1"
cases = ([("servo_motor", 5, 256)] if lane else
         [("servo_motor", 6, 3), ("copolymerization", 4, 2), ("spacecraft", 5, 2)])
for kind, N, B in cases:
    dims, d = fb.problems.ocp_batch(kind, N, count=B, config=3, rho=0.01)
    m = fb.FBstabMpc(*dims, max_batch=B)
    m.update_options(fb.FBstabMpc.default_options(max_newton_iters=2 if lane else 3))
    z, l, v = np.zeros(B * m.nz), np.zeros(B * m.nl), np.zeros(B * m.nv)
    out, y = m.solve_batch(d, z, l, v)
    print(kind, m.path[:40], out["eflag"][:4], out["newton_iters"][:4])
    one = {k: (a if k == "x0" else a[:m.field_sizes[k]].copy()) for k, a in d.items()}
    z, l, v = np.zeros(B * m.nz), np.zeros(B * m.nl), np.zeros(B * m.nv)
    out, y = m.solve_batch_shared(one, z, l, v)
    print(kind, "shared entry", out["eflag"][:4], out["newton_iters"][:4])
