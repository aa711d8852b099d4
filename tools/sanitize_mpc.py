import sys, numpy as np
sys.path.insert(0, '/root/repo')
import fbstab_b200 as fb
for kind, N, B in (("servo_motor", 6, 3), ("copolymerization", 4, 2)):
    dims, d = fb.problems.ocp_batch(kind, N, count=B, config=3, rho=0.01)
    m = fb.FBstabMpc(*dims, max_batch=B)
    m.update_options(fb.FBstabMpc.default_options(max_newton_iters=3))
    z, l, v = np.zeros(B * m.nz), np.zeros(B * m.nl), np.zeros(B * m.nv)
    out, y = m.solve_batch(d, z, l, v)
    print(kind, m.path, out["eflag"], out["newton_iters"])
