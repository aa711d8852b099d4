"""Exit flags and Newton counts of all 16,384 instances of BASELINE config 3a on the GPU --
structured (FBstabMpc) and as general sparse QPs (FBstabSparse) -- written to
gpurun_out/flags_3a.npz for the comparison with the CPU oracles in profiles/r2_flag_floor.txt.
Usage: python tools/dump_flags_3a.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb

n = 16384
dims, d = fb.problems.ocp_batch("servo_motor", 50, count=n, config=3, rho=0.02)
m = fb.FBstabMpc(*dims, max_batch=n)
z, l, v = np.zeros(n * m.nz), np.zeros(n * m.nl), np.zeros(n * m.nv)
om, _ = m.solve_batch(d, z, l, v)
(nz, nl, nv), pat, vals = fb.problems.ocp_as_sparse_qp(dims, d, n)
s = fb.FBstabSparse(nz, nl, nv, pat, max_batch=n)
z, l, v = np.zeros(n * nz), np.zeros(n * nl), np.zeros(n * nv)
os_, _ = s.solve_batch(vals, z, l, v)
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/flags_3a.npz", mpc_eflag=om["eflag"], mpc_newton=om["newton_iters"],
         sparse_eflag=os_["eflag"], sparse_newton=os_["newton_iters"])
print(m.path[:40], np.bincount(om["eflag"], minlength=6), "|", s.path[:40],
      np.bincount(os_["eflag"], minlength=6))
