"""FBstabSparse on an OCP of the reference's generator restated as a GENERAL sparse QP
(same instances as the MPC configs of bench.py: the plants differ in x0 only), next to
FBstabMpc on the structured form.  Usage: python tests/sparse_bench.py [kind] [N] [batch]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import fbstab_b200 as fb
from test_sparse import csc, ocp_as_qp

kind = sys.argv[1] if len(sys.argv) > 1 else "servo_motor"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
cfg = {"servo_motor": ("3a", 0.02), "double_integrator": ("3b", 0.1)}.get(kind, ("4", 0.05))
dims, d = fb.problems.ocp_batch(kind, N, count=B, config=3, rho=cfg[1])
nx = dims[1]
H, f, G, h, A, b = ocp_as_qp(dims, d, 0)
Hp, Hi, Hx = csc(H, upper=True)
Gp, Gi, Gx = csc(G)
Ap, Ai, Ax = csc(A)
nz, nl, nv = f.size, h.size, b.size
hs = np.tile(h, B).reshape(B, nl)
hs[:, :nx] = -d["x0"].reshape(B, nx)          # the instances differ in x0 only
vals = {"Hx": np.tile(Hx, B), "f": np.tile(f, B), "Gx": np.tile(Gx, B),
        "h": np.ascontiguousarray(hs.reshape(-1)), "Ax": np.tile(Ax, B), "b": np.tile(b, B)}
dev = torch.device("cuda:0")
s = fb.FBstabSparse(nz, nl, nv, (Hp, Hi, Gp, Gi, Ap, Ai), max_batch=B)
n, nnzK, nnzL, perm = s.analysis()
dv = {k: torch.from_numpy(a).to(dev) for k, a in vals.items()}
ts = []
for _ in range(3):
    z, l, v = (torch.zeros(B * m, dtype=torch.float64, device=dev) for m in (nz, nl, nv))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out, y = s.solve_batch(dv, z, l, v)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
# the structured solver on the same instances
m = fb.FBstabMpc(*dims, max_batch=B)
dm = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
zm, lm, vm = (torch.zeros(B * k_, dtype=torch.float64, device=dev) for k_ in (nz, nl, nv))
tm = []
for _ in range(2):
    zm.zero_(); lm.zero_(); vm.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    om, ym = m.solve_batch(dm, zm, lm, vm)
    torch.cuda.synchronize()
    tm.append(time.perf_counter() - t0)
omn = np.frombuffer(om.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
diff = float((z - zm).abs().max() / max(1.0, float(zm.abs().max())))
# CPU oracle (sparse components, same elimination order) on a prefix
from oracle import binding as ob
ns = min(B, 256)
thr = os.cpu_count() or 1
t0 = time.perf_counter()
oo, oz, ol, ov, oy = ob.sparse_solve_batch(
    nz, nl, nv, (Hp, Hi, Gp, Gi, Ap, Ai),
    [vals[k][:ns * s.field_sizes[k]] for k in ("Hx", "f", "Gx", "h", "Ax", "b")], perm=perm,
    nthreads=thr)
tc = time.perf_counter() - t0
zz = z.cpu().numpy()[:ns * nz]
same = (oo["newton_iters"] == o["newton_iters"][:ns]) & (oo["prox_iters"] == o["prox_iters"][:ns])
print(json.dumps({
    "workload": f"{kind} OCP N={N} as a general sparse QP, {B} instances (x0 = nominal + {cfg[1]}*U)",
    "n": n, "nnzK": nnzK, "nnzL": nnzL, "path": s.path,
    "ms": [round(t, 2) for t in ts], "solves_per_s": B / min(ts) * 1e3,
    "flags": np.bincount(o["eflag"], minlength=6).tolist(),
    "newton_mean": float(o["newton_iters"].mean()),
    "mpc_solver_solves_per_s": B / min(tm), "mpc_flags": np.bincount(omn["eflag"], minlength=6).tolist(),
    "max_rel_z_diff_vs_mpc_solver": diff,
    "cpu_oracle": {"solves_per_s": ns / tc, "threads": thr, "sample": ns,
                   "same_flags": bool((oo["eflag"] == o["eflag"][:ns]).all()),
                   "same_trajectory_frac": float(same.mean()),
                   "max_rel_z_diff": float(np.abs(zz - oz).max() / max(1.0, np.abs(oz).max()))}}))
