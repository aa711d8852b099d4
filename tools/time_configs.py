"""Times every BASELINE.json config on one GPU (inputs resident in HBM, CUDA
events) and prints solves/s, exit-flag histogram and iteration statistics.
Usage: python tools/time_configs.py [name ...] [--scale F]   (F scales batch sizes)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb  # noqa: E402

RHO = {"servo_motor": 0.02, "double_integrator": -0.1, "spacecraft": 0.05,
       "copolymerization": 0.05}
CONFIGS = {
    "cfg2_dense32": ("dense", (32, 8, 64), 65536, 2),
    "cfg3a_servo50": ("mpc", ("servo_motor", 50), 16384, 3),
    "cfg3b_dint50": ("mpc", ("double_integrator", 50), 16384, 3),
    "cfg4a_spacecraft100": ("mpc", ("spacecraft", 100), 4096, 4),
    "cfg4b_copoly100": ("mpc", ("copolymerization", 100), 4096, 4),
    "cfg5_dense512": ("dense", (512, 128, 1024), 1024, 5),
    "cfg1_dense50": ("dense", (50, 10, 100), 4096, 1),
}


def run(name, scale, reps=3):
    kind, spec, B, cfg = CONFIGS[name]
    B = max(1, int(B * scale))
    dev = torch.device("cuda:0")
    if kind == "dense":
        nz, nl, nv = spec
        d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=cfg, nthreads=16)
        s = fb.FBstabDense(nz, nl, nv, max_batch=B)
    else:
        okind, N = spec
        dims, d = fb.problems.ocp_batch(okind, N, count=B, config=cfg, rho=RHO[okind])
        s = fb.FBstabMpc(*dims, max_batch=B)
    dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
    best = 1e30
    for _ in range(reps):
        z = torch.zeros(B * s.nz, dtype=torch.float64, device=dev)
        l = torch.zeros(B * s.nl, dtype=torch.float64, device=dev)
        v = torch.zeros(B * s.nv, dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out, y = s.solve_batch(dd, z, l, v)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
    rec = {"config": name, "batch": B, "ms": round(best, 3),
           "solves_per_s": round(B / (best * 1e-3), 1), "path": s.path,
           "flags": np.bincount(o["eflag"], minlength=6).tolist(),
           "status_nonzero": int((o["status"] != 0).sum()),
           "newton_mean": float(o["newton_iters"].mean()),
           "newton_max": int(o["newton_iters"].max()),
           "prox_mean": float(o["prox_iters"].mean()),
           "backtracks_mean": float(o["ls_backtracks"].mean())}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:]]
    scale, reps = 1.0, 3
    if "--scale" in args:
        i = args.index("--scale")
        scale = float(args[i + 1])
        del args[i:i + 2]
    if "--reps" in args:
        i = args.index("--reps")
        reps = int(args[i + 1])
        del args[i:i + 2]
    for n in (args or list(CONFIGS)):
        run(n, scale, reps)
