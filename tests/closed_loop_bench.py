"""Closed-loop (receding-horizon) MPC throughput: B plants x T control steps,
data and iterates resident on the device, warm-started vs cold-started, next to
the same loop around the CPU oracle on a sample of the plants.
Usage: python tests/closed_loop_bench.py [kind N B T rho] ...   (default: the config-3 OCPs)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fbstab_b200 as fb  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle.closed_loop_ref import closed_loop_reference  # noqa: E402

CASES = [("servo_motor", 50, 16384, 40, 0.02), ("double_integrator", 50, 16384, 40, -0.1)]


def run(kind, N, B, T, rho, cpu_sample=256):
    dims, d = fb.problems.ocp_batch(kind, N, count=B, config=3, rho=rho)
    # ONE copy of the plant's OCP data, B initial states (the C-ABI's shared-data mode)
    s = fb.FBstabMpc(*dims, max_batch=1)
    one = {k: (a if k == "x0" else a[:s.field_sizes[k]].copy()) for k, a in d.items()}
    del s
    cl = fb.ClosedLoopMpc(dims, one, shared=True, max_steps=T)
    cl.run(2)  # warm-up
    warm = cl.run(T)
    cold = cl.run(T, warm_start=False)
    nthreads = os.cpu_count() or 1
    ds = {k: a[:cpu_sample * (a.size // B)].copy() for k, a in d.items()}

    def solve(dims_, dd, x0):
        return ob.mpc_solve_batch(*dims_, [dd[k] for k in fb.problems.MPC_FIELDS], x0=x0,
                                  nthreads=nthreads)[:4]

    t0 = time.perf_counter()
    ref = closed_loop_reference(dims, ds, T, solve)
    cpu_s = time.perf_counter() - t0
    same = (warm["eflag"][:, :cpu_sample] == ref["eflag"]).all()
    ok = (ref["eflag"] == 0).all(axis=0)
    err = np.abs(warm["U"][:cpu_sample][ok] - ref["U"][ok]).max() / max(1.0, np.abs(ref["U"][ok]).max())
    print(json.dumps({
        "workload": f"closed-loop {kind} OCP N={N}, {B} plants x {T} control steps",
        "path": warm["path"],
        "warm": {"ms": round(warm["ms"], 2), "control_steps_per_s": round(B * T / (warm["ms"] * 1e-3)),
                 "newton_per_step": round(float(warm["newton_iters"][1:].mean()), 2),
                 "flags": np.bincount(warm["eflag"].ravel(), minlength=6).tolist()},
        "cold": {"ms": round(cold["ms"], 2), "control_steps_per_s": round(B * T / (cold["ms"] * 1e-3)),
                 "newton_per_step": round(float(cold["newton_iters"][1:].mean()), 2)},
        "cpu_port": {"control_steps_per_s": round(cpu_sample * T / cpu_s), "threads": nthreads,
                     "sample": f"first {cpu_sample} plants, warm-started",
                     "same_flags": bool(same), "max_rel_input_diff": float(err)},
    }), flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    cases = CASES if not a else [(a[0], int(a[1]), int(a[2]), int(a[3]), float(a[4]))]
    for c in cases:
        run(*c)
