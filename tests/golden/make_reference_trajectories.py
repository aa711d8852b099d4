"""Golden TRAJECTORIES from the reference's own code.

Runs oracle/_ref/libfbstab_ref.so -- the reference's unmodified algorithm sources compiled
against oracle/eigen_shim (see oracle/README.md) -- on the first instances of every benchmark
family and writes, per instance, the exit flag and the Newton / proximal iteration counts to
tests/golden/reference_trajectories.json.  The fixture travels with the repository, so the GPU
parity tests compare the CUDA engine with what the reference's code did, instance by instance,
on a box that has neither the reference tree nor the library.

Needs the reference tree (or a prebuilt oracle/_ref).  Usage:
    python tests/golden/make_reference_trajectories.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FAMILIES = {
    # name: (kind, spec, count, generator config, rho)  -- the instances of bench.py's configs
    "dense_32_8_64": ("dense", (32, 8, 64), 256, 2, None),
    "dense_50_10_100": ("dense", (50, 10, 100), 64, 1, None),
    "servo_motor_N50": ("mpc", ("servo_motor", 50), 512, 3, 0.02),
    "double_integrator_N50": ("mpc", ("double_integrator", 50), 256, 3, -0.1),
    "spacecraft_N40": ("mpc", ("spacecraft", 40), 64, 4, 0.01),
    "copolymerization_N100": ("mpc", ("copolymerization", 100), 32, 4, 0.05),
    # large initial-state perturbations: feasible and infeasible OCPs in one batch
    "servo_motor_N25_mixed": ("mpc", ("servo_motor", 25), 96, 3, 0.3),
}


def family_data(fb, name):
    kind, spec, n, cfg, rho = FAMILIES[name]
    if kind == "dense":
        nz, nl, nv = spec
        d = fb.problems.random_dense_qp(nz, nl, nv, count=n, config=cfg)
        return kind, spec, d
    ocp, N = spec
    dims, d = fb.problems.ocp_batch(ocp, N, count=n, config=cfg, rho=rho)
    return kind, dims, d


def solve(ob, fb, name, which, threads=8):
    """which: 'ref' (the reference's own code) or 'oracle' (the restatement)."""
    kind, dims, d = family_data(fb, name)
    if kind == "dense":
        args = [d[k] for k in fb.problems.DENSE_FIELDS]
        f = ob.ref_dense_solve_batch if which == "ref" else ob.dense_solve_batch
        return f(*dims, *args, nthreads=threads)
    seqs = [d[k] for k in fb.problems.MPC_FIELDS]
    f = ob.ref_mpc_solve_batch if which == "ref" else ob.mpc_solve_batch
    return f(*dims, seqs, nthreads=threads)


if __name__ == "__main__":
    import fbstab_b200 as fb
    from oracle import binding as ob
    if ob.ref_lib() is None:
        raise SystemExit("oracle/_ref/libfbstab_ref.so is not available (no reference tree)")
    res = {}
    for name in FAMILIES:
        out = solve(ob, fb, name, "ref", threads=os.cpu_count() or 8)[0]
        res[name] = {"instances": int(out.size), "eflag": out["eflag"].tolist(),
                     "newton_iters": out["newton_iters"].tolist(),
                     "prox_iters": out["prox_iters"].tolist()}
        print(name, np.bincount(out["eflag"], minlength=6).tolist(),
              "newton mean %.2f" % out["newton_iters"].mean(), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "reference_trajectories.json"), "w") as fh:
        json.dump({"how": "tests/golden/make_reference_trajectories.py: the reference's own "
                          "algorithm sources (oracle/_ref/libfbstab_ref.so, compiled against "
                          "oracle/eigen_shim), default options, cold start, the first "
                          "instances of each family",
                   "families": res}, fh, separators=(",", ":"))
        fh.write("\n")
