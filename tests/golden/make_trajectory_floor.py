"""Measures the TRAJECTORY FLOOR: on how many instances of each benchmark family do
two builds of the same CPU restatement (oracle/liboracle.so without, and
oracle/liboracle_fma.so with FMA contraction) end with different iteration
counts?  Both builds evaluate the reference's formulas with correctly rounded
IEEE operations; they differ only in where a product is fused into a sum.  With
sigma = 1e-8 the Newton systems carry 1e8-scale entries, and a 1e-16 rounding
difference can move an instance across a convergence threshold.  No
implementation that is not the reference binary itself can be held to a tighter
"same trajectory" fraction than this.

Writes tests/golden/trajectory_floor.json (committed); tests/test_oracle_fma_floor.py
re-measures a subset on every CPU run and the GPU parity tests take their
thresholds from the file.   Usage: python tests/golden/make_trajectory_floor.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FAMILIES = {
    # name: (kind, spec, count, generator config, rho)
    "dense_32_8_64": ("dense", (32, 8, 64), 4096, 2, None),
    "dense_50_10_100": ("dense", (50, 10, 100), 512, 1, None),
    "dense_512_128_1024": ("dense", (512, 128, 1024), 16, 5, None),
    "servo_motor_N50": ("mpc", ("servo_motor", 50), 2048, 3, 0.02),
    "double_integrator_N50": ("mpc", ("double_integrator", 50), 2048, 3, -0.1),
    "spacecraft_N40": ("mpc", ("spacecraft", 40), 256, 4, 0.01),
    "copolymerization_N100": ("mpc", ("copolymerization", 100), 256, 4, 0.05),
    # the servo instances restated as general sparse QPs (bench.py config 3a-sparse); both
    # builds eliminate in the order the engine's symbolic analysis chooses (host only)
    "servo_motor_N50_sparse": ("sparse", ("servo_motor", 50), 512, 3, 0.02),
}


def measure(name, threads=8, count=None):
    import fbstab_b200 as fb
    from oracle import binding as ob
    kind, spec, n, cfg, rho = FAMILIES[name]
    n = count or n
    if kind == "dense":
        nz, nl, nv = spec
        d = fb.problems.random_dense_qp(nz, nl, nv, count=n, config=cfg)
        args = [d[k] for k in fb.problems.DENSE_FIELDS]
        a = ob.dense_solve_batch(nz, nl, nv, *args, nthreads=threads)
        b = ob.dense_solve_batch(nz, nl, nv, *args, nthreads=threads, fma=True)
        width = nz
    elif kind == "sparse":
        ocp, N = spec
        dims, d = fb.problems.ocp_batch(ocp, N, count=n, config=cfg, rho=rho)
        (nz, nl, nv), pat, vals = fb.problems.ocp_as_sparse_qp(dims, d, n)
        perm = fb.FBstabSparse.analyze(nz, nl, nv, pat)[3]
        V = [vals[k] for k in fb.problems.SPARSE_FIELDS]
        a = ob.sparse_solve_batch(nz, nl, nv, pat, V, perm=perm, nthreads=threads)
        b = ob.sparse_solve_batch(nz, nl, nv, pat, V, perm=perm, nthreads=threads, fma=True)
        width = nz
    else:
        ocp, N = spec
        dims, d = fb.problems.ocp_batch(ocp, N, count=n, config=cfg, rho=rho)
        seqs = [d[k] for k in fb.problems.MPC_FIELDS]
        a = ob.mpc_solve_batch(*dims, seqs, nthreads=threads)
        b = ob.mpc_solve_batch(*dims, seqs, nthreads=threads, fma=True)
        width = (N + 1) * (dims[1] + dims[2])
    oa, ob_ = a[0], b[0]
    same = ((oa["newton_iters"] == ob_["newton_iters"]) & (oa["prox_iters"] == ob_["prox_iters"]) &
            (oa["ls_backtracks"] == ob_["ls_backtracks"]) & (oa["eflag"] == ob_["eflag"]))
    Za, Zb = a[1].reshape(n, width), b[1].reshape(n, width)
    err = np.abs(Za - Zb).max(1) / np.maximum(1.0, np.abs(Za).max(1))
    okf = (oa["eflag"] == 0) & (ob_["eflag"] == 0)
    return {
        "instances": int(n),
        "same_flags": bool((oa["eflag"] == ob_["eflag"]).all()),
        "same_trajectory_frac": float(same.mean()),
        "max_abs_newton_diff": int(np.abs(oa["newton_iters"] - ob_["newton_iters"]).max()),
        "max_rel_solution_diff_same_trajectory": float(err[same & okf].max()) if (same & okf).any() else 0.0,
        "max_rel_solution_diff_off_trajectory": float(err[~same & okf].max()) if (~same & okf).any() else 0.0,
    }


if __name__ == "__main__":
    res = {}
    for name in FAMILIES:
        res[name] = measure(name, threads=os.cpu_count() or 8)
        print(name, res[name], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "trajectory_floor.json"), "w") as fh:
        json.dump({"how": "tests/golden/make_trajectory_floor.py: oracle built with -ffp-contract=off "
                          "against the same source built with -ffp-contract=fast -mfma, same instances",
                   "families": res}, fh, indent=1)
        fh.write("\n")
