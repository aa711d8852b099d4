"""The oracle against the REFERENCE'S OWN SOURCES (CPU).

oracle/_ref/libfbstab_ref.so is the reference's unmodified algorithm code --
fbstab/fbstab_algorithm-impl.h, fbstab/fbstab_dense.cc, fbstab/fbstab_mpc.cc and every file of
fbstab/components/ -- compiled where it lies under /root/reference against oracle/eigen_shim,
a stand-in for the Eigen API subset those files use (Eigen itself is not in this image), with
the recipe `make -C oracle _ref`.  It runs the reference's control flow, formulas, tolerances
and exit logic verbatim; only the dense linear algebra underneath is the stand-in's plain
loops instead of Eigen's kernels.

What this pins that the reference's golden vectors cannot: the iteration TRAJECTORY.  The
restated oracle (oracle/fbstab_oracle.cpp) and the reference's own code end every instance
with the same exit flag and the same Newton / proximal iteration counts, and on the MPC
path -- where the stand-in's loops and the oracle's happen to run in the same order -- with
the same BYTES.

The library is built here (the reference tree is present) and travels to the GPU box as a
prebuilt file; where neither exists the tests skip.
"""
import numpy as np
import pytest

from util import DENSE_CASES, DI2_L, DI2_V, DI2_Z, MPC_CASES, dense_case, rel_err


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libfbstab_ref.so: no reference tree and no prebuilt library")
    return oracle


def _opts(oracle, **kw):
    return oracle.default_options(abs_tol=1e-8, display_level=0, **kw)


def _same(a, b):
    return (a["eflag"] == b["eflag"]) & (a["newton_iters"] == b["newton_iters"]) & \
        (a["prox_iters"] == b["prox_iters"])


# ---- the reference's own live tests, run on the reference's own code --------------------
@pytest.mark.parametrize("name", list(DENSE_CASES))
def test_reference_sources_pass_the_dense_solver_tests(ref, name):
    """fbstab/test/fbstab_dense_unit_tests.cc:28-256 on libfbstab_ref.so: the stand-in linear
    algebra does not change what the reference's tests expect."""
    H, f, G, h, A, b, flag = dense_case(name)
    nz, nl, nv = f.size, h.size, b.size
    col = lambda M: np.ascontiguousarray(np.asarray(M, dtype=float).reshape(M.shape).T).reshape(-1)
    out, z, l, v, y = ref.ref_dense_solve_batch(nz, nl, nv, col(H), f, col(G), h, col(A), b,
                                                opts=_opts(ref))
    assert ref.EXIT_FLAGS[int(out["eflag"][0])] == flag and out["status"][0] == 0
    if name == "FeasibleQP":
        np.testing.assert_allclose(z, [0, -5], atol=1e-8)
        np.testing.assert_allclose(v, [5, 0], atol=1e-8)
    if name == "FeasibleQPwithEQ":
        np.testing.assert_allclose(z, [0.25, 0.75], atol=1e-8)
    # ... and the oracle (Eigen's LDLT restated, variant 0) walks the same trajectory
    oo, oz, ol, ov, oy = ref.dense_solve_batch(nz, nl, nv, col(H), f, col(G), h, col(A), b,
                                               opts=_opts(ref))
    assert _same(out, oo).all(), (out, oo)
    if flag == "SUCCESS" and name != "DegenerateQP":
        assert rel_err(z, oz) <= 1e-8 and rel_err(v, ov) <= 1e-8


@pytest.mark.parametrize("kind,N", MPC_CASES)
def test_reference_sources_pass_the_mpc_solver_tests(ref, fb, kind, N):
    """fbstab/test/fbstab_mpc_unit_tests.cc:15-148 on libfbstab_ref.so, incl. the quadprog
    golden of the double integrator at N = 2; the oracle returns the same bytes."""
    dims, d = fb.problems.ocp_batch(kind, N)
    seqs = [d[k] for k in fb.problems.MPC_FIELDS]
    out, z, l, v, y = ref.ref_mpc_solve_batch(*dims, seqs, opts=_opts(ref))
    assert out["eflag"][0] == 0 and out["residual"][0] <= 1e-6
    if (kind, N) == ("double_integrator", 2):
        np.testing.assert_allclose(z, DI2_Z, atol=1e-8)
        np.testing.assert_allclose(l, DI2_L, atol=1e-8)
        np.testing.assert_allclose(v, DI2_V, atol=1e-8)
    oo, oz, ol, ov, oy = ref.mpc_solve_batch(*dims, seqs, opts=_opts(ref))
    assert _same(out, oo).all()
    assert out["residual"][0] == oo["residual"][0]
    for a, b_ in ((z, oz), (l, ol), (v, ov), (y, oy)):
        assert a.tobytes() == b_.tobytes()


# ---- the bench families: trajectory of the oracle == trajectory of the reference's code ----
@pytest.mark.parametrize("kind,N,B,rho,cfg", [("servo_motor", 50, 512, 0.02, 3),
                                              ("double_integrator", 50, 512, -0.1, 3),
                                              ("spacecraft", 40, 48, 0.01, 4),
                                              # (N = 100: the reference's code, too, stops at
                                              # MAXITERATIONS after 200 Newton steps)
                                              ("spacecraft", 100, 8, 0.01, 4),
                                              ("copolymerization", 100, 32, 0.05, 4)])
def test_oracle_is_the_reference_code_on_the_mpc_families(ref, fb, kind, N, B, rho, cfg):
    """Same instances as bench.py's configs 3a, 3b, 4a40, 4b (prefixes): identical exit
    flags, iteration counts AND bytes -- including the rounding-sensitive servo family, where
    two compilations of one source already part ways on 2.4 % of the instances."""
    dims, d = fb.problems.ocp_batch(kind, N, count=B, config=cfg, rho=rho)
    seqs = [d[k] for k in fb.problems.MPC_FIELDS]
    ro, rz, rl, rv, ry = ref.ref_mpc_solve_batch(*dims, seqs, nthreads=8)
    oo, oz, ol, ov, oy = ref.mpc_solve_batch(*dims, seqs, nthreads=8)
    assert (ro["status"] == 0).all()
    assert (ro["eflag"] == oo["eflag"]).all()
    assert _same(ro, oo).all()
    assert (ro["residual"] == oo["residual"]).all()
    for a, b_ in ((rz, oz), (rl, ol), (rv, ov), (ry, oy)):
        assert a.tobytes() == b_.tobytes()


@pytest.mark.parametrize("sizes,B,cfg", [((32, 8, 64), 512, 2), ((50, 10, 100), 96, 1),
                                         ((9, 3, 4), 64, 2), ((512, 128, 1024), 2, 5)])
def test_oracle_follows_the_reference_code_on_dense_qps(ref, fb, sizes, B, cfg):
    """bench.py's configs 2, 1 and 5 (prefixes): same flags and iteration counts on every
    instance, solutions within 1e-8 (the pivoted LDL' of the stand-in and the oracle's
    restatement of Eigen's differ in the order of a few sums)."""
    nz, nl, nv = sizes
    d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=cfg)
    args = [d[k] for k in fb.problems.DENSE_FIELDS]
    ro, rz, rl, rv, ry = ref.ref_dense_solve_batch(nz, nl, nv, *args, nthreads=8)
    oo, oz, ol, ov, oy = ref.dense_solve_batch(nz, nl, nv, *args, nthreads=8)
    assert (ro["eflag"] == 0).all() and (ro["status"] == 0).all()
    assert _same(ro, oo).all()
    assert rel_err(rz, oz) <= 1e-8 and rel_err(rv, ov) <= 1e-8 and rel_err(ry, oy) <= 1e-8


def test_mixed_exit_flags_and_options_follow_the_reference_code(ref, fb):
    """Infeasible and feasible OCPs in one batch (large initial-state perturbations), an
    iteration cap, and ReliableOptions: flags and counts of the oracle are the reference
    code's, instance by instance."""
    dims, d = fb.problems.ocp_batch("servo_motor", 25, count=96, config=3, rho=0.3)
    seqs = [d[k] for k in fb.problems.MPC_FIELDS]
    for opts in (ref.default_options(display_level=0),
                 ref.default_options(display_level=0, max_newton_iters=7),
                 ref.reliable_options(display_level=0)):
        ro = ref.ref_mpc_solve_batch(*dims, seqs, opts=opts, nthreads=8)[0]
        oo = ref.mpc_solve_batch(*dims, seqs, opts=opts, nthreads=8)[0]
        assert (ro["eflag"] == oo["eflag"]).all()
        assert _same(ro, oo).all()
    assert len(np.unique(ro["eflag"])) >= 2, "the batch should mix exit flags"


# ---- the committed fixture: what the reference's code did, instance by instance ----------
def _fixture():
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location(
        "make_reference_trajectories", os.path.join(here, "golden", "make_reference_trajectories.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, "golden", "reference_trajectories.json")) as fh:
        return mod, json.load(fh)["families"]


@pytest.mark.parametrize("which", ["oracle", "ref"])
def test_committed_reference_trajectories(oracle, fb, which):
    """tests/golden/reference_trajectories.json (exit flag, Newton and proximal counts of the
    reference's own code on the first instances of every family) is reproduced by the
    reference's code where it is available, and by the oracle everywhere: the GPU parity
    tests may use the fixture as the reference's word."""
    mod, fam = _fixture()
    if which == "ref" and oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libfbstab_ref.so is not available")
    assert set(fam) == set(mod.FAMILIES)
    for name, rec in fam.items():
        out = mod.solve(oracle, fb, name, which)[0]
        assert out.size == rec["instances"]
        for f in ("eflag", "newton_iters", "prox_iters"):
            assert out[f].tolist() == rec[f], (name, f)
    assert len(set(fam["servo_motor_N25_mixed"]["eflag"])) >= 2


# ---- the reference's own unit tests and its own problem generator ------------------------
def test_the_reference_unit_tests_pass_on_the_reference_code(oracle):
    """fbstab/test/fbstab_dense_unit_tests.cc and fbstab_mpc_unit_tests.cc -- the ten live
    tests of the reference, compiled UNMODIFIED together with its ocp_generator.cc and its
    algorithm sources against oracle/eigen_shim and oracle/gtest_shim (`make -C oracle
    _ref_tests`): all pass.  This is what qualifies the stand-in linear algebra."""
    import subprocess
    binary = oracle.build_ref_tests()
    if binary is None:
        pytest.skip("oracle/_ref/ref_unit_tests: no reference tree and no prebuilt binary")
    p = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "10 tests, 0 failed" in p.stdout, p.stdout[-3000:]
    for name in ("FBstabDense.FeasibleQP", "FBstabDense.InfeasibleQP", "FBstabDense.UnboundedQP",
                 "FBstabMpc.DoubleIntegrator", "FBstabMpc.CopolymerizationReactor"):
        assert "[  OK  ] " + name in p.stdout


@pytest.mark.parametrize("kind,name,N", [(0, "double_integrator", 10), (0, "double_integrator", 2),
                                         (1, "servo_motor", 20), (1, "servo_motor", 50),
                                         (2, "spacecraft", 40), (3, "copolymerization", 70),
                                         (3, "copolymerization", 100)])
def test_the_engines_ocp_generator_is_the_reference_generator(ref, fb, kind, name, N):
    """fbstab_ocp_generate (fbstab_b200/csrc/problems.cpp, the restated constant tables of
    fbstab/test/ocp_generator.cc:73-421) returns the BYTES the reference's own OcpGenerator
    produces: all eleven sequences and x0, every problem, any horizon.  (Exact equality of
    every value; the only byte that may differ is the sign bit of a zero, -Q xtrg = -0.0 in
    the reference's q where the restated table holds +0.0.)"""
    dims, d = ref.ref_ocp_generate(kind, N)
    dims2, d2 = fb.problems.ocp_batch(name, N)
    assert tuple(dims) == tuple(dims2)
    for k in fb.problems.MPC_FIELDS:
        a, b = d[k], np.ascontiguousarray(d2[k]).reshape(-1)
        assert a.shape == b.shape and np.array_equal(a, b), k
        diff = a.tobytes() != b.tobytes()
        if diff:
            assert ((a == 0) | (a.view(np.int64) == b.view(np.int64))).all(), k

