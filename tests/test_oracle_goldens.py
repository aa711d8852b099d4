"""Pins the CPU oracle against every golden vector the reference's own tests
hold for the hot path.  CPU only."""
import numpy as np
import pytest

from util import (DENSE_CASES, DI2_L, DI2_V, DI2_Z, MPC_CASES, component_ocp,
                  dense_case)


def _opts(oracle):
    # every live reference test: DefaultOptions, abs_tol = 1e-8, display OFF
    return oracle.default_options(abs_tol=1e-8, display_level=0)


# ---- fbstab/test/fbstab_dense_unit_tests.cc --------------------------------
@pytest.mark.parametrize("name", list(DENSE_CASES))
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_dense_solver_cases(oracle, name, variant):
    H, f, G, h, A, b, flag = dense_case(name)
    p = oracle.Problem.dense(H, f, G, h, A, b)
    out, (z, l, v, y), _ = p.solve(_opts(oracle), variant=variant)
    assert out["flag"] == flag
    assert out["status"] == 0
    if name == "FeasibleQP":  # :51-60
        np.testing.assert_allclose(z, [0, -5], atol=1e-8)
        np.testing.assert_allclose(v, [5, 0], atol=1e-8)
    if name == "FeasibleQPwithEQ":  # :99-103
        np.testing.assert_allclose(z, [0.25, 0.75], atol=1e-8)
    if name == "DegenerateQP":  # :168-176
        assert abs(z[0] - 1) <= 1e-8 and 1 <= z[1] <= 3
        r1 = H @ z + f + A.T @ v
        r2 = np.minimum(y, v)
        assert np.linalg.norm(r1) + np.linalg.norm(r2) <= 1e-6


# ---- fbstab/components/test/dense_unit_tests.h (stale, numbers still valid) -
def _dense_component_problem(oracle):
    H = np.array([[3., 1], [1, 1]])
    A = np.array([[-1., 0], [0, 1]])
    return oracle.Problem.dense(H, [1., 6], np.zeros((0, 2)), [], A, [0., -1]), H, A


def test_dense_variable(oracle):  # dense_unit_tests.h:51-92
    p, H, A = _dense_component_problem(oracle)
    b = np.array([0., -1])
    xz = np.ones(2)
    np.testing.assert_array_equal(p.margin(xz), b - A @ xz)
    # y <- a*x + y keeps y.y == b - A*y.z
    a = 0.35
    x = (np.ones(2), np.zeros(0), np.ones(2), p.margin(np.ones(2)))
    yv = (-np.ones(2), np.zeros(0), -np.ones(2), p.margin(-np.ones(2)))
    z, l, v, y = p.variable_axpy(a, x, yv)
    np.testing.assert_allclose(z, a * 1 - 1, rtol=1e-15)
    np.testing.assert_allclose(v, a * 1 - 1, rtol=1e-15)
    np.testing.assert_allclose(y, b - A @ z, rtol=1e-15, atol=1e-15)


def test_dense_inner_residual(oracle):  # dense_unit_tests.h:100-131
    p, _, _ = _dense_component_problem(oracle)
    rz, rl, rv, _ = p.residual("inner", ([1, 5], [], [0.4, 2]),
                               ([-5, 6], [], [-9, 1]), sigma=0.5)
    np.testing.assert_allclose(rz, [11.6, 13.5], atol=1e-14)
    np.testing.assert_allclose(rv, [0.480683041678573, -8.88473245759182],
                               atol=1e-14)


def test_dense_natural_residual(oracle):  # dense_unit_tests.h:137-160
    p, _, _ = _dense_component_problem(oracle)
    rz, rl, rv, _ = p.residual("natural", ([1, 5], [], [0.4, 2]))
    np.testing.assert_allclose(rz, [8.6, 14.0], atol=1e-14)
    np.testing.assert_allclose(rv, [0.4, -6], atol=1e-14)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_dense_linear_solver_residual(oracle, variant):  # dense_unit_tests.h:172-213
    p, H, A = _dense_component_problem(oracle)
    sigma = 0.5
    r = (np.ones(2), np.zeros(0), np.ones(2))
    rc, (dz, dl, dv, dy), gamma, mus = p.linear_solve(
        ([1, 5], [], [0.4, 2]), ([-5, 6], [], [-9, 1]), sigma, r, variant=variant)
    assert rc == 0
    K = np.block([[H + sigma * np.eye(2), A.T],
                  [-np.diag(gamma) @ A, np.diag(mus)]])
    res = K @ np.concatenate([dz, dv]) - np.concatenate([r[0], r[2]])
    assert np.linalg.norm(res) <= 1e-12
    np.testing.assert_allclose(dy, np.array([0., -1]) - A @ dz, atol=1e-14)


def test_dense_infeasibility_certificates(oracle):  # dense_unit_tests.h:223-293
    H, f, G, h, A, b, _ = dense_case("InfeasibleQP")
    p = oracle.Problem.dense(H, f, G, h, A, b)
    assert p.feasibility([0, 0], [], [1, 0, 0, 1, 1], 1e-8) == 1  # primal infeasible
    H, f, G, h, A, b, _ = dense_case("UnboundedQP")
    p = oracle.Problem.dense(H, f, G, h, A, b)
    assert p.feasibility([0, 1], [], [0, 0, 0, 0], 1e-8) == 2  # dual infeasible


# ---- fbstab/components/test/mpc_component_unit_tests.h ----------------------
@pytest.fixture()
def comp_ocp(oracle, fb):
    dims, d = component_ocp(fb)
    p = oracle.Problem.mpc(*dims, *[d[k] for k in fb.problems.MPC_FIELDS])
    return dims, p


def test_mpc_gemv_goldens(comp_ocp):  # mpc_component_unit_tests.h:99-214
    _, p = comp_ocp
    z = np.arange(1, p.nz + 1, dtype=float)
    l = np.arange(1, p.nl + 1, dtype=float)
    v = np.arange(1, p.nv + 1, dtype=float)
    np.testing.assert_array_equal(p.gemv("H", z, 1.0, 0.0, np.zeros(p.nz)),
                                  [5, 2, 10, 14, 5, 22, 23, 8, 34])
    np.testing.assert_array_equal(
        p.gemv("A", z, 1.0, 0.0, np.zeros(p.nv)),
        [-1, -2, 1, 2, -3, 3, -4, -5, 4, 5, -6, 6, -7, -8, 7, 8, -9, 9])
    np.testing.assert_array_equal(p.gemv("G", z, 1.0, 0.0, np.zeros(p.nl)),
                                  [-1, -2, -1, 0, 2, 3])
    np.testing.assert_array_equal(p.gemv("GT", l, 1.0, 0.0, np.zeros(p.nz)),
                                  [2, 5, 4, 2, 7, 6, -5, -6, 0])
    np.testing.assert_array_equal(p.gemv("AT", v, 1.0, 0.0, np.zeros(p.nz)),
                                  [2, 2, 1, 2, 2, 1, 2, 2, 1])


def test_mpc_axpy_goldens(comp_ocp):  # mpc_component_unit_tests.h:219-272
    _, p = comp_ocp
    np.testing.assert_array_equal(
        p.axpy("f", 2.0, [5, 2, 10, 14, 5, 22, 23, 8, 34]),
        [1, 2, 10, 10, 5, 22, 19, 8, 34])
    np.testing.assert_array_equal(p.axpy("h", 2.0, [-1, -2, -1, 0, 2, 3]),
                                  [-1, -2, -1, 0, 2, 3])
    np.testing.assert_array_equal(
        p.axpy("b", 2.0, [-1, -2, 1, 2, -3, 3, -4, -5, 4, 5, -6, 6, -7, -8, 7, 8, -9, 9]),
        [-1, -2, 5, 6, -1, 5, -4, -5, 8, 9, -4, 8, -7, -8, 11, 12, -7, 11])


def test_mpc_variable_axpy(comp_ocp):  # mpc_component_unit_tests.h:277-311
    _, p = comp_ocp
    one = lambda n: np.ones(n)
    x = (one(p.nz), one(p.nl), one(p.nv), p.margin(one(p.nz)))
    y = (one(p.nz), one(p.nl), one(p.nv), p.margin(one(p.nz)))
    z, l, v, yy = p.variable_axpy(-2.0, y, x)
    np.testing.assert_array_equal(z, -one(p.nz))
    np.testing.assert_array_equal(l, -one(p.nl))
    np.testing.assert_array_equal(v, -one(p.nv))
    np.testing.assert_array_equal(yy, [-1, -1, 3, 3, 0, 2] * 3)


def test_mpc_inner_residual(comp_ocp):  # mpc_component_unit_tests.h:316-355
    _, p = comp_ocp
    two = lambda n: 2.0 * np.ones(n)
    rz, rl, rv, _ = p.residual("inner", (two(p.nz), two(p.nl), two(p.nv)),
                               (-two(p.nz), -two(p.nl), -two(p.nv)), sigma=1.0)
    np.testing.assert_allclose(rz, [8, 8, 14, 8, 8, 14, 6, 4, 12], atol=1e-14)
    np.testing.assert_allclose(rl, [6, 6, 2, 2, 2, 2], atol=1e-14)
    np.testing.assert_allclose(
        rv, [2.19167244568008, 2.19167244568008, 1.85147084275040,
             1.85147084275040, 2.33389560518351, 1.62472628830921] * 3, atol=1e-14)


def test_mpc_feasibility_sanity(comp_ocp):  # mpc_component_unit_tests.h:359-373
    _, p = comp_ocp
    assert p.feasibility(np.zeros(p.nz), np.zeros(p.nl), np.zeros(p.nv), 1e-8) == 0


def test_riccati_recursion(comp_ocp):  # mpc_component_unit_tests.h:386-461
    _, p = comp_ocp
    f = lambda n, a: a * np.ones(n)
    x = (f(p.nz, 1), f(p.nl, 2), f(p.nv, 4))
    xb = (f(p.nz, 2), f(p.nl, 1), f(p.nv, 3))
    r = (f(p.nz, 2.5), f(p.nl, 2.5), f(p.nv, 2.5))
    sigma = 1.0
    rc, (dz, dl, dv, dy), gamma, mus = p.linear_solve(x, xb, sigma, r)
    assert rc == 0
    r1 = p.gemv("H", dz, 1.0, 1.0, np.zeros(p.nz)) + sigma * dz
    r1 = p.gemv("GT", dl, 1.0, 1.0, r1)
    r1 = p.gemv("AT", dv, 1.0, 1.0, r1)
    np.testing.assert_allclose(r[0] - r1, 0, atol=1e-14)
    r2 = p.gemv("G", dz, -1.0, 1.0, np.zeros(p.nl)) + sigma * dl
    np.testing.assert_allclose(r[1] - r2, 0, atol=1e-14)
    r3 = gamma * p.gemv("A", dz, -1.0, 1.0, np.zeros(p.nv)) + mus * dv
    np.testing.assert_allclose(r[2] - r3, 0, atol=1e-14)
    r4 = p.axpy("b", 1.0, p.gemv("A", dz, -1.0, 1.0, np.zeros(p.nv)))
    np.testing.assert_allclose(dy - r4, 0, atol=1e-14)


# ---- fbstab/test/fbstab_mpc_unit_tests.cc -----------------------------------
@pytest.mark.parametrize("kind,N", MPC_CASES)
def test_mpc_solver_cases(oracle, fb, kind, N):
    dims, d = fb.problems.ocp_batch(kind, N)
    p = oracle.Problem.mpc(*dims, *[d[k] for k in fb.problems.MPC_FIELDS])
    out, (z, l, v, y), _ = p.solve(_opts(oracle))
    assert out["flag"] == "SUCCESS"      # :29 etc.
    assert out["residual"] <= 1e-6       # :30 etc.
    if (kind, N) == ("double_integrator", 2):  # :49-59
        np.testing.assert_allclose(z, DI2_Z, atol=1e-8)
        np.testing.assert_allclose(l, DI2_L, atol=1e-8)
        np.testing.assert_allclose(v, DI2_V, atol=1e-8)


# ---- options: fbstab_algorithm-impl.h:7-74 -----------------------------------
def test_options_defaults_and_clamps(oracle):
    o = oracle.default_options()
    assert (o.sigma0, o.alpha, o.beta, o.eta, o.delta) == (1e-8, 0.95, 0.75, 1e-8, 0.2)
    assert (o.max_newton_iters, o.max_prox_iters, o.max_inner_iters,
            o.max_linesearch_iters) == (200, 30, 50, 20)
    r = oracle.reliable_options()
    assert (r.sigma0, r.beta, r.abs_tol, r.max_newton_iters,
            r.nonmonotone_linesearch) == (1e-4, 0.9, 1e-4, 500, 0)
    o.alpha, o.sigma0, o.max_prox_iters = 5.0, 1.0, -3
    assert oracle.validate_options(o) == 0
    assert (o.alpha, o.sigma0, o.max_prox_iters) == (0.999, 1e-6, 1)


def test_trajectory_is_recorded(oracle):
    H, f, G, h, A, b, _ = dense_case("FeasibleQP")
    p = oracle.Problem.dense(H, f, G, h, A, b)
    out, _, traj = p.solve(_opts(oracle), traj_cap=64)
    newton = traj[traj[:, 0] == 1]
    assert len(newton) == out["newton_iters"]
    assert int(newton[:, 6].sum()) == out["ls_backtracks"]
    assert (traj[:, 0] == 0).sum() == out["prox_iters"] + 1
