"""FBstabSparse: sparse QPs with a common pattern.

The reference has no sparse solver yet -- it plans one (ROADMAP.md:10) on the LDL'
wrapper tools/qdldl/qdldl_wrapper.h:19-84 -- so what its tree pins is (1) the wrapper's
own known-answer test (tools/qdldl/test/qdldl_test.cc:33-60) and (2) the QPs of the live
dense and MPC solver tests, which a sparse solver must solve to the same answers.  CPU
tests pin the oracle's restated QDLDL and sparse components against those; the GPU tests
compare the lane-per-instance CUDA path with that oracle through the C-ABI.
"""
import numpy as np
import pytest

from util import DENSE_CASES, dense_case, rel_err


# ---- helpers: sparse formats ----------------------------------------------------------
def csc(M, upper=False, keep=None):
    """Compressed-column pattern and values of a 2-D array (entries with keep != 0, or
    the non-zeros; upper: only i <= j)."""
    M = np.asarray(M, dtype=float)
    K = (M != 0) if keep is None else (np.asarray(keep) != 0)
    p, i, x = [0], [], []
    for c in range(M.shape[1]):
        for r in range(M.shape[0]):
            if K[r, c] and (not upper or r <= c):
                i.append(r)
                x.append(M[r, c])
        p.append(len(i))
    return np.array(p, np.int32), np.array(i, np.int32), np.array(x, float)


def sparse_from_dense(H, G, A):
    Hp, Hi, Hx = csc(H, upper=True)
    Gp, Gi, Gx = csc(G) if G.shape[0] else (np.zeros(H.shape[0] + 1, np.int32),
                                            np.zeros(0, np.int32), np.zeros(0))
    Ap, Ai, Ax = csc(A)
    return (Hp, Hi, Gp, Gi, Ap, Ai), (Hx, Gx, Ax)


def random_sparse_qp(rng, nz, nl, nv, band=3, count=1):
    """`count` strictly convex, strictly feasible QPs with ONE pattern: banded H,
    G and A with a few entries per row (the recipe of SURVEY.md 8(d) on a pattern)."""
    mask_H = np.zeros((nz, nz), bool)
    for i in range(nz):
        mask_H[i, max(0, i - band):i + band + 1] = True
    mask_G = np.zeros((nl, nz), bool)
    for r in range(nl):
        mask_G[r, rng.choice(nz, size=min(nz, 4), replace=False)] = True
    mask_A = np.zeros((nv, nz), bool)
    for r in range(nv):
        mask_A[r, rng.choice(nz, size=min(nz, 3), replace=False)] = True
    dense, pat = [], None
    for _ in range(count):
        M = rng.standard_normal((nz, nz)) * mask_H
        H = M @ M.T / nz + 1e-2 * np.eye(nz)
        keepH = (np.abs(M) @ np.abs(M.T)) > 0
        G = rng.standard_normal((nl, nz)) * mask_G
        A = rng.standard_normal((nv, nz)) * mask_A
        zs = rng.standard_normal(nz)
        h = G @ zs
        b = A @ zs + rng.uniform(0, 1, nv)
        f = rng.standard_normal(nz)
        dense.append((H, f, G, h, A, b))
        if pat is None:
            Hp, Hi, _ = csc(H, upper=True, keep=keepH)
            Gp, Gi, _ = csc(G, keep=mask_G)
            Ap, Ai, _ = csc(A, keep=mask_A)
            pat = (Hp, Hi, Gp, Gi, Ap, Ai), (keepH, mask_G, mask_A)
    vals = {k: [] for k in ("Hx", "f", "Gx", "h", "Ax", "b")}
    for H, f, G, h, A, b in dense:
        vals["Hx"].append(csc(H, upper=True, keep=pat[1][0])[2])
        vals["Gx"].append(csc(G, keep=pat[1][1])[2])
        vals["Ax"].append(csc(A, keep=pat[1][2])[2])
        vals["f"].append(f)
        vals["h"].append(h)
        vals["b"].append(b)
    return pat[0], {k: np.ascontiguousarray(np.concatenate(v)) for k, v in vals.items()}, dense


def ocp_as_qp(dims, d, inst=0):
    """One instance of the OCP wire format as a dense QP (H, f, G, h, A, b): the sign and
    ordering conventions of mpc_data.cc:17-289."""
    N, nx, nu, nc = dims
    K, ns = N + 1, nx + nu
    sz = {"Q": K * nx * nx, "R": K * nu * nu, "S": K * nu * nx, "q": K * nx, "r": K * nu,
          "A": N * nx * nx, "B": N * nx * nu, "c": N * nx, "E": K * nc * nx, "L": K * nc * nu,
          "d": K * nc, "x0": nx}
    g = {k: d[k][inst * n:(inst + 1) * n] for k, n in sz.items()}
    mat = lambda a, i, r, c: a[i * r * c:(i + 1) * r * c].reshape(c, r).T  # column-major
    nz, nl, nv = K * ns, K * nx, K * nc
    H, G, A = np.zeros((nz, nz)), np.zeros((nl, nz)), np.zeros((nv, nz))
    f, h, b = np.zeros(nz), np.zeros(nl), np.zeros(nv)
    for i in range(K):
        o = i * ns
        H[o:o + nx, o:o + nx] = mat(g["Q"], i, nx, nx)
        H[o + nx:o + ns, o + nx:o + ns] = mat(g["R"], i, nu, nu)
        S = mat(g["S"], i, nu, nx)
        H[o + nx:o + ns, o:o + nx] = S
        H[o:o + nx, o + nx:o + ns] = S.T
        f[o:o + nx] = g["q"][i * nx:(i + 1) * nx]
        f[o + nx:o + ns] = g["r"][i * nu:(i + 1) * nu]
        A[i * nc:(i + 1) * nc, o:o + nx] = mat(g["E"], i, nc, nx)
        A[i * nc:(i + 1) * nc, o + nx:o + ns] = mat(g["L"], i, nc, nu)
        b[i * nc:(i + 1) * nc] = -g["d"][i * nc:(i + 1) * nc]
        G[i * nx:(i + 1) * nx, o:o + nx] = -np.eye(nx)
        if i == 0:
            h[:nx] = -g["x0"]
        else:
            p = (i - 1) * ns
            G[i * nx:(i + 1) * nx, p:p + nx] = mat(g["A"], i - 1, nx, nx)
            G[i * nx:(i + 1) * nx, p + nx:p + ns] = mat(g["B"], i - 1, nx, nu)
            h[i * nx:(i + 1) * nx] = -g["c"][(i - 1) * nx:i * nx]
    return H, f, G, h, A, b


def _opts(mod, **kw):
    return mod.default_options(abs_tol=1e-8, display_level=0, **kw)


# ---- CPU: the oracle's QDLDL and sparse components against the reference's numbers -----
def test_qdldl_known_answer(oracle):
    """tools/qdldl/test/qdldl_test.cc:33-60 (the upstream QDLDL example)."""
    Ap = [0, 1, 2, 4, 5, 6, 8, 10, 12, 14, 17]
    Ai = [0, 1, 1, 2, 3, 4, 1, 5, 0, 6, 3, 7, 6, 8, 1, 2, 9]
    Ax = [1.0, 0.460641, -0.121189, 0.417928, 0.177828, 0.1, -0.0290058, -1.0, 0.350321,
          -0.441092, -0.0845395, -0.316228, 0.178663, -0.299077, 0.182452, -1.56506, -0.1]
    b = np.arange(1.0, 11.0)
    rc, x = oracle.qdldl_solve(10, Ap, Ai, Ax, b)
    assert rc == 0
    A = np.zeros((10, 10))
    for c in range(10):
        for e in range(Ap[c], Ap[c + 1]):
            A[Ai[e], c] = A[c, Ai[e]] = Ax[e]
    assert np.linalg.norm(A @ x - b) <= 1e-12


@pytest.mark.parametrize("name", list(DENSE_CASES))
def test_sparse_oracle_on_the_dense_solver_cases(oracle, name):
    """fbstab_dense_unit_tests.cc:28-256 solved through the sparse components."""
    H, f, G, h, A, b, flag = dense_case(name)
    pat, (Hx, Gx, Ax) = sparse_from_dense(H, G, A)
    p = oracle.Problem.sparse(f.size, h.size, b.size, pat[0], pat[1], Hx, f, pat[2], pat[3], Gx,
                              h, pat[4], pat[5], Ax, b)
    out, (z, l, v, y), _ = p.solve(_opts(oracle))
    assert out["flag"] == flag and out["status"] == 0
    if name == "FeasibleQP":
        np.testing.assert_allclose(z, [0, -5], atol=1e-8)
        np.testing.assert_allclose(v, [5, 0], atol=1e-8)
    if name == "FeasibleQPwithEQ":
        np.testing.assert_allclose(z, [0.25, 0.75], atol=1e-8)
    if name == "DegenerateQP":
        assert abs(z[0] - 1) <= 1e-8 and 1 <= z[1] <= 3


def test_sparse_oracle_components_match_dense(oracle):
    """gemv / residual / Newton step of the sparse components == the dense ones (pinned to
    the reference's goldens in test_oracle_goldens.py) on a random sparse QP."""
    rng = np.random.default_rng(5)
    nz, nl, nv = 12, 3, 15
    pat, vals, dense = random_sparse_qp(rng, nz, nl, nv)
    H, f, G, h, A, b = dense[0]
    ps = oracle.Problem.sparse(nz, nl, nv, pat[0], pat[1], vals["Hx"], f, pat[2], pat[3],
                               vals["Gx"], h, pat[4], pat[5], vals["Ax"], b,
                               perm=np.concatenate([nz + nl + rng.permutation(nv),
                                                    rng.permutation(nz), nz + rng.permutation(nl)]))
    pd = oracle.Problem.dense(H, f, G, h, A, b)
    x = rng.standard_normal(nz)
    for op, n_in, n_out in (("H", nz, nz), ("A", nz, nv), ("AT", nv, nz), ("G", nz, nl),
                            ("GT", nl, nz)):
        xi, y0 = rng.standard_normal(n_in), rng.standard_normal(n_out)
        np.testing.assert_allclose(ps.gemv(op, xi, 0.7, -0.3, y0), pd.gemv(op, xi, 0.7, -0.3, y0),
                                   atol=1e-13)
    xv = (x, rng.standard_normal(nl), np.abs(rng.standard_normal(nv)))
    xb = (rng.standard_normal(nz), rng.standard_normal(nl), np.abs(rng.standard_normal(nv)))
    for a_, b_ in zip(ps.residual("inner", xv, xb, sigma=0.5), pd.residual("inner", xv, xb, sigma=0.5)):
        np.testing.assert_allclose(a_, b_, atol=1e-12)
    r = (rng.standard_normal(nz), rng.standard_normal(nl), rng.standard_normal(nv))
    rc1, dx1, g1, m1 = ps.linear_solve(xv, xb, 1e-4, r)
    rc2, dx2, g2, m2 = pd.linear_solve(xv, xb, 1e-4, r)
    assert rc1 == 0 and rc2 == 0
    np.testing.assert_allclose(g1, g2, rtol=1e-15)
    for a_, b_ in zip(dx1, dx2):
        np.testing.assert_allclose(a_, b_, rtol=1e-8, atol=1e-9)


def test_sparse_oracle_solves_match_dense(oracle):
    rng = np.random.default_rng(11)
    nz, nl, nv = 20, 4, 30
    pat, vals, dense = random_sparse_qp(rng, nz, nl, nv, count=6)
    out, z, l, v, y = oracle.sparse_solve_batch(nz, nl, nv, pat, [vals[k] for k in
                                                ("Hx", "f", "Gx", "h", "Ax", "b")], nthreads=2)
    for i, (H, f, G, h, A, b) in enumerate(dense):
        o, (zd, ld, vd, yd), _ = oracle.Problem.dense(H, f, G, h, A, b).solve()
        assert out["eflag"][i] == o["eflag"] == 0
        assert rel_err(z[i * nz:(i + 1) * nz], zd) <= 1e-6
        assert rel_err(v[i * nv:(i + 1) * nv], vd) <= 1e-6


def test_sparse_oracle_on_an_ocp(oracle, fb):
    """fbstab_mpc_unit_tests.cc:15-60: DoubleIntegrator N=2 as a general sparse QP reaches
    the quadprog golden."""
    from util import DI2_L, DI2_V, DI2_Z
    dims, d = fb.problems.ocp_batch("double_integrator", 2)
    H, f, G, h, A, b = ocp_as_qp(dims, d)
    pat, (Hx, Gx, Ax) = sparse_from_dense(H, G, A)
    p = oracle.Problem.sparse(f.size, h.size, b.size, pat[0], pat[1], Hx, f, pat[2], pat[3], Gx,
                              h, pat[4], pat[5], Ax, b)
    out, (z, l, v, y), _ = p.solve(_opts(oracle))
    assert out["flag"] == "SUCCESS"
    np.testing.assert_allclose(z, DI2_Z, atol=1e-8)
    np.testing.assert_allclose(l, DI2_L, atol=1e-8)
    np.testing.assert_allclose(v, DI2_V, atol=1e-8)


def test_sparse_pattern_validation_and_no_cpu_fallback(fb):
    """The symbolic analysis rejects inconsistent patterns with FBSTAB_ERR_INVALID before
    any device work; a valid pattern without a GPU fails loudly (no CPU fallback)."""
    import torch
    Hp, Hi = np.array([0, 1, 3], np.int32), np.array([0, 0, 1], np.int32)
    Ap, Ai = np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32)
    nog = (np.zeros(3, np.int32), np.zeros(0, np.int32))
    bad = [((Hp, np.array([0, 1, 0], np.int32)) + nog + (Ap, Ai), "strictly increasing"),
           ((np.array([0, 1, 2], np.int32), np.array([1, 1], np.int32)) + nog + (Ap, Ai),
            "upper triangle"),
           ((Hp, Hi) + nog + (Ap, np.array([0, 5], np.int32)), "out of range")]
    for pat, what in bad:
        with pytest.raises(fb.FbstabError) as e:
            fb.FBstabSparse(2, 0, 2, pat)
        assert e.value.code == fb.capi.ERR_INVALID and what in str(e.value), str(e.value)
    with pytest.raises(fb.FbstabError) as e:
        fb.FBstabSparse(2, 0, 2, (Hp, Hi) + nog + (Ap, Ai), perm=[0, 1, 1, 3])
    assert "permutation" in str(e.value)
    with pytest.raises(RuntimeError):
        fb.FBstabSparse(0, 0, 2, (Hp, Hi) + nog + (Ap, Ai))
    if not torch.cuda.is_available():
        with pytest.raises(fb.FbstabError) as e:
            fb.FBstabSparse(2, 0, 2, (Hp, Hi) + nog + (Ap, Ai))
        assert e.value.code == fb.capi.ERR_NOGPU


# ---- GPU: the lane-per-instance CUDA path against the oracle ---------------------------
def test_sparse_analyze_runs_on_the_host(fb, oracle):
    """fbstab_sparse_analyze: the handle's symbolic analysis without a handle or a device --
    a valid elimination order of [z; l; w], the fill it implies, the caller's order kept,
    bad patterns rejected; the oracle solves in that order."""
    nz, nl, nv, B = 24, 4, 40, 3
    rng = np.random.default_rng(3)
    pat, vals, _ = random_sparse_qp(rng, nz, nl, nv, count=B)
    n, nnzK, nnzL, perm = fb.FBstabSparse.analyze(nz, nl, nv, pat)
    assert n == nz + nl + nv and sorted(perm.tolist()) == list(range(n))
    assert nnzK >= n and nnzL > 0
    # the natural order [w; z; l] fills at least as much as the minimum-degree order
    natural = np.concatenate([nz + nl + np.arange(nv), np.arange(nz), nz + np.arange(nl)])
    n2, k2, l2, p2 = fb.FBstabSparse.analyze(nz, nl, nv, pat, perm=natural)
    assert (p2 == natural).all() and k2 == nnzK and l2 >= nnzL
    V = [vals[k] for k in ("Hx", "f", "Gx", "h", "Ax", "b")]
    oa = oracle.sparse_solve_batch(nz, nl, nv, pat, V, perm=perm)
    ob_ = oracle.sparse_solve_batch(nz, nl, nv, pat, V, perm=natural)
    assert (oa[0]["eflag"] == 0).all() and (ob_[0]["eflag"] == 0).all()
    assert rel_err(oa[1], ob_[1]) <= 1e-7
    Hp, Hi, Gp, Gi, Ap, Ai = pat
    with pytest.raises(fb.FbstabError):
        fb.FBstabSparse.analyze(nz, nl, nv, (Hp, Hi[::-1].copy(), Gp, Gi, Ap, Ai))
    with pytest.raises(fb.FbstabError):
        fb.FBstabSparse.analyze(nz, nl, nv, pat, perm=np.zeros(n, np.int32))


def _gpu_solve(fb, nz, nl, nv, pat, vals, B, opts=None, perm=None, x0=None):
    s = fb.FBstabSparse(nz, nl, nv, pat, max_batch=B, perm=perm)
    if opts is not None:
        s.update_options(opts)
    z, l, v = (np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)) if x0 is None else \
        [a.copy() for a in x0]
    out, y = s.solve_batch(vals, z, l, v)
    return s, out, z, l, v, y


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(DENSE_CASES))
def test_sparse_gpu_on_the_dense_solver_cases(fb, name):
    H, f, G, h, A, b, flag = dense_case(name)
    pat, (Hx, Gx, Ax) = sparse_from_dense(H, G, A)
    vals = {"Hx": Hx, "f": f, "Gx": Gx, "h": h, "Ax": Ax, "b": b}
    s, out, z, l, v, y = _gpu_solve(fb, f.size, h.size, b.size, pat, vals, 1,
                                    _opts(fb.FBstabSparse))
    assert fb.EXIT_FLAGS[int(out["eflag"][0])] == flag and out["status"][0] == 0
    if name == "FeasibleQP":
        np.testing.assert_allclose(z, [0, -5], atol=1e-8)
        np.testing.assert_allclose(v, [5, 0], atol=1e-8)
    if name == "FeasibleQPwithEQ":
        np.testing.assert_allclose(z, [0.25, 0.75], atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("team", ["1", "0"])
@pytest.mark.parametrize("shape", [(24, 4, 40, 300), (60, 10, 90, 70), (9, 0, 14, 33)])
def test_sparse_gpu_parity_with_the_oracle(fb, oracle, monkeypatch, shape, team):
    """Both device paths (warp per instance, sparse_team.cu; lane per instance,
    sparse_lane.cu): same exit flags; same trajectory (Newton, proximal and backtrack
    counts) and solutions within 1e-8 relative (north_star's tolerance) -- the oracle
    eliminates in the order the handle chose."""
    monkeypatch.setenv("FBSTAB_SPARSE_TEAM", team)
    nz, nl, nv, B = shape
    rng = np.random.default_rng(nz)
    pat, vals, _ = random_sparse_qp(rng, nz, nl, nv, count=B)
    s, out, z, l, v, y = _gpu_solve(fb, nz, nl, nv, pat, vals, B)
    assert s.path.startswith("sparse-team" if team == "1" else "sparse-lane"), s.path
    n, nnzK, nnzL, perm = s.analysis()
    assert n == nz + nl + nv and sorted(perm.tolist()) == list(range(n))
    oo, oz, ol, ov, oy = oracle.sparse_solve_batch(
        nz, nl, nv, pat, [vals[k] for k in ("Hx", "f", "Gx", "h", "Ax", "b")], perm=perm,
        nthreads=4)
    assert (out["eflag"] == oo["eflag"]).all() and (out["eflag"] == 0).all()
    assert (out["status"] == 0).all()
    same = (out["newton_iters"] == oo["newton_iters"]) & (out["prox_iters"] == oo["prox_iters"]) \
        & (out["ls_backtracks"] == oo["ls_backtracks"])
    assert same.mean() >= 0.97, same.mean()
    assert np.abs(out["newton_iters"] - oo["newton_iters"]).max() <= 3
    for i in range(B):
        tol = 1e-8 if same[i] else 1e-5
        assert rel_err(z[i * nz:(i + 1) * nz], oz[i * nz:(i + 1) * nz]) <= tol
        assert rel_err(v[i * nv:(i + 1) * nv], ov[i * nv:(i + 1) * nv]) <= tol
        assert rel_err(y[i * nv:(i + 1) * nv], oy[i * nv:(i + 1) * nv]) <= tol
        if nl:
            assert rel_err(l[i * nl:(i + 1) * nl], ol[i * nl:(i + 1) * nl]) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("team", ["1", "0"])
def test_sparse_gpu_wide_rows(fb, oracle, monkeypatch, team):
    """A nearly dense pattern: columns of K, rows of the LDL' schedule and columns of L
    with more than 32 entries (the strided loops of the warp-per-instance kernel)."""
    monkeypatch.setenv("FBSTAB_SPARSE_TEAM", team)
    nz, nl, nv, B = 44, 5, 60, 48
    rng = np.random.default_rng(44)
    pat, vals, _ = random_sparse_qp(rng, nz, nl, nv, band=nz, count=B)
    s, out, z, l, v, y = _gpu_solve(fb, nz, nl, nv, pat, vals, B)
    Lp, Li = s.factor_pattern()
    assert np.diff(Lp).max() > 32 and np.diff(pat[0]).max() > 32
    oo, oz, ol, ov, oy = oracle.sparse_solve_batch(
        nz, nl, nv, pat, [vals[k] for k in ("Hx", "f", "Gx", "h", "Ax", "b")],
        perm=s.analysis()[3], nthreads=4)
    assert (out["eflag"] == oo["eflag"]).all() and (out["eflag"] == 0).all()
    same = (out["newton_iters"] == oo["newton_iters"]) & (out["prox_iters"] == oo["prox_iters"])
    assert same.mean() >= 0.95
    for i in np.nonzero(same)[0]:
        assert rel_err(z[i * nz:(i + 1) * nz], oz[i * nz:(i + 1) * nz]) <= 1e-8
        assert rel_err(v[i * nv:(i + 1) * nv], ov[i * nv:(i + 1) * nv]) <= 1e-8


@pytest.mark.gpu
def test_sparse_gpu_batch_equals_single_solves_and_device_pointers(fb):
    import torch
    nz, nl, nv, B = 16, 3, 24, 37
    rng = np.random.default_rng(3)
    pat, vals, _ = random_sparse_qp(rng, nz, nl, nv, count=B)
    s, out, z, l, v, y = _gpu_solve(fb, nz, nl, nv, pat, vals, B)
    sizes = s.field_sizes
    one = fb.FBstabSparse(nz, nl, nv, pat, max_batch=1)
    for i in (0, 5, B - 1):
        vi = {k: a[i * sizes[k]:(i + 1) * sizes[k]].copy() for k, a in vals.items()}
        z1, l1, v1 = np.zeros(nz), np.zeros(nl), np.zeros(nv)
        o1, y1 = one.solve_batch(vi, z1, l1, v1)
        assert np.array_equal(z1, z[i * nz:(i + 1) * nz])
        assert np.array_equal(v1, v[i * nv:(i + 1) * nv])
        assert o1["newton_iters"][0] == out["newton_iters"][i]
    dev = torch.device("cuda:0")
    dv = {k: torch.from_numpy(a).to(dev) for k, a in vals.items()}
    zt, lt, vt = (torch.zeros(B * n_, dtype=torch.float64, device=dev) for n_ in (nz, nl, nv))
    s.solve_batch(dv, zt, lt, vt)
    torch.cuda.synchronize()
    assert np.array_equal(zt.cpu().numpy(), z) and np.array_equal(vt.cpu().numpy(), v)
    # warm start at the solution: no Newton step (impl:162-169)
    o2, _ = s.solve_batch(vals, z.copy(), l.copy(), v.copy())
    assert (o2["eflag"] == 0).all() and o2["newton_iters"].max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("team", ["1", "0"])
def test_sparse_gpu_mixed_exit_flags(fb, oracle, monkeypatch, team):
    """Feasible, primal-infeasible and unbounded instances in one batch keep their own
    flags (the cases of fbstab_dense_unit_tests.cc:121-256 share a pattern once the
    zero entries are stored)."""
    monkeypatch.setenv("FBSTAB_SPARSE_TEAM", team)
    names = ["DegenerateQP", "InfeasibleQP", "DegenerateQP", "InfeasibleQP"]
    cases = [dense_case(n) for n in names]
    keepH = np.ones((2, 2), bool)
    keepA = np.ones((5, 2), bool)
    Hp, Hi, _ = csc(cases[0][0], upper=True, keep=keepH)
    Ap, Ai, _ = csc(cases[0][4], keep=keepA)
    pat = (Hp, Hi, np.zeros(3, np.int32), np.zeros(0, np.int32), Ap, Ai)
    vals = {"Hx": np.concatenate([csc(c[0], upper=True, keep=keepH)[2] for c in cases]),
            "f": np.concatenate([c[1] for c in cases]), "Gx": np.zeros(0), "h": np.zeros(0),
            "Ax": np.concatenate([csc(c[4], keep=keepA)[2] for c in cases]),
            "b": np.concatenate([c[5] for c in cases])}
    s, out, z, l, v, y = _gpu_solve(fb, 2, 0, 5, pat, vals, 4, _opts(fb.FBstabSparse))
    assert [fb.EXIT_FLAGS[int(e)] for e in out["eflag"]] == [c[6] for c in cases]
    oo, *_ = oracle.sparse_solve_batch(2, 0, 5, pat, [vals[k] for k in
                                       ("Hx", "f", "Gx", "h", "Ax", "b")],
                                       perm=s.analysis()[3], opts=_opts(oracle))
    assert (oo["eflag"] == out["eflag"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("team", ["1", "0"])
@pytest.mark.parametrize("kind,N", [("double_integrator", 20), ("servo_motor", 25)])
def test_sparse_gpu_on_ocps_matches_the_mpc_solver(fb, monkeypatch, kind, N, team):
    """The reference's live MPC tests (fbstab_mpc_unit_tests.cc:62-104) as general sparse
    QPs: same flag and solution as FBstabMpc on the structured form."""
    monkeypatch.setenv("FBSTAB_SPARSE_TEAM", team)
    B = 40
    dims, d = fb.problems.ocp_batch(kind, N, count=B, config=3, rho=0.01)
    qps = [ocp_as_qp(dims, d, i) for i in range(B)]
    keepH = sum(np.abs(q[0]) for q in qps) > 0
    keepG = sum(np.abs(q[2]) for q in qps) > 0
    keepA = sum(np.abs(q[4]) for q in qps) > 0
    Hp, Hi, _ = csc(qps[0][0], upper=True, keep=keepH)
    Gp, Gi, _ = csc(qps[0][2], keep=keepG)
    Ap, Ai, _ = csc(qps[0][4], keep=keepA)
    vals = {"Hx": np.concatenate([csc(q[0], upper=True, keep=keepH)[2] for q in qps]),
            "f": np.concatenate([q[1] for q in qps]),
            "Gx": np.concatenate([csc(q[2], keep=keepG)[2] for q in qps]),
            "h": np.concatenate([q[3] for q in qps]),
            "Ax": np.concatenate([csc(q[4], keep=keepA)[2] for q in qps]),
            "b": np.concatenate([q[5] for q in qps])}
    nz, nl, nv = qps[0][1].size, qps[0][3].size, qps[0][5].size
    s, out, z, l, v, y = _gpu_solve(fb, nz, nl, nv, (Hp, Hi, Gp, Gi, Ap, Ai), vals, B)
    m = fb.FBstabMpc(*dims, max_batch=B)
    zm, lm, vm = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)
    om, ym = m.solve_batch(d, zm, lm, vm)
    # (the perturbed initial states make some instances infeasible: both solvers say so)
    assert (out["eflag"] == om["eflag"]).all() and (out["eflag"] == 0).sum() >= B // 3
    # two different linear solvers, each stopping at the absolute tolerance 1e-6 on its own
    # trajectory: they agree like off-trajectory instances do (tests/golden/
    # trajectory_floor.json: two builds of ONE solver differ by up to 3.6e-6 on the servo
    # family, 3.5e-5 in the sparse form), not to 1e-8
    tol = 1e-5
    for i in np.nonzero(out["eflag"] == 0)[0]:
        assert rel_err(z[i * nz:(i + 1) * nz], zm[i * nz:(i + 1) * nz]) <= tol
        assert rel_err(l[i * nl:(i + 1) * nl], lm[i * nl:(i + 1) * nl]) <= tol
        assert rel_err(v[i * nv:(i + 1) * nv], vm[i * nv:(i + 1) * nv]) <= tol
        assert out["residual"][i] <= 1e-6


@pytest.mark.gpu
def test_sparse_multi_device_solve_matches_one_handle(fb):
    """fbstab_sparse_multi_gpu_solve over every visible device (host arrays, contiguous
    instance ranges, the same elimination order on every device): the same bytes as one
    handle on one device.  (On the one-GPU test box this is a single shard; the N-GPU run
    is tests/multi_gpu_check.py.)"""
    nz, nl, nv, B = 24, 4, 40, 203
    rng = np.random.default_rng(7)
    pat, vals, _ = random_sparse_qp(rng, nz, nl, nv, count=B)
    s, out, z, l, v, y = _gpu_solve(fb, nz, nl, nv, pat, vals, B)
    z2, l2, v2 = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)
    devices = list(range(fb.capi.device_count()))
    out2, y2 = s.solve_batch_devices(vals, z2, l2, v2, devices)
    for a, b in ((z, z2), (l, l2), (v, v2), (y, y2)):
        assert a.tobytes() == b.tobytes()
    for f in ("eflag", "newton_iters", "prox_iters", "ls_backtracks", "status"):
        assert (out[f] == out2[f]).all(), f
    with pytest.raises(fb.FbstabError):
        s.solve_batch_devices(vals, z2, l2, v2, [0, 0])  # duplicate device

