"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by
fbstab_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
# Second build of the SAME restatement with FMA contraction allowed
# (-ffp-contract=fast -mfma).  Not a second oracle: it exists to measure how far
# two correctly rounded evaluations of the reference's arithmetic drift apart in
# iteration counts (tests/test_oracle_fma_floor.py) -- the floor under any
# "same trajectory" requirement on the GPU path.
_FMA_LIB_PATH = os.path.join(_HERE, "liboracle_fma.so")

EXIT_FLAGS = {0: "SUCCESS", 1: "DIVERGENCE", 2: "MAXITERATIONS",
              3: "PRIMAL_INFEASIBLE", 4: "DUAL_INFEASIBLE",
              5: "PRIMAL_DUAL_INFEASIBLE"}
TRAJ_STRIDE = 8


class Options(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "sigma0", "sigma_max", "sigma_min", "alpha", "beta", "eta", "delta",
        "gamma", "abs_tol", "rel_tol", "stall_tol", "infeas_tol",
        "inner_tol_max", "inner_tol_min")] + [(n, C.c_int) for n in (
            "max_newton_iters", "max_prox_iters", "max_inner_iters",
            "max_linesearch_iters", "check_feasibility",
            "nonmonotone_linesearch", "display_level", "refine_steps",
            "regularize_retries")]


class Out(C.Structure):
    _fields_ = [("eflag", C.c_int), ("newton_iters", C.c_int),
                ("prox_iters", C.c_int), ("status", C.c_int),
                ("residual", C.c_double), ("initial_residual", C.c_double),
                ("solve_time", C.c_double), ("ls_backtracks", C.c_int),
                ("residual_evals", C.c_int)]


OUT_DTYPE = np.dtype([("eflag", "i4"), ("newton_iters", "i4"),
                      ("prox_iters", "i4"), ("status", "i4"),
                      ("residual", "f8"), ("initial_residual", "f8"),
                      ("solve_time", "f8"), ("ls_backtracks", "i4"),
                      ("residual_evals", "i4")])
assert OUT_DTYPE.itemsize == C.sizeof(Out) == 48


def build(force=False):
    """Compile liboracle.so (and the FMA-contracted build) with the committed Makefile."""
    newest = max(os.path.getmtime(os.path.join(_HERE, f))
                 for f in ("fbstab_oracle.cpp", "fbstab_oracle.h", "Makefile"))
    if force or any(not os.path.exists(q) or os.path.getmtime(q) < newest
                    for q in (_LIB_PATH, _FMA_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None
_fma_lib = None
_dp = C.POINTER(C.c_double)


def fma_lib():
    """The FMA-contracted build (batch solvers only; see _FMA_LIB_PATH)."""
    global _fma_lib
    if _fma_lib is None:
        if not os.path.exists(_FMA_LIB_PATH):
            build()
        L = C.CDLL(_FMA_LIB_PATH)
        L.oracle_dense_solve_batch.argtypes = (
            [C.c_int] * 4 + [_dp] * 10 + [C.POINTER(Options), C.c_void_p,
                                          C.c_int, C.c_int])
        L.oracle_mpc_solve_batch.argtypes = (
            [C.c_int] * 5 + [_dp] * 16 + [C.POINTER(Options), C.c_void_p,
                                          C.c_int])
        _ip = C.POINTER(C.c_int)
        L.oracle_sparse_solve_batch.argtypes = (
            [C.c_int] * 4 + [_ip, _ip, _dp, _dp] * 3 + [_ip] + [_dp] * 4 +
            [C.POINTER(Options), C.c_void_p, C.c_int])
        _fma_lib = L
    return _fma_lib


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_dense_create.restype = C.c_void_p
        L.oracle_dense_create.argtypes = [C.c_int] * 3 + [_dp] * 6
        L.oracle_mpc_create.restype = C.c_void_p
        L.oracle_mpc_create.argtypes = [C.c_int] * 4 + [_dp] * 12
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_forcing_norm.restype = C.c_double
        L.oracle_forcing_norm.argtypes = [C.c_void_p]
        L.oracle_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.oracle_gemv.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double,
                                  C.c_double, _dp]
        L.oracle_axpy.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp]
        L.oracle_margin.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_variable_axpy.argtypes = [C.c_void_p, C.c_double] + [_dp] * 8
        L.oracle_residual.argtypes = [C.c_void_p, C.c_int, C.c_double,
                                      C.c_double] + [_dp] * 11
        L.oracle_linear_solve.argtypes = [C.c_void_p, C.c_int, C.c_double,
                                          C.c_double] + [_dp] * 16
        L.oracle_feasibility.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double]
        L.oracle_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(Options),
                                   _dp, _dp, _dp, _dp, C.POINTER(Out), _dp,
                                   C.c_int, C.POINTER(C.c_int)]
        L.oracle_dense_solve_batch.argtypes = (
            [C.c_int] * 4 + [_dp] * 10 + [C.POINTER(Options), C.c_void_p,
                                          C.c_int, C.c_int])
        L.oracle_mpc_solve_batch.argtypes = (
            [C.c_int] * 5 + [_dp] * 16 + [C.POINTER(Options), C.c_void_p,
                                          C.c_int])
        _ip = C.POINTER(C.c_int)
        L.oracle_sparse_create.restype = C.c_void_p
        L.oracle_sparse_create.argtypes = [C.c_int] * 3 + [_ip, _ip, _dp, _dp] * 3 + [_ip]
        L.oracle_qdldl_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp]
        L.oracle_sparse_solve_batch.argtypes = (
            [C.c_int] * 4 + [_ip, _ip, _dp, _dp] * 3 + [_ip] + [_dp] * 4 +
            [C.POINTER(Options), C.c_void_p, C.c_int])
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_dp)


def _ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f(a):
    """float64 contiguous 1-D copy, column-major flattening for 2-D input."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        a = a.T  # column-major
    return np.ascontiguousarray(a).reshape(-1).copy()


def default_options(**kw):
    o = Options()
    lib().oracle_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def reliable_options(**kw):
    o = Options()
    lib().oracle_reliable_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def validate_options(o):
    return lib().oracle_validate_options(C.byref(o))


class Problem:
    """Owns flat copies of the data and an oracle_problem handle."""

    def __init__(self, handle, keep, nz, nl, nv):
        self._h = handle
        self._keep = keep
        self.nz, self.nl, self.nv = nz, nl, nv

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    @staticmethod
    def dense(H, f, G, h, A, b):
        """Matrices as 2-D numpy arrays (any order) or flat column-major."""
        f, h, b = _f(f), _f(h), _f(b)
        nz, nl, nv = f.size, h.size, b.size
        H, G, A = _f(H), _f(G), _f(A)
        assert H.size == nz * nz and G.size == nl * nz and A.size == nv * nz
        hd = lib().oracle_dense_create(nz, nl, nv, _p(H), _p(f), _p(G), _p(h),
                                       _p(A), _p(b))
        if not hd:
            raise ValueError("oracle_dense_create rejected the sizes")
        return Problem(hd, (H, f, G, h, A, b), nz, nl, nv)

    @staticmethod
    def mpc(N, nx, nu, nc, Q, R, S, q, r, A, B, c, E, L, d, x0):
        """Sequences as flat arrays in the wire format."""
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
                for a in (Q, R, S, q, r, A, B, c, E, L, d, x0)]
        hd = lib().oracle_mpc_create(N, nx, nu, nc, *[_p(a) for a in arrs])
        if not hd:
            raise ValueError("oracle_mpc_create rejected the sizes")
        return Problem(hd, arrs, (N + 1) * (nx + nu), (N + 1) * nx,
                       (N + 1) * nc)

    @staticmethod
    def sparse(nz, nl, nv, Hp, Hi, Hx, f, Gp, Gi, Gx, h, Ap, Ai, Ax, b, perm=None):
        """H by its upper triangle, G and A: compressed-column int32 patterns + values."""
        ia = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
        fa = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
        ints = [ia(a) for a in (Hp, Hi, Gp, Gi, Ap, Ai)]
        vals = [fa(a) for a in (Hx, f, Gx, h, Ax, b)]
        pm = ia(perm) if perm is not None else None
        hd = lib().oracle_sparse_create(
            nz, nl, nv, _ip(ints[0]), _ip(ints[1]), _p(vals[0]), _p(vals[1]), _ip(ints[2]),
            _ip(ints[3]), _p(vals[2]), _p(vals[3]), _ip(ints[4]), _ip(ints[5]), _p(vals[4]),
            _p(vals[5]), _ip(pm))
        if not hd:
            raise ValueError("oracle_sparse_create rejected the sizes")
        return Problem(hd, (ints, vals, pm), nz, nl, nv)

    # -- data ops ---------------------------------------------------------
    def forcing_norm(self):
        return lib().oracle_forcing_norm(self._h)

    def gemv(self, op, x, a, b, y):
        ops = {"H": 0, "A": 1, "AT": 2, "G": 3, "GT": 4}
        x = _f(x)
        y = _f(y)
        lib().oracle_gemv(self._h, ops[op], _p(x), a, b, _p(y))
        return y

    def axpy(self, which, a, y):
        y = _f(y)
        lib().oracle_axpy(self._h, {"f": 0, "h": 1, "b": 2}[which], a, _p(y))
        return y

    def margin(self, z):
        z = _f(z)
        y = np.zeros(self.nv)
        lib().oracle_margin(self._h, _p(z), _p(y))
        return y

    def variable_axpy(self, a, dx, x):
        dz, dl, dv, dy = [_f(t) for t in dx]
        z, l, v, y = [_f(t) for t in x]
        lib().oracle_variable_axpy(self._h, a, _p(dz), _p(dl), _p(dv), _p(dy),
                                   _p(z), _p(l), _p(v), _p(y))
        return z, l, v, y

    def residual(self, kind, x, xbar=None, sigma=0.0, alpha=0.95):
        """kind: 'inner' | 'natural' | 'penalized'.  x=(z,l,v[,y])."""
        z, l, v = _f(x[0]), _f(x[1]), _f(x[2])
        y = _f(x[3]) if len(x) > 3 and x[3] is not None else self.margin(z)
        if xbar is None:
            xbar = (z, l, v)
        zb, lb, vb = _f(xbar[0]), _f(xbar[1]), _f(xbar[2])
        rz, rl, rv = np.zeros(self.nz), np.zeros(self.nl), np.zeros(self.nv)
        norms = np.zeros(3)
        k = {"inner": 0, "natural": 1, "penalized": 2}[kind]
        lib().oracle_residual(self._h, k, alpha, sigma, _p(z), _p(l), _p(v),
                              _p(y), _p(zb), _p(lb), _p(vb), _p(rz), _p(rl),
                              _p(rv), _p(norms))
        return rz, rl, rv, norms

    def linear_solve(self, x, xbar, sigma, r, alpha=0.95, variant=0):
        z, l, v = _f(x[0]), _f(x[1]), _f(x[2])
        y = _f(x[3]) if len(x) > 3 and x[3] is not None else self.margin(z)
        zb, lb, vb = _f(xbar[0]), _f(xbar[1]), _f(xbar[2])
        rz, rl, rv = _f(r[0]), _f(r[1]), _f(r[2])
        dz, dl = np.zeros(self.nz), np.zeros(self.nl)
        dv, dy = np.zeros(self.nv), np.zeros(self.nv)
        gamma, mus = np.zeros(self.nv), np.zeros(self.nv)
        rc = lib().oracle_linear_solve(
            self._h, variant, alpha, sigma, _p(z), _p(l), _p(v), _p(y), _p(zb),
            _p(lb), _p(vb), _p(rz), _p(rl), _p(rv), _p(dz), _p(dl), _p(dv),
            _p(dy), _p(gamma), _p(mus))
        return rc, (dz, dl, dv, dy), gamma, mus

    def feasibility(self, dz, dl, dv, tol):
        dz, dl, dv = _f(dz), _f(dl), _f(dv)
        return lib().oracle_feasibility(self._h, _p(dz), _p(dl), _p(dv), tol)

    def solve(self, opts=None, x0=None, variant=0, traj_cap=0):
        """Returns (out dict, (z,l,v,y), traj ndarray or None)."""
        if opts is None:
            opts = default_options()
        if x0 is None:
            z, l, v = np.zeros(self.nz), np.zeros(self.nl), np.zeros(self.nv)
        else:
            z, l, v = _f(x0[0]), _f(x0[1]), _f(x0[2])
        y = np.zeros(self.nv)
        out = Out()
        traj = np.zeros((traj_cap, TRAJ_STRIDE)) if traj_cap else None
        tl = C.c_int(0)
        lib().oracle_solve(self._h, variant, C.byref(opts), _p(z), _p(l),
                           _p(v), _p(y), C.byref(out),
                           _p(traj) if traj is not None else None, traj_cap,
                           C.byref(tl))
        o = {n: getattr(out, n) for n, _ in Out._fields_}
        o["flag"] = EXIT_FLAGS[out.eflag]
        if traj is not None:
            traj = traj[:min(tl.value, traj_cap)]
        return o, (z, l, v, y), traj


def dense_solve_batch(nz, nl, nv, H, f, G, h, A, b, opts=None, x0=None,
                      variant=0, nthreads=1, fma=False):
    """Instance-major flat arrays.  Returns (out structured array, z,l,v,y)."""
    if opts is None:
        opts = default_options()
    batch = f.size // nz
    if x0 is None:
        z, l, v = (np.zeros(batch * nz), np.zeros(batch * nl),
                   np.zeros(batch * nv))
    else:
        z, l, v = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1).copy()
                   for t in x0]
    y = np.zeros(batch * nv)
    out = np.zeros(batch, dtype=OUT_DTYPE)
    (fma_lib() if fma else lib()).oracle_dense_solve_batch(
        nz, nl, nv, batch, _p(H), _p(f), _p(G), _p(h), _p(A), _p(b), _p(z),
        _p(l), _p(v), _p(y), C.byref(opts), out.ctypes.data, variant, nthreads)
    return out, z, l, v, y


def mpc_solve_batch(N, nx, nu, nc, seqs, opts=None, x0=None, nthreads=1, fma=False):
    """seqs = (Q,R,S,q,r,A,B,c,E,L,d,x0) instance-major flat arrays."""
    if opts is None:
        opts = default_options()
    batch = seqs[11].size // nx
    nz, nl, nv = (N + 1) * (nx + nu), (N + 1) * nx, (N + 1) * nc
    if x0 is None:
        z, l, v = (np.zeros(batch * nz), np.zeros(batch * nl),
                   np.zeros(batch * nv))
    else:
        z, l, v = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1).copy()
                   for t in x0]
    y = np.zeros(batch * nv)
    out = np.zeros(batch, dtype=OUT_DTYPE)
    (fma_lib() if fma else lib()).oracle_mpc_solve_batch(
        N, nx, nu, nc, batch, *[_p(a) for a in seqs], _p(z), _p(l), _p(v),
        _p(y), C.byref(opts), out.ctypes.data, nthreads)
    return out, z, l, v, y


# ---- the reference's own sources (oracle/_ref, `make -C oracle _ref`) ---------------------
# libfbstab_ref.so = the reference's unmodified algorithm sources compiled against
# oracle/eigen_shim (Eigen itself is not in this image); built only where /root/reference
# exists, shipped to the GPU box as a prebuilt file.  Used by tests to pin the oracle's
# TRAJECTORY to the reference's own code, never by the product.
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libfbstab_ref.so")
_ref_lib = None


def build_ref(reference="/root/reference"):
    """Builds oracle/_ref/libfbstab_ref.so when the reference tree is present; returns the
    path, or None when there is neither a tree nor a prebuilt library."""
    if os.path.isdir(os.path.join(reference, "fbstab", "components")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref", "REF=" + reference])
    return _REF_LIB_PATH if os.path.exists(_REF_LIB_PATH) else None


_REF_TESTS_PATH = os.path.join(_HERE, "_ref", "ref_unit_tests")


def build_ref_tests(reference="/root/reference"):
    """Builds oracle/_ref/ref_unit_tests -- the reference's own live unit tests
    (fbstab/test/*_unit_tests.cc), unmodified, on its own code, against eigen_shim/ and
    gtest_shim/ -- where the reference tree is present; returns the path of the binary, or
    None when there is neither a tree nor a prebuilt binary."""
    if os.path.isdir(os.path.join(reference, "fbstab", "test")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref_tests", "REF=" + reference])
    return _REF_TESTS_PATH if os.path.exists(_REF_TESTS_PATH) else None


def ref_lib():
    """The reference-sources library, or None when it is not available."""
    global _ref_lib
    if _ref_lib is None:
        if not os.path.exists(_REF_LIB_PATH) and build_ref() is None:
            return None
        L = C.CDLL(_REF_LIB_PATH)
        L.ref_dense_solve_batch.argtypes = (
            [C.c_int] * 4 + [_dp] * 10 + [C.POINTER(Options), C.c_void_p, C.c_int])
        L.ref_mpc_solve_batch.argtypes = (
            [C.c_int] * 5 + [_dp] * 16 + [C.POINTER(Options), C.c_void_p, C.c_int])
        _ref_lib = L
    return _ref_lib


def ref_ocp_generate(kind, N):
    """The reference's own OcpGenerator (fbstab/test/ocp_generator.cc) at horizon N:
    kind 0 DoubleIntegrator, 1 ServoMotor, 2 SpacecraftRelativeMotion,
    3 CopolymerizationReactor.  Returns ((N, nx, nu, nc), {field: flat array}) in the wire
    format of fbstab_b200.problems.ocp_batch."""
    L = ref_lib()
    ip = C.POINTER(C.c_int)
    L.ref_ocp_generate.argtypes = [C.c_int, C.c_int, ip] + [_dp] * 12
    sizes = (C.c_int * 3)()
    if L.ref_ocp_generate(kind, N, sizes, *([None] * 12)) != 0:
        raise RuntimeError("ref_ocp_generate failed")
    nx, nu, nc = sizes[0], sizes[1], sizes[2]
    shp = {"Q": (N + 1) * nx * nx, "R": (N + 1) * nu * nu, "S": (N + 1) * nu * nx,
           "q": (N + 1) * nx, "r": (N + 1) * nu, "A": N * nx * nx, "B": N * nx * nu,
           "c": N * nx, "E": (N + 1) * nc * nx, "L": (N + 1) * nc * nu, "d": (N + 1) * nc,
           "x0": nx}
    d = {k: np.zeros(n) for k, n in shp.items()}
    if L.ref_ocp_generate(kind, N, None, *[_p(d[k]) for k in shp]) != 0:
        raise RuntimeError("ref_ocp_generate failed")
    return (N, nx, nu, nc), d


def ref_dense_solve_batch(nz, nl, nv, H, f, G, h, A, b, opts=None, x0=None, nthreads=1):
    """FBstabDense::Solve of the reference's own sources on every instance of the batch;
    same arguments and results as dense_solve_batch (ls_backtracks / residual_evals = -1:
    the reference does not count them)."""
    if opts is None:
        opts = default_options()
    batch = f.size // nz
    if x0 is None:
        z, l, v = np.zeros(batch * nz), np.zeros(batch * nl), np.zeros(batch * nv)
    else:
        z, l, v = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1).copy() for t in x0]
    y = np.zeros(batch * nv)
    out = np.zeros(batch, dtype=OUT_DTYPE)
    ref_lib().ref_dense_solve_batch(
        nz, nl, nv, batch, _p(H), _p(f), _p(G), _p(h), _p(A), _p(b), _p(z), _p(l), _p(v),
        _p(y), C.byref(opts), out.ctypes.data, nthreads)
    return out, z, l, v, y


def ref_mpc_solve_batch(N, nx, nu, nc, seqs, opts=None, x0=None, nthreads=1):
    """FBstabMpc::Solve of the reference's own sources; arguments as mpc_solve_batch."""
    if opts is None:
        opts = default_options()
    nz, nl, nv = (N + 1) * (nx + nu), (N + 1) * nx, (N + 1) * nc
    batch = seqs[-1].size // nx
    if x0 is None:
        z, l, v = np.zeros(batch * nz), np.zeros(batch * nl), np.zeros(batch * nv)
    else:
        z, l, v = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1).copy() for t in x0]
    y = np.zeros(batch * nv)
    out = np.zeros(batch, dtype=OUT_DTYPE)
    ref_lib().ref_mpc_solve_batch(
        N, nx, nu, nc, batch, *[_p(a) for a in seqs], _p(z), _p(l), _p(v), _p(y),
        C.byref(opts), out.ctypes.data, nthreads)
    return out, z, l, v, y


def qdldl_solve(n, Ap, Ai, Ax, b):
    """QdldlWrapper: factor the upper-triangular CSC matrix, solve A x = b."""
    Ap, Ai = [np.ascontiguousarray(np.asarray(a, dtype=np.int32)) for a in (Ap, Ai)]
    Ax = np.ascontiguousarray(np.asarray(Ax, dtype=np.float64))
    x = np.ascontiguousarray(np.asarray(b, dtype=np.float64)).copy()
    rc = lib().oracle_qdldl_solve(n, _ip(Ap), _ip(Ai), _p(Ax), _p(x))
    return rc, x


def sparse_solve_batch(nz, nl, nv, pattern, vals, perm=None, opts=None, x0=None, nthreads=1,
                       fma=False):
    """pattern = (Hp,Hi,Gp,Gi,Ap,Ai) int32; vals = (Hx,f,Gx,h,Ax,b) instance-major."""
    if opts is None:
        opts = default_options()
    Hp, Hi, Gp, Gi, Ap, Ai = [np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
                              for a in pattern]
    Hx, f, Gx, h, Ax, b = vals
    batch = f.size // nz
    if x0 is None:
        z, l, v = np.zeros(batch * nz), np.zeros(batch * nl), np.zeros(batch * nv)
    else:
        z, l, v = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1).copy() for t in x0]
    y = np.zeros(batch * nv)
    out = np.zeros(batch, dtype=OUT_DTYPE)
    pm = np.ascontiguousarray(np.asarray(perm, dtype=np.int32)) if perm is not None else None
    (fma_lib() if fma else lib()).oracle_sparse_solve_batch(
        nz, nl, nv, batch, _ip(Hp), _ip(Hi), _p(Hx), _p(f), _ip(Gp), _ip(Gi), _p(Gx), _p(h),
        _ip(Ap), _ip(Ai), _p(Ax), _p(b), _ip(pm), _p(z), _p(l), _p(v), _p(y), C.byref(opts),
        out.ctypes.data, nthreads)
    return out, z, l, v, y
