"""Synthetic problems: the reference's four OCP fixtures and the seeded random
dense QP family (host-side generators in csrc/problems.cpp).

Mirrors reference fbstab/test/ocp_generator.h (OcpGenerator) for the OCPs.
"""
import ctypes as C
import os

import numpy as np

from . import capi

# The generators are plain host code (csrc/problems.cpp).  They are exported by the
# engine library for the C++ facade, and ALSO built into their own small library:
# whoever only needs problem data -- bench.py's CPU reference arm, the oracle-only
# tests -- never maps the CUDA engine.
_GEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfbstab_problems.so")
_gen = None


def _lib():
    global _gen
    if _gen is None:
        if not os.path.exists(_GEN_PATH):
            raise ImportError(f"{_GEN_PATH} is missing: build it with "
                              "`python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_GEN_PATH)
        L.fbstab_ocp_dims.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 3
        L.fbstab_ocp_generate.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 12
        L.fbstab_ocp_generate_batch.argtypes = (
            [C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, C.c_double] + [C.c_void_p] * 12)
        L.fbstab_random_dense_qp.argtypes = (
            [C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] +
            [C.c_void_p] * 6 + [C.c_int])
        _gen = L
    return _gen


def _check(rc):
    if rc != capi.OK:
        raise capi.FbstabError(rc, "problem generator rejected its arguments")


OCP_KINDS = {"double_integrator": capi.OCP_DOUBLE_INTEGRATOR,
             "servo_motor": capi.OCP_SERVO_MOTOR,
             "spacecraft": capi.OCP_SPACECRAFT,
             "copolymerization": capi.OCP_COPOLYMERIZATION}
MPC_FIELDS = ("Q", "R", "S", "q", "r", "A", "B", "c", "E", "L", "d", "x0")
DENSE_FIELDS = ("H", "f", "G", "h", "A", "b")


def ocp_dims(kind):
    nx, nu, nc = C.c_int(), C.c_int(), C.c_int()
    _check(_lib().fbstab_ocp_dims(OCP_KINDS[kind], C.byref(nx), C.byref(nu), C.byref(nc)))
    return nx.value, nu.value, nc.value


def mpc_field_sizes(N, nx, nu, nc):
    """Doubles per instance of each wire-format sequence."""
    K = N + 1
    return {"Q": K * nx * nx, "R": K * nu * nu, "S": K * nu * nx, "q": K * nx,
            "r": K * nu, "A": N * nx * nx, "B": N * nx * nu, "c": N * nx,
            "E": K * nc * nx, "L": K * nc * nu, "d": K * nc, "x0": nx}


def ocp_batch(kind, N, count=1, config=0, first=0, rho=0.0, alloc=None):
    """`count` instances of an OCP in the wire format, instance-major.

    Returns (dims, dict of flat float64 arrays).  alloc(n) may return a
    pinned/other float64 buffer of n elements (numpy array).
    """
    nx, nu, nc = ocp_dims(kind)
    sizes = mpc_field_sizes(N, nx, nu, nc)
    mk = alloc or (lambda n: np.empty(n, dtype=np.float64))
    arrs = {k: mk(count * sizes[k]) for k in MPC_FIELDS}
    _check(_lib().fbstab_ocp_generate_batch(
        OCP_KINDS[kind], N, config, first, count, rho,
        *[capi.ptr(arrs[k]) for k in MPC_FIELDS]))
    return (N, nx, nu, nc), arrs


def random_dense_qp(nz, nl, nv, count=1, config=0, first=0, kind=0,
                    nthreads=8, alloc=None):
    """Seeded random dense QPs, instance-major flat arrays (column-major
    matrices).  kind 0 feasible, 1 primal infeasible, 2 unbounded."""
    mk = alloc or (lambda n: np.empty(n, dtype=np.float64))
    sizes = {"H": nz * nz, "f": nz, "G": nl * nz, "h": nl, "A": nv * nz,
             "b": nv}
    arrs = {k: mk(count * sizes[k]) for k in DENSE_FIELDS}
    _check(_lib().fbstab_random_dense_qp(
        config, first, count, nz, nl, nv, kind,
        *[capi.ptr(arrs[k]) for k in DENSE_FIELDS], nthreads))
    return arrs


SPARSE_FIELDS = ("Hx", "f", "Gx", "h", "Ax", "b")


def ocp_as_sparse_qp(dims, d, count, alloc=None):
    """An OCP batch in the wire format as GENERAL sparse QPs on one pattern: the sign and
    ordering conventions of the reference's MpcData (mpc_data.cc:17-289: z = [x0;u0;..;
    xN;uN], H = blkdiag([Q S';S R]), G row-block 0 = [-I 0], row-block i = [A B] at stage
    i-1 and -I at x_i, h = -[x0;c], A_qp = blkdiag([E L]), b = -d).

    The pattern holds the entries that are non-zero in instance 0 (plus the -I blocks), so
    it is valid for batches that share their stage matrices -- as the generator's batches
    do: the instances differ in x0 only -- and the values of every instance are gathered
    on it.  Returns ((nz, nl, nv), (Hp, Hi, Gp, Gi, Ap, Ai), dict of instance-major values).
    """
    N, nx, nu, nc = dims
    K, ns = N + 1, nx + nu
    sizes = mpc_field_sizes(N, nx, nu, nc)
    nz, nl, nv = K * ns, K * nx, K * nc
    H, G, A = [], [], []  # entries (row, col, field, offset, scale): value = scale * field[offset]
    for i in range(K):
        o = i * ns
        for c in range(nx):
            for r in range(nx):
                if r <= c:
                    H.append((o + r, o + c, "Q", i * nx * nx + r + c * nx, 1.0))
        for c in range(nu):
            for r in range(nx):  # S' above the diagonal block of R: entry (x_r, u_c) = S(c, r)
                H.append((o + r, o + nx + c, "S", i * nu * nx + c + r * nu, 1.0))
            for r in range(nu):
                if r <= c:
                    H.append((o + nx + r, o + nx + c, "R", i * nu * nu + r + c * nu, 1.0))
        for r in range(nc):
            for c in range(nx):
                A.append((i * nc + r, o + c, "E", i * nc * nx + r + c * nc, 1.0))
            for c in range(nu):
                A.append((i * nc + r, o + nx + c, "L", i * nc * nu + r + c * nc, 1.0))
        for r in range(nx):
            G.append((i * nx + r, o + r, None, 0, -1.0))
        if i > 0:
            p = (i - 1) * ns
            for r in range(nx):
                for c in range(nx):
                    G.append((i * nx + r, p + c, "A", (i - 1) * nx * nx + r + c * nx, 1.0))
                for c in range(nu):
                    G.append((i * nx + r, p + nx + c, "B", (i - 1) * nx * nu + r + c * nx, 1.0))

    def compress(entries, cols):
        keep = [e for e in entries if e[2] is None or d[e[2]][e[3]] != 0.0]
        keep.sort(key=lambda e: (e[1], e[0]))
        ptr = np.zeros(cols + 1, dtype=np.int32)
        for e in keep:
            ptr[e[1] + 1] += 1
        ptr = np.cumsum(ptr).astype(np.int32)
        idx = np.array([e[0] for e in keep], dtype=np.int32)
        return ptr, idx, keep

    def gather(keep, out):
        out = out.reshape(count, len(keep))
        for field in set(e[2] for e in keep):
            cols = [k for k, e in enumerate(keep) if e[2] == field]
            if field is None:
                out[:, cols] = np.array([keep[k][4] for k in cols])
            else:
                src = d[field].reshape(count, sizes[field])
                out[:, cols] = src[:, [keep[k][3] for k in cols]] * np.array(
                    [keep[k][4] for k in cols])

    Hp, Hi, Hk = compress(H, nz)
    Gp, Gi, Gk = compress(G, nz)
    Ap, Ai, Ak = compress(A, nz)
    mk = alloc or (lambda n: np.empty(n, dtype=np.float64))
    vals = {"Hx": mk(count * len(Hk)), "f": mk(count * nz), "Gx": mk(count * len(Gk)),
            "h": mk(count * nl), "Ax": mk(count * len(Ak)), "b": mk(count * nv)}
    gather(Hk, vals["Hx"])
    gather(Gk, vals["Gx"])
    gather(Ak, vals["Ax"])
    f = vals["f"].reshape(count, K, ns)
    f[:, :, :nx] = d["q"].reshape(count, K, nx)
    f[:, :, nx:] = d["r"].reshape(count, K, nu)
    h = vals["h"].reshape(count, K, nx)
    h[:, 0, :] = -d["x0"].reshape(count, nx)
    h[:, 1:, :] = -d["c"].reshape(count, N, nx)
    vals["b"][:] = -d["d"]
    return (nz, nl, nv), (Hp, Hi, Gp, Gi, Ap, Ai), vals
