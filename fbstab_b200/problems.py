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
