"""ctypes binding of the C-ABI in include/fbstab_b200.h.

This is the same binding a reference maintainer would write;
tests and bench.py call the engine through it.  There is no CPU fallback: if
the CUDA library is missing the import fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FBSTAB_B200_LIB") or os.path.join(_HERE, "libfbstab_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOGPU, ERR_ALLOC, ERR_NCCL = 0, 1, 2, 3, 4, 5
EXIT_FLAGS = {0: "SUCCESS", 1: "DIVERGENCE", 2: "MAXITERATIONS",
              3: "PRIMAL_INFEASIBLE", 4: "DUAL_INFEASIBLE",
              5: "PRIMAL_DUAL_INFEASIBLE"}
COMP_MARGIN, COMP_RESIDUAL, COMP_NEWTON, COMP_FEAS = 0, 1, 2, 3
OCP_DOUBLE_INTEGRATOR, OCP_SERVO_MOTOR, OCP_SPACECRAFT, OCP_COPOLYMERIZATION = 0, 1, 2, 3

# every symbol include/fbstab_b200.h declares
SYMBOLS = [
    "fbstab_default_options", "fbstab_reliable_options",
    "fbstab_validate_options", "fbstab_last_error", "fbstab_device_count",
    "fbstab_dense_batch_create", "fbstab_dense_batch_destroy",
    "fbstab_dense_batch_set_options", "fbstab_dense_batch_get_options",
    "fbstab_dense_batch_solve", "fbstab_dense_batch_last_launches",
    "fbstab_dense_batch_path", "fbstab_dense_batch_component",
    "fbstab_mpc_batch_create", "fbstab_mpc_batch_destroy",
    "fbstab_mpc_batch_set_options", "fbstab_mpc_batch_get_options",
    "fbstab_mpc_batch_solve", "fbstab_mpc_batch_last_launches",
    "fbstab_mpc_batch_path", "fbstab_mpc_batch_component",
    "fbstab_mpc_batch_solve_shared", "fbstab_mpc_batch_solve_lti",
    "fbstab_sparse_batch_create", "fbstab_sparse_batch_destroy",
    "fbstab_sparse_batch_set_options", "fbstab_sparse_batch_get_options",
    "fbstab_sparse_batch_solve", "fbstab_sparse_batch_last_launches",
    "fbstab_sparse_batch_path", "fbstab_sparse_batch_analysis",
    "fbstab_sparse_batch_factor_pattern", "fbstab_sparse_analyze",
    "fbstab_mpc_closed_loop_create", "fbstab_mpc_closed_loop_destroy",
    "fbstab_mpc_closed_loop_set_options", "fbstab_mpc_closed_loop_reset",
    "fbstab_mpc_closed_loop_step", "fbstab_mpc_closed_loop_run",
    "fbstab_mpc_closed_loop_path",
    "fbstab_multi_gpu_unique_id", "fbstab_multi_gpu_create", "fbstab_multi_gpu_destroy",
    "fbstab_multi_gpu_shard", "fbstab_multi_gpu_gather",
    "fbstab_dense_multi_gpu_create", "fbstab_dense_multi_gpu_destroy",
    "fbstab_dense_multi_gpu_set_options", "fbstab_dense_multi_gpu_solve",
    "fbstab_mpc_multi_gpu_create", "fbstab_mpc_multi_gpu_destroy",
    "fbstab_mpc_multi_gpu_set_options", "fbstab_mpc_multi_gpu_solve",
    "fbstab_sparse_multi_gpu_create", "fbstab_sparse_multi_gpu_destroy",
    "fbstab_sparse_multi_gpu_set_options", "fbstab_sparse_multi_gpu_solve",
    "fbstab_ocp_dims", "fbstab_ocp_generate", "fbstab_ocp_generate_batch",
    "fbstab_ocp_simulation",
    "fbstab_random_dense_qp", "fbstab_fp64_peak",
    "fbstab_fp64_peak_concurrent",
]


class Options(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "sigma0", "sigma_max", "sigma_min", "alpha", "beta", "eta", "delta",
        "gamma", "abs_tol", "rel_tol", "stall_tol", "infeas_tol",
        "inner_tol_max", "inner_tol_min")] + [(n, C.c_int32) for n in (
            "max_newton_iters", "max_prox_iters", "max_inner_iters",
            "max_linesearch_iters", "check_feasibility",
            "nonmonotone_linesearch", "display_level", "refine_steps",
            "regularize_retries")]


OUT_DTYPE = np.dtype([("eflag", "i4"), ("newton_iters", "i4"),
                      ("prox_iters", "i4"), ("status", "i4"),
                      ("residual", "f8"), ("initial_residual", "f8"),
                      ("solve_time", "f8"), ("ls_backtracks", "i4"),
                      ("residual_evals", "i4")])
assert OUT_DTYPE.itemsize == 48


class ComponentIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "z", "l", "v", "y", "zbar", "lbar", "vbar", "rz", "rl", "rv", "dz",
        "dl", "dv", "dy", "gamma", "mus", "norms", "status")] + [
            ("sigma", C.c_double), ("tol", C.c_double)]


class FbstabError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fbstab_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libfbstab_b200.so; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.fbstab_last_error.restype = C.c_char_p
        L.fbstab_dense_batch_path.restype = C.c_char_p
        L.fbstab_mpc_batch_path.restype = C.c_char_p
        L.fbstab_dense_batch_path.argtypes = [C.c_void_p]
        L.fbstab_mpc_batch_path.argtypes = [C.c_void_p]
        L.fbstab_dense_batch_create.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_void_p)]
        L.fbstab_mpc_batch_create.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_void_p)]
        for n in ("fbstab_dense_batch_destroy", "fbstab_mpc_batch_destroy",
                  "fbstab_dense_batch_last_launches",
                  "fbstab_mpc_batch_last_launches"):
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("fbstab_dense_batch_set_options", "fbstab_mpc_batch_set_options",
                  "fbstab_dense_batch_get_options", "fbstab_mpc_batch_get_options"):
            getattr(L, n).argtypes = [C.c_void_p, C.POINTER(Options)]
        L.fbstab_dense_batch_solve.argtypes = (
            [C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_void_p, C.c_void_p])
        L.fbstab_mpc_batch_solve.argtypes = (
            [C.c_void_p, C.c_int] + [C.c_void_p] * 16 + [C.c_void_p, C.c_void_p])
        L.fbstab_dense_batch_component.argtypes = (
            [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6 +
            [C.POINTER(ComponentIO), C.c_void_p])
        L.fbstab_mpc_batch_component.argtypes = (
            [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 12 +
            [C.POINTER(ComponentIO), C.c_void_p])
        ip = C.POINTER(C.c_int)
        L.fbstab_sparse_batch_path.restype = C.c_char_p
        L.fbstab_sparse_batch_path.argtypes = [C.c_void_p]
        L.fbstab_sparse_batch_create.argtypes = (
            [C.c_int] * 3 + [C.c_void_p] * 7 + [C.c_int, C.c_int, C.POINTER(C.c_void_p)])
        L.fbstab_sparse_batch_destroy.argtypes = [C.c_void_p]
        L.fbstab_sparse_batch_last_launches.argtypes = [C.c_void_p]
        L.fbstab_sparse_batch_set_options.argtypes = [C.c_void_p, C.POINTER(Options)]
        L.fbstab_sparse_batch_get_options.argtypes = [C.c_void_p, C.POINTER(Options)]
        L.fbstab_sparse_batch_solve.argtypes = (
            [C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_void_p, C.c_void_p])
        L.fbstab_sparse_batch_analysis.argtypes = [C.c_void_p, ip, ip, ip, C.c_void_p]
        L.fbstab_sparse_batch_factor_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fbstab_sparse_analyze.argtypes = ([C.c_int] * 3 + [C.c_void_p] * 7 + [ip, ip, ip, C.c_void_p])
        L.fbstab_ocp_dims.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 3
        L.fbstab_ocp_generate.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 12
        L.fbstab_ocp_generate_batch.argtypes = (
            [C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, C.c_double] +
            [C.c_void_p] * 12)
        L.fbstab_random_dense_qp.argtypes = (
            [C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] +
            [C.c_void_p] * 6 + [C.c_int])
        L.fbstab_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise FbstabError(rc, lib().fbstab_last_error().decode())


def ptr(a):
    """Address of a numpy array, a torch tensor, an int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise RuntimeError("arrays must be contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (host or cuda)
        if not a.is_contiguous():
            raise RuntimeError("tensors must be contiguous")
        return a.data_ptr()
    raise TypeError(type(a))


def default_options(**kw):
    o = Options()
    lib().fbstab_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def reliable_options(**kw):
    o = Options()
    lib().fbstab_reliable_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def validate_options(o):
    return lib().fbstab_validate_options(C.byref(o))


def device_count():
    return lib().fbstab_device_count()


def fp64_peak(device=0):
    """(DFMA, DMMA) measured FP64 peaks in TFLOP/s."""
    a, b = C.c_double(), C.c_double()
    check(lib().fbstab_fp64_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def fp64_peak_concurrent(device=0):
    """Combined TFLOP/s with DFMA and DMMA warps running at the same time."""
    a = C.c_double()
    L = lib()
    L.fbstab_fp64_peak_concurrent.argtypes = [C.c_int, C.POINTER(C.c_double)]
    check(L.fbstab_fp64_peak_concurrent(device, C.byref(a)))
    return a.value
