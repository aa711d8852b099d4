"""The C-ABI library loads without a GPU and exports every symbol that
include/fbstab_b200.h declares (no compute calls here); the engine entry points
fail loudly -- FBSTAB_ERR_NOGPU, never a CPU fallback -- when no device exists."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "fbstab_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fbstab_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    from fbstab_b200 import capi
    L = capi.lib()
    names = _declared()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # the Python mirror's own list is the same set
    assert sorted(capi.SYMBOLS) == names, sorted(set(capi.SYMBOLS) ^ set(names))


def test_generator_library_is_separate_from_the_engine():
    """bench.py's CPU reference arm takes its problem data from the generator
    library alone: importing fbstab_b200.problems does not map the CUDA engine."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import fbstab_b200.problems as p; "
            "d = p.random_dense_qp(4, 1, 3, count=2, config=2); "
            "maps = open('/proc/self/maps').read(); "
            "print('engine-mapped' if 'libfbstab_b200.so' in maps else 'generator-only')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "generator-only", out.stdout + out.stderr


def test_no_gpu_means_error_not_fallback():
    from fbstab_b200 import capi
    if capi.device_count() > 0:
        return
    L = capi.lib()
    h = C.c_void_p()
    assert L.fbstab_dense_batch_create(4, 1, 3, 8, 0, C.byref(h)) == capi.ERR_NOGPU
    assert L.fbstab_mpc_batch_create(5, 2, 1, 2, 8, 0, C.byref(h)) == capi.ERR_NOGPU
    L.fbstab_mpc_closed_loop_create.argtypes = ([C.c_int] * 7 + [C.c_void_p] * 14 +
                                                [C.c_int, C.POINTER(C.c_void_p)])
    z = np.zeros(64)
    p = z.ctypes.data
    assert L.fbstab_mpc_closed_loop_create(5, 2, 1, 2, 1, 0, 1, *([p] * 12), None, None, 4,
                                           C.byref(h)) == capi.ERR_NOGPU
    assert b"no CPU path" in L.fbstab_last_error()
