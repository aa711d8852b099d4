"""fbstab_b200 -- B200-native batched FBstab (proximally stabilised
Fischer-Burmeister QP solver).  The compute path is the CUDA library
libfbstab_b200.so behind the C-ABI in include/fbstab_b200.h; this package is
the thin host-side mirror of the reference's solver interface."""
from . import capi, problems, sharding  # noqa: F401
from .capi import EXIT_FLAGS, OUT_DTYPE, FbstabError, Options  # noqa: F401
from .solver import FBstabDense, FBstabMpc, FBstabSparse  # noqa: F401
from .closed_loop import ClosedLoopMpc  # noqa: F401
