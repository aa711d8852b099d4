"""Phase-level cycle accounting of the dense-large kernel (development build).
Builds build/libfbstab_b200_phases.so with -DFBSTAB_PHASE_TIMERS, solves a
config-5 batch and prints the share of CTA cycles per phase.
Usage: python tools/phase_timers.py [batch]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

tbn = "64"
if "--tbn" in sys.argv:
    i = sys.argv.index("--tbn")
    tbn = sys.argv[i + 1]
    del sys.argv[i:i + 2]
lib = os.path.join(ROOT, "build", f"libfbstab_b200_phases{tbn}.so")
if "--build" in sys.argv or not os.path.exists(lib):
    objs = []
    procs = []
    for src in g.CUDA_SOURCES + g.HOST_SOURCES:
        obj = os.path.join(ROOT, "build", f"ph{tbn}_" + src + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["nvcc"] + g.NVCC_FLAGS + ["-DFBSTAB_PHASE_TIMERS", "-DFBS_DL_TBN=" + tbn, "-c",
                                      os.path.join(g.CSRC, src), "-o", obj]))
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                           "-o", lib] + objs)
    if "--build" in sys.argv:
        sys.exit(0)
os.environ["FBSTAB_B200_LIB"] = lib
import numpy as np  # noqa: E402
import torch  # noqa: E402
import fbstab_b200 as fb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
nz, nl, nv = 512, 128, 1024
d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=5, nthreads=16)
s = fb.FBstabDense(nz, nl, nv, max_batch=B)
dev = torch.device("cuda:0")
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
L = fb.capi.lib()
buf = (C.c_ulonglong * 32)()
for it in range(2):
    z = torch.zeros(B * nz, dtype=torch.float64, device=dev)
    l = torch.zeros(B * nl, dtype=torch.float64, device=dev)
    v = torch.zeros(B * nv, dtype=torch.float64, device=dev)
    L.fbstab_debug_phase_cycles(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, y = s.solve_batch(dd, z, l, v)
    e1.record()
    torch.cuda.synchronize()
    L.fbstab_debug_phase_cycles(buf, 0)
names = {1: "A'GammaA (DMMA tiles)", 2: "G rows / S init", 3: "diagonal-block Cholesky",
         4: "panel solves", 5: "trailing updates (DMMA)", 6: "rhs: A'(rv/mu)",
         7: "forward E", 8: "Schur fwd/bwd + W dl", 9: "backward E", 10: "dv, dy (A dz)",
         11: "residual: Hz, Gz", 12: "residual: G'l, A'v", 15: "other (axpy, PFB, norms, control)"}
tot = sum(buf)
o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
print(f"batch {B}: {e0.elapsed_time(e1):.1f} ms, newton mean {o['newton_iters'].mean():.1f}, "
      f"evals mean {o['residual_evals'].mean():.1f}")
for i in range(32):
    if buf[i]:
        print(f"  {names.get(i, i)!s:36s} {100.0 * buf[i] / tot:6.2f} %  "
              f"{buf[i] / B / 1e6:9.2f} Mcycles/instance")
