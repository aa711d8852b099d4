"""Batched receding-horizon (closed-loop) MPC: Python mirror of the C-ABI's
fbstab_mpc_closed_loop_* (include/fbstab_b200.h, csrc/closed_loop.cu).

SURVEY.md section 8(f), rank 1: what the reference's
OcpGenerator::GetSimulationInputs (fbstab/test/ocp_generator.h:31-38,69;
ocp_generator.cc:56-71) exists for and what its README calls "can be easily
warmstarted" (README.md:20).  B plants are simulated for T control steps; the
whole loop -- shift of the warm start, solve, plant update, logging -- runs in
the library on the device, this class only hands over the data and reads the
logged trajectory back once at the end.
"""
import ctypes as C

import numpy as np

from . import capi, problems
from .capi import OUT_DTYPE


def _bind(L):
    if getattr(L, "_fbstab_closed_loop_bound", False):
        return L
    L.fbstab_mpc_closed_loop_create.argtypes = ([C.c_int] * 7 + [C.c_void_p] * 14 +
                                                [C.c_int, C.POINTER(C.c_void_p)])
    L.fbstab_mpc_closed_loop_destroy.argtypes = [C.c_void_p]
    L.fbstab_mpc_closed_loop_set_options.argtypes = [C.c_void_p, C.POINTER(capi.Options)]
    L.fbstab_mpc_closed_loop_reset.argtypes = [C.c_void_p, C.c_void_p]
    L.fbstab_mpc_closed_loop_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.fbstab_mpc_closed_loop_run.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    L.fbstab_mpc_closed_loop_path.argtypes = [C.c_void_p]
    L.fbstab_mpc_closed_loop_path.restype = C.c_char_p
    L._fbstab_closed_loop_bound = True
    return L


class ClosedLoopMpc:
    """data: the wire-format dict of problems.ocp_batch.  shared=True: the 11
    sequences hold ONE copy for all plants and data["x0"] the B initial states."""

    def __init__(self, dims, data, device=0, shared=False, Asim=None, Bsim=None, max_steps=64):
        self.N, self.nx, self.nu, self.nc = dims
        self.B = data["x0"].size // self.nx
        self.max_steps = max_steps
        self.device = device
        self._h = C.c_void_p()
        self._L = _bind(capi.lib())
        keep = [np.ascontiguousarray(data[k], dtype=np.float64) for k in problems.MPC_FIELDS]
        plant = [None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)
                 for a in (Asim, Bsim)]  # column-major
        capi.check(self._L.fbstab_mpc_closed_loop_create(
            self.N, self.nx, self.nu, self.nc, self.B, device, 1 if shared else 0,
            *[capi.ptr(a) for a in keep], capi.ptr(plant[0]), capi.ptr(plant[1]), max_steps,
            C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.fbstab_mpc_closed_loop_destroy(h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def update_options(self, opts):
        capi.check(self._L.fbstab_mpc_closed_loop_set_options(self._h, C.byref(opts)))

    @property
    def path(self):
        return self._L.fbstab_mpc_closed_loop_path(self._h).decode()

    def run(self, T, warm_start=True, stream=None, time_it=True):
        """Simulates T steps from x_init.  Returns a dict with X (B,T+1,nx), U (B,T,nu),
        the per-step exit flags / Newton iterations (T,B) and the device time in ms."""
        B, nx, nu = self.B, self.nx, self.nu
        X, U = np.zeros((B, T + 1, nx)), np.zeros((B, T, nu))
        out = np.zeros((T, B), dtype=OUT_DTYPE)
        ms = None
        if time_it:
            import torch
            # (the stream and the events of the solver's OWN device, whatever the current one is)
            with torch.cuda.device(self.device):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
            s = torch.cuda.current_stream(self.device)
            stream = s.cuda_stream
            # device time of the loop alone: steps only, logs are read afterwards
            capi.check(self._L.fbstab_mpc_closed_loop_reset(self._h, stream))
            e0.record(s)
            for _ in range(T):
                capi.check(self._L.fbstab_mpc_closed_loop_step(self._h, 1 if warm_start else 0, stream))
            e1.record(s)
            torch.cuda.synchronize(self.device)
            ms = e0.elapsed_time(e1)
        capi.check(self._L.fbstab_mpc_closed_loop_run(self._h, T, 1 if warm_start else 0,
                                                       capi.ptr(X), capi.ptr(U), capi.ptr(out), stream))
        return {"X": X, "U": U, "eflag": out["eflag"].copy(),
                "newton_iters": out["newton_iters"].copy(), "status": out["status"].copy(),
                "ms": ms, "path": self.path}
