"""Batched receding-horizon (closed-loop) MPC on top of FBstabMpc.

SURVEY.md section 8(f), rank 1: what the reference's
OcpGenerator::GetSimulationInputs (fbstab/test/ocp_generator.h:31-38,69;
ocp_generator.cc:56-71) exists for and what its README calls "can be easily
warmstarted" (README.md:20).  B plants, each with its own OCP data and state,
are simulated for T control steps:

    solve the OCP from x(t), warm-started with the previous solution
    apply the first input:  x(t+1) = A x(t) + B u0(t) + c      (the plant model
    is stage 0 of the instance's own OCP data, as in the reference generator,
    whose Asim/Bsim are the OCP's A/B)

Everything stays on the device between steps: the problem data and the iterate
(z, l, v) are torch CUDA tensors handed to the C-ABI as device pointers (used in
place, the call only enqueues the kernel), the state update is a batched
mat-vec, and the host reads back the logged trajectory once at the end.
"""
import numpy as np

from . import problems
from .capi import OUT_DTYPE
from .solver import FBstabMpc


def _shift(t, B, K, w):
    """Receding-horizon shift of a stage-major iterate: stage i <- stage i+1,
    the last stage is repeated."""
    a = t.view(B, K, w)
    a[:, :-1] = a[:, 1:].clone()


class ClosedLoopMpc:
    def __init__(self, dims, data, device=0):
        import torch
        self.torch = torch
        self.N, self.nx, self.nu, self.nc = dims
        self.B = data["x0"].size // self.nx
        self.dev = torch.device(f"cuda:{device}")
        self.solver = FBstabMpc(*dims, max_batch=self.B, device=device)
        self.data = {k: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
                     for k, a in data.items()}
        B, N, nx, nu = self.B, self.N, self.nx, self.nu
        # plant model = stage 0 of every instance's own OCP (column-major blocks)
        self.A = self.data["A"].view(B, N, nx, nx)[:, 0].transpose(1, 2).contiguous()
        self.Bm = self.data["B"].view(B, N, nu, nx)[:, 0].transpose(1, 2).contiguous()
        self.c = self.data["c"].view(B, N, nx)[:, 0].contiguous()
        self.x_init = self.data["x0"].clone()

    def run(self, T, warm_start=True, shift=True):
        """Simulates T steps.  Returns a dict with X (B,T+1,nx), U (B,T,nu), the
        per-step exit flags / Newton iterations (T,B) and the device time in ms."""
        torch = self.torch
        s, B, K = self.solver, self.B, self.N + 1
        nx, nu, nc = self.nx, self.nu, self.nc
        z = torch.zeros(B * s.nz, dtype=torch.float64, device=self.dev)
        l = torch.zeros(B * s.nl, dtype=torch.float64, device=self.dev)
        v = torch.zeros(B * s.nv, dtype=torch.float64, device=self.dev)
        y = torch.zeros(B * s.nv, dtype=torch.float64, device=self.dev)
        outs = torch.zeros(T, B * OUT_DTYPE.itemsize, dtype=torch.uint8, device=self.dev)
        X = torch.zeros(B, T + 1, nx, dtype=torch.float64, device=self.dev)
        U = torch.zeros(B, T, nu, dtype=torch.float64, device=self.dev)
        self.data["x0"].copy_(self.x_init)
        x = self.data["x0"].view(B, nx)
        X[:, 0] = x
        stream = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(T):
            if not warm_start:
                z.zero_(), l.zero_(), v.zero_()
            elif shift and t > 0:
                _shift(z, B, K, nx + nu), _shift(l, B, K, nx), _shift(v, B, K, nc)
            s.solve_batch(self.data, z, l, v, y=y, out=outs[t], stream=stream)
            u0 = z.view(B, K, nx + nu)[:, 0, nx:]
            U[:, t] = u0
            xn = (torch.bmm(self.A, x.unsqueeze(2)) + torch.bmm(self.Bm, u0.unsqueeze(2))
                  ).squeeze(2) + self.c
            x.copy_(xn)  # in place: data["x0"] is the next solve's initial state
            X[:, t + 1] = x
        e1.record()
        torch.cuda.synchronize()
        o = np.frombuffer(outs.cpu().numpy().tobytes(), dtype=OUT_DTYPE).reshape(T, B)
        return {"X": X.cpu().numpy(), "U": U.cpu().numpy(), "eflag": o["eflag"].copy(),
                "newton_iters": o["newton_iters"].copy(), "status": o["status"].copy(),
                "ms": e0.elapsed_time(e1), "path": s.path}
