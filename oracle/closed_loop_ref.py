"""Checker for fbstab_b200.closed_loop (test infrastructure, like everything
under oracle/): the same receding-horizon loop in numpy around any batched
solver -- the tests and tests/closed_loop_bench.py pass the CPU oracle."""
import numpy as np


def closed_loop_reference(dims, data, T, solve, warm_start=True, shift=True):
    """The same loop in numpy around any batched solver `solve(dims, data, (z,l,v))
    -> (out, z, l, v)` -- the tests pass the CPU oracle."""
    N, nx, nu, nc = dims
    K = N + 1
    d = {k: a.copy() for k, a in data.items()}
    B = d["x0"].size // nx
    A = d["A"].reshape(B, N, nx, nx)[:, 0].transpose(0, 2, 1)
    Bm = d["B"].reshape(B, N, nu, nx)[:, 0].transpose(0, 2, 1)
    c = d["c"].reshape(B, N, nx)[:, 0]
    z, l, v = np.zeros(B * K * (nx + nu)), np.zeros(B * K * nx), np.zeros(B * K * nc)
    x = d["x0"].reshape(B, nx)
    X, U = np.zeros((B, T + 1, nx)), np.zeros((B, T, nu))
    X[:, 0] = x
    u_prev = np.zeros((B, nu))
    flags, newton = np.zeros((T, B), dtype=np.int32), np.zeros((T, B), dtype=np.int32)
    for t in range(T):
        if not warm_start:
            z[:], l[:], v[:] = 0.0, 0.0, 0.0
        elif shift and t > 0:
            for arr, w in ((z, nx + nu), (l, nx), (v, nc)):
                a = arr.reshape(B, K, w)
                a[:, :-1] = a[:, 1:].copy()
        out, z, l, v = solve(dims, d, (z, l, v))
        flags[t], newton[t] = out["eflag"], out["newton_iters"]
        # csrc/closed_loop.cu: a plant whose solve ends infeasible / failed holds its
        # previous input and restarts cold (the iterate holds a certificate, not a control)
        good = (out["status"] == 0) & ((out["eflag"] == 0) | (out["eflag"] == 2))
        u_prev[good] = z.reshape(B, K, nx + nu)[good, 0, nx:]
        for arr, w in ((z, nx + nu), (l, nx), (v, nc)):
            arr.reshape(B, K * w)[~good] = 0.0
        u0 = u_prev
        U[:, t] = u0
        x[:] = np.einsum("bij,bj->bi", A, x) + np.einsum("bij,bj->bi", Bm, u0) + c
        X[:, t + 1] = x
    return {"X": X, "U": U, "eflag": flags, "newton_iters": newton}
