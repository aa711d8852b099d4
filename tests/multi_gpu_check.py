"""The C-ABI's NCCL result gather on real GPUs (one process per GPU, torchrun):
every rank solves its shard of a seeded batch, fbstab_multi_gpu_gather collects the
rows on rank 0, and rank 0 compares every byte with its own single-GPU solve of the
whole batch ("same batch at 1/2/4/8 shards must return identical
bytes").  Also runs the single-process fbstab_dense_multi_gpu_solve on all visible
devices.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi_gpu_check.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import fbstab_b200 as fb
from fbstab_b200 import capi, sharding


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mg = sharding.MultiGpu(rank, world, local, dist=dist, torch=torch)
    ok = True
    for (nz, nl, nv, B, cfg) in ((32, 8, 64, 4099, 2), (9, 3, 4, 7, 21), (50, 10, 100, 130, 1)):
        lo, hi = mg.shard(B)
        n = hi - lo
        d = fb.problems.random_dense_qp(nz, nl, nv, count=max(n, 1), config=cfg, first=lo)
        dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
        s = fb.FBstabDense(nz, nl, nv, max_batch=max(n, 1), device=local)
        f64 = lambda m: torch.zeros(m, dtype=torch.float64, device=dev)
        z, l, v, y = f64(n * nz), f64(n * nl), f64(n * nv), f64(n * nv)
        out = torch.zeros(n * fb.OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream()
        if n:
            s.solve_batch(dd, z, l, v, y=y, out=out, stream=stream.cuda_stream)
        full = None
        if rank == 0:
            full = (f64(B * nz), f64(B * nl), f64(B * nv), f64(B * nv),
                    torch.zeros(B * fb.OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev))
        mg.gather(B, (nz, nl, nv), (z, l, v, y, out), full, root=0, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        if rank == 0:
            d1 = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=cfg)
            s1 = fb.FBstabDense(nz, nl, nv, max_batch=B, device=local)
            dd1 = {k: torch.from_numpy(a).to(dev) for k, a in d1.items()}
            z1, l1, v1 = f64(B * nz), f64(B * nl), f64(B * nv)
            o1, y1 = s1.solve_batch(dd1, z1, l1, v1)
            torch.cuda.synchronize()
            same = all(torch.equal(a, b) for a, b in zip(full[:4], (z1, l1, v1, y1)))
            oa = np.frombuffer(full[4].cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE).copy()
            ob = np.frombuffer(o1.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE).copy()
            oa["solve_time"] = ob["solve_time"] = 0
            same = same and oa.tobytes() == ob.tobytes()
            print(f"gather {nz}/{nl}/{nv} x {B} on {world} ranks: identical bytes = {same}",
                  flush=True)
            ok = ok and same
    if rank == 0:
        # single process, all visible devices, host buffers (the facade's SolveBatch(..., devices))
        L = capi.lib()
        nd = capi.device_count()
        nz, nl, nv, B = 32, 8, 64, 3001
        d = fb.problems.random_dense_qp(nz, nl, nv, count=B, config=2)
        devs = (C.c_int * nd)(*range(nd))
        h = C.c_void_p()
        L.fbstab_dense_multi_gpu_create.argtypes = [C.c_int, C.c_void_p] + [C.c_int] * 3 + [
            C.c_long, C.POINTER(C.c_void_p)]
        L.fbstab_dense_multi_gpu_solve.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 11
        L.fbstab_dense_multi_gpu_destroy.argtypes = [C.c_void_p]
        capi.check(L.fbstab_dense_multi_gpu_create(nd, devs, nz, nl, nv, B, C.byref(h)))
        z, l, v, y = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv), np.zeros(B * nv)
        out = np.zeros(B, dtype=fb.OUT_DTYPE)
        capi.check(L.fbstab_dense_multi_gpu_solve(
            h, B, *[capi.ptr(d[k]) for k in fb.problems.DENSE_FIELDS], capi.ptr(z), capi.ptr(l),
            capi.ptr(v), capi.ptr(y), capi.ptr(out)))
        L.fbstab_dense_multi_gpu_destroy(h)
        s1 = fb.FBstabDense(nz, nl, nv, max_batch=B, device=local)
        z1, l1, v1 = np.zeros(B * nz), np.zeros(B * nl), np.zeros(B * nv)
        o1, y1 = s1.solve_batch(d, z1, l1, v1)
        same = all(a.tobytes() == b.tobytes() for a, b in ((z, z1), (l, l1), (v, v1), (y, y1)))
        same = same and (out["eflag"] == o1["eflag"]).all() and (
            out["newton_iters"] == o1["newton_iters"]).all()
        print(f"single-process solve on {nd} device(s): identical bytes = {same}", flush=True)
        ok = ok and same
        # the sparse entry (fbstab_sparse_multi_gpu_solve): the servo OCP as general sparse QPs
        Bs = 1003
        dims, dm = fb.problems.ocp_batch("servo_motor", 20, count=Bs, config=3, rho=0.02)
        (snz, snl, snv), pat, vals = fb.problems.ocp_as_sparse_qp(dims, dm, Bs)
        ss = fb.FBstabSparse(snz, snl, snv, pat, max_batch=Bs, device=local)
        z, l, v = np.zeros(Bs * snz), np.zeros(Bs * snl), np.zeros(Bs * snv)
        out, y = ss.solve_batch_devices(vals, z, l, v, list(range(nd)))
        z1, l1, v1 = np.zeros(Bs * snz), np.zeros(Bs * snl), np.zeros(Bs * snv)
        o1, y1 = ss.solve_batch(vals, z1, l1, v1)
        same = all(a.tobytes() == b.tobytes() for a, b in ((z, z1), (l, l1), (v, v1), (y, y1)))
        same = same and (out["eflag"] == o1["eflag"]).all() and (
            out["newton_iters"] == o1["newton_iters"]).all()
        print(f"single-process sparse solve on {nd} device(s): identical bytes = {same}",
              flush=True)
        ok = ok and same
    mg.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
