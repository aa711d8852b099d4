"""Multi-rank path on CPU: world_size-2 and -3 `gloo` process groups run the
partition -> solve -> gather logic of fbstab_b200.sharding with the CPU oracle
standing in for the GPU solve, and rank 0 must receive exactly the bytes a single
process produces (identical bytes for every G).  The NCCL gather
of the C-ABI itself (fbstab_multi_gpu_gather) needs GPUs: tests/multi_gpu_check.py
runs it under torchrun and compares the gathered bytes with a single-GPU solve;
here the library's shard arithmetic and argument checks are covered."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NZ, NL, NV, BATCH = 6, 2, 9, 11  # BATCH not divisible by 2 or 3: ragged shards


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    import fbstab_b200 as fb
    return fb.problems.random_dense_qp(NZ, NL, NV, count=BATCH, config=21)


def _oracle_solve(d, count):
    from oracle import binding as ob
    oo, z, l, v, y = ob.dense_solve_batch(
        NZ, NL, NV, *[d[k] for k in ("H", "f", "G", "h", "A", "b")])
    return z, l, v, y, oo


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import fbstab_b200 as fb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = {"H": NZ * NZ, "f": NZ, "G": NL * NZ, "h": NL, "A": NV * NZ, "b": NV}
        res = fb.sharding.solve_sharded(dist, _oracle_solve, _problem(), sizes, BATCH)
        if rank == 0:
            q.put([a.tobytes() for a in res])
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    from fbstab_b200 import sharding
    for batch in (0, 1, 7, 64, 1000, 65536):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(batch, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert sum(sharding.shard_sizes(batch, world)) == batch
            sizes = sharding.shard_sizes(batch, world)
            assert max(sizes) == -(-batch // world) and max(sizes) - min(sizes) <= 1
            # the C-ABI's fbstab_multi_gpu_shard is the same arithmetic
            for k in range(world):
                assert sharding.c_shard_range(batch, world, k) == r[k]


def test_multi_gpu_entry_points_reject_bad_arguments():
    """No GPU here: argument checks and the no-GPU error of the multi-GPU C-ABI."""
    import ctypes as C
    from fbstab_b200 import capi, sharding
    L = sharding._bind(capi.lib())
    first, count = C.c_long(), C.c_long()
    assert L.fbstab_multi_gpu_shard(0, 0, 10, C.byref(first), C.byref(count)) == capi.ERR_INVALID
    assert L.fbstab_multi_gpu_shard(2, 2, 10, C.byref(first), C.byref(count)) == capi.ERR_INVALID
    h = C.c_void_p()
    rc = L.fbstab_multi_gpu_create(0, 1, None, 0, C.byref(h))
    if capi.device_count() == 0:
        assert rc == capi.ERR_NOGPU and b"no CPU path" in L.fbstab_last_error()
    else:
        assert rc == capi.OK
        L.fbstab_multi_gpu_destroy(h)
    assert L.fbstab_multi_gpu_create(3, 2, None, 0, C.byref(h)) == capi.ERR_INVALID


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_solve_equals_single_process(world):
    import torch.multiprocessing as mp
    single = [a.tobytes() for a in _oracle_solve(_problem(), BATCH)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # solve_time is wall clock: compare everything else byte for byte
    from fbstab_b200.capi import OUT_DTYPE
    for k in range(4):
        assert got[k] == single[k]
    a = np.frombuffer(got[4], dtype=OUT_DTYPE).copy()
    b = np.frombuffer(single[4], dtype=OUT_DTYPE).copy()
    a["solve_time"] = b["solve_time"] = 0
    assert a.tobytes() == b.tobytes()
