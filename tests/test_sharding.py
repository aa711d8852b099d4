"""Multi-rank path on CPU: world_size-2 and -3 `gloo` process groups run the
partition -> solve -> pack -> gather logic of fbstab_b200.sharding with the CPU
oracle standing in for the GPU solve, and rank 0 must receive exactly the bytes
a single process produces (identical bytes for every G)."""
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
        res = fb.sharding.solve_sharded(torch, dist, _oracle_solve, _problem(), sizes,
                                        BATCH, (NZ, NL, NV))
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
            assert max(sharding.shard_sizes(batch, world)) == -(-batch // world)


def test_pack_unpack_round_trip():
    import torch
    from fbstab_b200 import sharding
    from fbstab_b200.capi import OUT_DTYPE
    rng = np.random.default_rng(3)
    count, cap = 3, 5
    z, l = rng.normal(size=count * NZ), rng.normal(size=count * NL)
    v, y = rng.normal(size=count * NV), rng.normal(size=count * NV)
    out = np.zeros(count, dtype=OUT_DTYPE)
    out["eflag"] = [0, 3, 4]
    out["residual"] = rng.normal(size=count)
    buf = sharding.pack(torch, z, l, v, y, out, count, cap, (NZ, NL, NV))
    assert buf.numel() == cap * sharding.record_bytes(NZ, NL, NV)
    z2, l2, v2, y2, out2 = sharding.unpack(buf, count, cap, (NZ, NL, NV))
    for a, b in ((z, z2), (l, l2), (v, v2), (y, y2), (out, out2)):
        assert a.tobytes() == b.tobytes()


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
