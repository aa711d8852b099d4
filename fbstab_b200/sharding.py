"""Multi-GPU sharding of a batch of independent QPs.

Every QP instance is an independent solve (reference fbstab_dense.h:136-142:
one `Solve` = one problem), so a batch shards by instance with no collective
on the data path: one process per GPU solves a contiguous index range, and the
only exchange is the gather of the packed results to one rank over
`torch.distributed` (NCCL on NVLink for CUDA tensors, gloo for the CPU tests).

Packed record of one shard, as one byte buffer:
    z (nz) | l (nl) | v (nv) | y (nv)   float64, instance-major
    out                                   OUT_DTYPE records (48 bytes each)
Shards are padded to the size of the largest one so that a plain `gather`
works; rank 0 drops the padding and concatenates in rank order, which is the
global instance order, so the result is byte-identical for every world size.
"""
import numpy as np

from .capi import OUT_DTYPE


def shard_range(batch, world, rank):
    """Contiguous index range [lo, hi) of `rank`: ceil(batch/world) instances
    per rank, the last ranks possibly short or empty."""
    per = -(-batch // world)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def shard_sizes(batch, world):
    return [hi - lo for lo, hi in (shard_range(batch, world, r) for r in range(world))]


def record_bytes(nz, nl, nv):
    """Bytes of one instance in the packed result buffer."""
    return 8 * (nz + nl + 2 * nv) + OUT_DTYPE.itemsize


def slice_data(data, field_sizes, lo, hi):
    """The rows [lo, hi) of every field of an instance-major problem dict."""
    return {k: a[lo * field_sizes[k]:hi * field_sizes[k]] for k, a in data.items()}


def pack(torch, z, l, v, y, out, count, capacity, sizes):
    """Packs `count` solved instances into a uint8 tensor with room for
    `capacity` instances (same device as z).  `out` is a uint8 tensor (device
    path) or a structured numpy array (host path)."""
    nz, nl, nv = sizes
    if isinstance(z, np.ndarray):
        z, l, v, y = (torch.from_numpy(np.ascontiguousarray(a)) for a in (z, l, v, y))
    if isinstance(out, np.ndarray):
        out = torch.from_numpy(np.frombuffer(out.tobytes(), dtype=np.uint8).copy())
    buf = torch.zeros(capacity * record_bytes(nz, nl, nv), dtype=torch.uint8, device=z.device)
    off = 0
    for t, n in ((z, nz), (l, nl), (v, nv), (y, nv)):
        nb = 8 * n * count
        if nb:
            buf[off:off + nb].copy_(t[:n * count].contiguous().view(torch.uint8))
        off += 8 * n * capacity
    nb = OUT_DTYPE.itemsize * count
    buf[off:off + nb].copy_(out[:nb])
    return buf


def unpack(buf, count, capacity, sizes):
    """Inverse of pack(): numpy (z, l, v, y, out) of the first `count` instances."""
    nz, nl, nv = sizes
    raw = buf.cpu().numpy() if hasattr(buf, "cpu") else np.asarray(buf)
    res, off = [], 0
    for n in (nz, nl, nv, nv):
        res.append(np.frombuffer(raw[off:off + 8 * n * count].tobytes(), dtype=np.float64))
        off += 8 * n * capacity
    out = np.frombuffer(raw[off:off + OUT_DTYPE.itemsize * count].tobytes(), dtype=OUT_DTYPE)
    return (*res, out)


def gather_results(torch, dist, packed, batch, sizes, dst=0, group=None):
    """Gathers every rank's packed shard to `dst`.  Returns (z, l, v, y, out)
    for the whole batch in global instance order on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = shard_sizes(batch, world)
    cap = max(counts)
    assert packed.numel() == cap * record_bytes(*sizes), "shards must be padded to capacity"
    bufs = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [unpack(b, c, cap, sizes) for b, c in zip(bufs, counts)]
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(5))


def solve_sharded(torch, dist, solve_shard, data, field_sizes, batch, sizes, group=None,
                  device=None):
    """Shards `batch` instances over the ranks of `group`, calls
    `solve_shard(data_shard, count) -> (z, l, v, y, out)` on this rank's range
    and gathers the results to rank 0 (None on the other ranks).

    `data` holds the WHOLE batch on every rank (instance-major numpy arrays or
    tensors); a rank only touches its own rows.  `solve_shard` is normally
    `lambda d, n: FBstabDense(...).solve_batch(...)`; the CPU tests pass a
    stand-in so that the partition + pack + gather logic runs under gloo."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(batch, world, rank)
    cap = max(shard_sizes(batch, world))
    nz, nl, nv = sizes
    if hi > lo:
        z, l, v, y, out = solve_shard(slice_data(data, field_sizes, lo, hi), hi - lo)
    else:
        e = np.zeros(0)
        z, l, v, y, out = e, e, e, e, np.zeros(0, dtype=OUT_DTYPE)
    packed = pack(torch, z, l, v, y, out, hi - lo, cap, sizes)
    if device is not None:
        packed = packed.to(device)
    return gather_results(torch, dist, packed, batch, sizes, group=group)
