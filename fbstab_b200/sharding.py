"""Multi-GPU sharding of a batch of independent QPs: the Python
mirror of the C-ABI's multi-GPU section (include/fbstab_b200.h, csrc/multi_gpu.cu).

Every QP instance is an independent solve (reference fbstab_dense.h:136-142:
one `Solve` = one problem), so a batch shards by contiguous instance ranges with
no collective on the data path: one process per GPU solves its range with its
own batch handle, and the only exchange is the result gather.  That gather is
C++ host code in the library -- one grouped NCCL send/recv per result array,
straight from each rank's result buffers into the root's global arrays over
NVLink, no packing and no padding.  `torch.distributed` is used for exactly one
thing: broadcasting the 128-byte NCCL unique id at start-up (the launcher's own
channel).  The CPU tests run the same partition / ordering logic under gloo with
host arrays (`gather_host`).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import OUT_DTYPE


def shard_range(batch, world, rank):
    """Contiguous index range [lo, hi) of `rank` (fbstab_multi_gpu_shard): the
    first batch % world ranks own one instance more than the others."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(batch, world):
    return [hi - lo for lo, hi in (shard_range(batch, world, r) for r in range(world))]


def slice_data(data, field_sizes, lo, hi):
    """The rows [lo, hi) of every field of an instance-major problem dict."""
    return {k: a[lo * field_sizes[k]:hi * field_sizes[k]] for k, a in data.items()}


def _bind(L):
    if getattr(L, "_fbstab_multi_gpu_bound", False):
        return L
    L.fbstab_multi_gpu_unique_id.argtypes = [C.c_char_p]
    L.fbstab_multi_gpu_create.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int,
                                          C.POINTER(C.c_void_p)]
    L.fbstab_multi_gpu_destroy.argtypes = [C.c_void_p]
    L.fbstab_multi_gpu_shard.argtypes = [C.c_int, C.c_int, C.c_long, C.POINTER(C.c_long),
                                         C.POINTER(C.c_long)]
    L.fbstab_multi_gpu_gather.argtypes = ([C.c_void_p, C.c_int, C.c_long] + [C.c_int] * 3 +
                                          [C.c_void_p] * 10 + [C.c_void_p])
    L._fbstab_multi_gpu_bound = True
    return L


def c_shard_range(batch, world, rank):
    """The same range from the library (tests check it equals shard_range)."""
    first, count = C.c_long(), C.c_long()
    capi.check(_bind(capi.lib()).fbstab_multi_gpu_shard(world, rank, batch, C.byref(first),
                                                        C.byref(count)))
    return first.value, first.value + count.value


class MultiGpu:
    """One rank of a one-process-per-GPU job (fbstab_multi_gpu_*)."""

    def __init__(self, rank, world, device, dist=None, torch=None):
        self.rank, self.world, self.device = rank, world, device
        self._h = C.c_void_p()
        L = _bind(capi.lib())
        ident = C.create_string_buffer(128)
        if world > 1:
            if rank == 0:
                capi.check(L.fbstab_multi_gpu_unique_id(ident))
            # the launcher's channel carries the id to the other ranks
            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else "cpu"
            t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).to(dev)
            dist.broadcast(t, src=0)
            ident = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
        capi.check(L.fbstab_multi_gpu_create(rank, world, ident, device, C.byref(self._h)))

    def close(self):
        if self._h:
            capi.lib().fbstab_multi_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def shard(self, global_batch):
        return shard_range(global_batch, self.world, self.rank)

    def gather(self, global_batch, sizes, shard, full=None, root=0, stream=None):
        """shard = (z, l, v, y, out) device tensors of this rank's instances;
        full = (Z, L, V, Y, OUT) device tensors of the whole batch on `root`
        (None elsewhere).  Enqueues on `stream`; returns nothing."""
        nz, nl, nv = sizes
        full = full if full is not None else (None,) * 5
        capi.check(capi.lib().fbstab_multi_gpu_gather(
            self._h, root, global_batch, nz, nl, nv, *[capi.ptr(t) for t in shard],
            *[capi.ptr(t) for t in full], stream))


# ---- host-array path (gloo): the partition / ordering logic on CPU ---------------
def gather_host(dist, parts, batch, dst=0, group=None):
    """parts = (z, l, v, y, out) numpy arrays of this rank's shard.  Returns the
    five arrays of the whole batch in global instance order on `dst`, None on the
    other ranks.  Shards arrive in rank order, which IS the global order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    objs = [None] * world if rank == dst else None
    dist.gather_object([np.ascontiguousarray(a) for a in parts], objs, dst=dst, group=group)
    if rank != dst:
        return None
    counts = shard_sizes(batch, world)
    assert [len(o[4]) for o in objs] == counts, "every rank must send exactly its range"
    return tuple(np.concatenate([o[k] for o in objs]) for k in range(5))


def solve_sharded(dist, solve_shard, data, field_sizes, batch, group=None):
    """Shards `batch` instances over the ranks of `group`, calls
    `solve_shard(data_shard, count) -> (z, l, v, y, out)` (numpy) on this rank's
    range and gathers the results to rank 0 (None on the other ranks).  `data`
    holds the whole batch on every rank; a rank only touches its own rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(batch, world, rank)
    if hi > lo:
        parts = solve_shard(slice_data(data, field_sizes, lo, hi), hi - lo)
    else:
        e = np.zeros(0)
        parts = (e, e, e, e, np.zeros(0, dtype=OUT_DTYPE))
    return gather_host(dist, parts, batch, group=group)
