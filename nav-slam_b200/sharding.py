"""Multi-GPU partitioning of the front end.

The path shards in exactly two ways, both without touching the kernels:

  * independent sequences (config 5a): sequence s runs on rank s % world; nothing is exchanged on
    the data path (poses / CSVs are gathered by the host afterwards);
  * large-map queries (config 5b): the map is replicated (rank 0 broadcasts the points, every
    rank builds the same flat kd-tree -- the build is deterministic and takes milliseconds), the
    query set is cut into `world` contiguous shards, and ONE all_gather of packed (dist:f64, idx:int32)
    records reassembles the answers on every rank.

Per-row maps of the SLAM step (<= cols points) are never worth sharding.  One process per GPU;
torch.distributed (NCCL over NVLink on GPUs, gloo on CPU for the tests) is only the plumbing.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items; the first n % world ranks get one item more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assign_sequences(n_seq: int, world: int) -> List[List[int]]:
    """Round-robin sequence -> rank assignment (config 5a); ranks may get an empty list."""
    return [list(range(r, n_seq, world)) for r in range(world)]


def max_shard(n: int, world: int) -> int:
    return (n + world - 1) // world


def sharded_nn(nn_fn: Callable, queries, *, group=None, device=None):
    """Answer `queries` ([nq,3] float64 torch tensor, identical on every rank) with this rank
    computing only its shard through nn_fn(q_shard) -> (idx int32 [m], dist float64 [m]) and one
    all_gather putting the full (idx, dist) on every rank.  Works with any backend: the tensors
    live on `device` (cuda for NCCL, cpu for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nq = int(queries.shape[0])
    lo, hi = shard_bounds(nq, world, rank)
    idx_s, dist_s = nn_fn(queries[lo:hi])
    if world == 1:
        return idx_s, dist_s
    device = device if device is not None else queries.device
    m = max_shard(nq, world)
    # pack (dist, idx) into one padded float64 buffer so that a single collective moves both
    buf = torch.zeros((m, 2), dtype=torch.float64, device=device)
    buf[: hi - lo, 0] = dist_s.to(torch.float64)
    buf[: hi - lo, 1] = idx_s.to(torch.float64)  # int32 is exact in float64
    out = torch.empty((world, m, 2), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out.view(world * m, 2), buf, group=group)
    idx = torch.empty(nq, dtype=torch.int32, device=device)
    dd = torch.empty(nq, dtype=torch.float64, device=device)
    for r in range(world):
        a, b = shard_bounds(nq, world, r)
        dd[a:b] = out[r, : b - a, 0]
        idx[a:b] = out[r, : b - a, 1].to(torch.int32)
    return idx, dd


_PACKED = {}   # (device, world, shard size) -> byte buffer of the packed exchange


def sharded_nn_into(nn_into: Callable, queries, idx_out, dist_out, *, group=None):
    """The same exchange for the hot loop, with ONE collective and no packing kernels in front of it:
    idx_out [nq] int32 and dist_out [nq] float64 are preallocated on the collective's device (same shapes on
    every rank).  A persistent byte buffer holds, for every rank r, the record [dist: m x f64 | idx: m x i32]
    of its contiguous shard (m = nq / world); nn_into(q_shard, idx_view, dist_view) writes this rank's
    answers straight into the views of its own record, one in-place all_gather (the record is the rank's
    send buffer) completes the buffer on every rank, and two strided copies unpack it into idx_out /
    dist_out.  Needs equal shards (nq % world == 0, m even); otherwise the padded path of sharded_nn()."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nq = int(queries.shape[0])
    lo, hi = shard_bounds(nq, world, rank)
    if world == 1:
        nn_into(queries, idx_out, dist_out)
        return idx_out, dist_out
    m = nq // world
    if nq % world != 0 or m % 2 != 0:
        def nn_fn(qs):
            k = int(qs.shape[0])
            nn_into(qs, idx_out[lo:lo + k], dist_out[lo:lo + k])
            return idx_out[lo:lo + k].clone(), dist_out[lo:lo + k].clone()
        i, d = sharded_nn(nn_fn, queries, group=group)
        idx_out.copy_(i)
        dist_out.copy_(d)
        return idx_out, dist_out
    key = (str(idx_out.device), world, m)
    packed = _PACKED.get(key)
    if packed is None:
        packed = _PACKED[key] = torch.empty((world, 12 * m), dtype=torch.uint8, device=idx_out.device)
    mine = packed[rank]
    nn_into(queries[lo:hi], mine[8 * m:].view(torch.int32), mine[: 8 * m].view(torch.float64))
    on_cpu = idx_out.device.type == "cpu"   # gloo wants distinct send / receive storage
    dist.all_gather_into_tensor(packed.view(-1), mine.clone() if on_cpu else mine, group=group)
    dist_out.view(world, m).copy_(packed[:, : 8 * m].view(torch.float64))
    idx_out.view(world, m).copy_(packed[:, 8 * m:].view(torch.int32))
    return idx_out, dist_out


def broadcast_points(points, *, src: int = 0, group=None):
    """Replicate the map points (torch tensor [n,3] float64 on the collective's device)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(points, src=src, group=group)
    return points


def gather_sequence_results(local: Sequence[Tuple[int, np.ndarray]], world: int, *, group=None):
    """Host-side gather of per-sequence results [(sequence id, array)] to every rank (config 5a:
    poses per sequence).  Not on the data path; uses all_gather_object."""
    import torch.distributed as dist

    if not dist.is_initialized() or world == 1:
        return dict(local)
    box = [None] * world
    dist.all_gather_object(box, list(local), group=group)
    merged = {}
    for part in box:
        for sid, arr in part:
            merged[sid] = arr
    return merged


class PeerGather:
    """Config 5b over peer memory (include/navslam_b200.h, nav_peer_*): the answers of every rank's query shard
    are stored by the search kernel itself into the result buffers of all ranks of the node (CUDA IPC mappings,
    NVLink stores), so no collective follows the search.  torch.distributed only carries the 64-byte memory
    handles once, at construction.  Raises NavError when the buffers cannot be mapped (no peer access between the
    GPUs, or more than eight ranks): callers then use sharded_nn_into()."""

    def __init__(self, lib, device: int, nq_total: int, *, group=None):
        import ctypes as C

        import torch.distributed as dist
        from .binding import NavError
        self.L, self.device, self.nq = lib, device, int(nq_total)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        handle = C.create_string_buffer(64)
        self.h = lib.nav_peer_create(device, self.nq, handle)
        # every step that can fail on one rank is followed by an exchange of the outcome, so that all ranks raise
        # together instead of one leaving the others inside a collective
        handles = [None] * self.world
        mine = handle.raw if self.h else None
        if self.world > 1:
            dist.all_gather_object(handles, mine, group=group)
        else:
            handles[0] = mine
        if any(h is None for h in handles):
            msg = lib.nav_last_error().decode() if not self.h else "another rank could not allocate its peer buffer"
            self.close()
            raise NavError(msg)
        rc = lib.nav_peer_connect(self.h, self.world, self.rank, b"".join(handles))
        ok = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(ok, rc == 0, group=group)   # all ranks use the peer path, or none does
        else:
            ok[0] = rc == 0
        if not all(ok):
            msg = lib.nav_last_error().decode() if rc else "another rank could not map the peer buffers"
            self.close()
            raise NavError(msg)

    def nn(self, tree, queries, stream: int):
        """queries: [nq_total, 3] float64 CUDA tensor, identical on every rank.  Returns (idx int32 [nq_total],
        dist float64 [nq_total]) as tensors aliasing this rank's buffer: complete in stream order, valid until the
        call after the next one."""
        import ctypes as C

        import torch
        from .binding import NavError
        lo, hi = shard_bounds(self.nq, self.world, self.rank)
        shard = queries[lo:hi]
        pi, pd = C.c_void_p(), C.c_void_p()
        rc = self.L.nav_kdtree_nn_allgather_dev(tree.h, self.h, shard.data_ptr(), lo, hi - lo, C.byref(pi), C.byref(pd),
                                                stream)
        if rc:
            raise NavError(self.L.nav_last_error().decode())

        class _Raw:
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev = queries.device
        return (torch.as_tensor(_Raw(pi.value, self.nq, "<i4"), device=dev),
                torch.as_tensor(_Raw(pd.value, self.nq, "<f8"), device=dev))

    def nn_sharded_map(self, tree_part, queries, idx_offset: int, stream: int):
        """The map sharded instead of the queries: tree_part holds this rank's points (points idx_offset ... of the
        whole map), queries are all nq_total queries on every rank.  Same return value and answers as nn()."""
        import ctypes as C

        import torch
        from .binding import NavError
        pi, pd = C.c_void_p(), C.c_void_p()
        rc = self.L.nav_kdtree_nn_sharded_map_dev(tree_part.h, self.h, queries.data_ptr(), self.nq, int(idx_offset),
                                                  C.byref(pi), C.byref(pd), stream)
        if rc:
            raise NavError(self.L.nav_last_error().decode())

        class _Raw:
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev = queries.device
        return (torch.as_tensor(_Raw(pi.value, self.nq, "<i4"), device=dev),
                torch.as_tensor(_Raw(pd.value, self.nq, "<f8"), device=dev))

    def check(self):
        from .binding import NavError
        if self.L.nav_peer_check(self.h):
            raise NavError(self.L.nav_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.nav_peer_destroy(self.h)
            self.h = None
