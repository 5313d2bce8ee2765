"""Multi-GPU plumbing: one process per GPU, torch.distributed for bootstrap only.

The data path never goes through torch: gradient all-reduce is `b200_all_reduce` (NCCL on a
dedicated stream inside libburn_b200.so).  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used to agree on the NCCL unique id, for barriers and for max-over-ranks timing —
the role `DistributedContext::init` / the gradient-sync server bootstrap plays in the reference
(crates/burn-tensor/src/tensor/distributed.rs:30-38,
 crates/burn-backend/src/backend/distributed/server.rs:60-139).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

UNIQUE_ID_BYTES = 128


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of `n_units` independent units (batch rows, tensors…)
    across ranks — the split `split_dataloader` performs per device
    (crates/burn-train/src/learner/supervised/strategies/ddp/strategy.rs:91)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_unique_id(make_id: Callable[[], bytes], rank: int, device=None) -> bytes:
    """Rank 0 produces the 128-byte NCCL unique id; everyone receives it."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        if len(raw) != UNIQUE_ID_BYTES:
            raise ValueError("unique id must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class Communicator:
    """Owns the per-process b200_comm (NCCL communicator + collective stream)."""

    def __init__(self, rank: int, world: int, device=None):
        from . import _abi as abi
        self._abi = abi
        self.lib = abi.load()
        self.rank, self.world = rank, world

        def make_id() -> bytes:
            raw = (C.c_uint8 * UNIQUE_ID_BYTES)()
            abi.check(self.lib.b200_comm_unique_id(raw))
            return bytes(raw)

        uid = broadcast_unique_id(make_id, rank, device) if world > 1 else make_id()
        self.handle = C.c_void_p()
        arr = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(uid)
        abi.check(self.lib.b200_comm_init(C.byref(self.handle), arr, rank, world))

    def all_reduce(self, tensor, mean: bool = True, producer_stream=None) -> None:
        """In-place all-reduce of one gradient tensor (DistributedOps::all_reduce)."""
        abi = self._abi
        abi.check(self.lib.b200_all_reduce(self.handle, tensor.data_ptr(), tensor.numel, tensor.dtype,
                                           abi.REDUCE_MEAN if mean else abi.REDUCE_SUM, producer_stream))

    def all_reduce_bucket(self, tensors: Sequence, mean: bool = True, producer_stream=None) -> None:
        abi = self._abi
        n = len(tensors)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        counts = (C.c_uint64 * n)(*[t.numel for t in tensors])
        abi.check(self.lib.b200_all_reduce_multi(self.handle, ptrs, counts, n, tensors[0].dtype,
                                                 abi.REDUCE_MEAN if mean else abi.REDUCE_SUM, producer_stream))

    def sync(self, consumer_stream=None) -> None:
        """DistributedOps::sync_collective: later work on `consumer_stream` waits for the collectives."""
        self._abi.check(self.lib.b200_collective_sync(self.handle, consumer_stream))

    def close(self) -> None:
        if self.handle:
            self.lib.b200_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def gather_bytes(raw: bytes, rank: int, world: int, device=None) -> bytes:
    """All-gather of one fixed-size byte string per rank (host bootstrap only: IPC handles, ids)."""
    import torch
    import torch.distributed as dist
    mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device) if device is not None else \
        torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


class PeerRegion:
    """Borrowed storage inside a peer region (DeviceTensor.storage duck type: never freed by a view)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.ptr, self.nbytes, self.owner, self.stream = C.c_void_p(ptr), nbytes, owner, None


class PeerGroup:
    """One rank's view of the peer-memory group (include/burn_b200.h, burn_b200/csrc/peer.cu): a cudaMalloc'd region
    holding the flat gradient / parameter buckets at the same offsets on every rank, mapped into all ranks over
    NVLink.  `all_reduce` is DistributedOps::all_reduce without NCCL; `adam` is the all-reduce fused with the
    optimizer step on a 1/N shard (see peer.cu).  One process per GPU: handles travel through torch.distributed."""

    def __init__(self, rank: int, world: int, data_bytes: int, device=None):
        from . import _abi as abi
        self._abi, self.lib = abi, abi.load()
        self.rank, self.world = rank, world
        self.flag_bytes = int(self.lib.b200_peer_flag_bytes())
        self.bytes = self.flag_bytes + (data_bytes + 255) // 256 * 256
        self.region = C.c_void_p()
        abi.check(self.lib.b200_peer_alloc(C.byref(self.region), self.bytes))
        raw = (C.c_uint8 * 64)()
        abi.check(self.lib.b200_peer_export(self.region, raw))
        handles = gather_bytes(bytes(raw), rank, world, device) if world > 1 else bytes(raw)
        self.handle = C.c_void_p()
        arr = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        abi.check(self.lib.b200_peer_group_create(C.byref(self.handle), rank, world, self.region, self.bytes, arr))
        self.data_ptr = int(self.lib.b200_peer_data(self.handle))
        self.cursor = 0          # bump allocator over the data area (elements); identical on every rank
        self.next_slot = 0

    def carve(self, numel: int, dtype=None):
        """A contiguous f32 tensor of `numel` elements in the data area; returns (DeviceTensor, element offset)."""
        from .device import DeviceTensor
        abi = self._abi
        off = self.cursor
        n = (numel + 3) // 4 * 4
        if (off + n) * 4 + self.flag_bytes > self.bytes:
            raise MemoryError("peer region exhausted")
        self.cursor += n
        st = PeerRegion(self.data_ptr + off * 4, n * 4, self)
        return DeviceTensor(st, abi.F32, (numel,)), off

    def slot(self) -> int:
        s = self.next_slot
        self.next_slot += 1
        return s

    def all_reduce(self, offset: int, count: int, slot: int, mean: bool = True, producer_stream=None) -> None:
        abi = self._abi
        abi.check(self.lib.b200_launch_peer_all_reduce(self.handle, offset, count,
                                                       abi.REDUCE_MEAN if mean else abi.REDUCE_SUM, slot, producer_stream))

    def adam(self, g_off: int, p_off: int, m, v, coef, count: int, lr: float, b1: float, b2: float, slot: int,
             producer_stream=None) -> None:
        self._abi.check(self.lib.b200_launch_peer_adam(self.handle, g_off, p_off, m.data_ptr(), v.data_ptr(), coef.data_ptr(),
                                                       count, lr, b1, b2, slot, producer_stream))

    def sync(self, consumer_stream=None) -> None:
        self._abi.check(self.lib.b200_peer_sync(self.handle, consumer_stream))

    def mark(self, event) -> None:
        self._abi.check(self.lib.b200_peer_mark(self.handle, event))

    def close(self) -> None:
        """Every rank must have stopped using the group (host barrier) before anyone closes it."""
        if self.handle:
            self.lib.b200_peer_group_destroy(self.handle)
            self.handle = C.c_void_p()
        if self.region:
            self.lib.b200_peer_free(self.region)
            self.region = C.c_void_p()
