"""Batch sharding across the GPUs of one box + the two collectives the path needs.

Every image's forward / CAM is independent given the per-rank batch composition, so the image range is cut into
contiguous, balanced shards (one process per GPU, `torchrun`); NCCL is used only to gather the per-rank CAM / label
buffers and to sum the metric counters -- never inside the forward.  The reference has no inference sharding at all
(validate.py:97-102 is single-process, batch 1) and its `reduce_value` helper (distributed_utils.py:60-70) is never
called; this module is the role those helpers were meant to play.  Works with the `gloo` backend on CPU tensors too
(that is how the host logic is tested without GPUs)."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torchrun).  Returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced shard [lo, hi): the first n % world ranks get one extra item (10,582 images over 8 ranks ->
    1323,1323,1323,1323,1323,1323,1322,1322)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def batches(lo: int, hi: int, batch: int) -> List[Tuple[int, int]]:
    """Balanced batches of a shard: ceil(n / batch) batches whose sizes differ by at most one (a 1,323-image shard at
    batch 256 runs 6 x ~221 instead of 5 x 256 + 43)."""
    n = hi - lo
    if n <= 0:
        return []
    k = -(-n // batch)
    out, start = [], lo
    for i in range(k):
        size = n // k + (1 if i < n % k else 0)
        out.append((start, start + size))
        start += size
    return out


def gather_shards(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """All-gather per-rank results of a `shard_range` partition back into global image order: local [n_r, ...] ->
    [n_items, ...] on every rank.  Shards are padded to equal length for the collective and the padding is dropped."""
    rank, world = rank_world()
    if world == 1:
        return local
    sizes = shard_sizes(n_items, world)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    pad = max(sizes)
    buf = local
    if local.shape[0] < pad:
        buf = torch.cat([local, local.new_zeros((pad - local.shape[0], *local.shape[1:]))])
    out = local.new_empty((world * pad, *local.shape[1:]))       # concatenated form: accepted by both NCCL and gloo
    dist.all_gather_into_tensor(out, buf.contiguous())
    out = out.view(world, pad, *local.shape[1:])
    return torch.cat([out[r, :sizes[r]] for r in range(world)])


class SideStreamGather:
    """All-gather of per-step results on a side stream, so the collective overlaps the forward of the following steps
    (collectives outside the forward, on a side stream, once per k batches).

    every == 1: `gather(out, local)` enqueues all_gather_into_tensor(out, local) behind everything already enqueued on the
    caller's stream; out is [world * B, ...].
    every == k > 1: the results of k consecutive steps are staged in a device buffer (one 4 MB device-to-device copy per
    step at the bench shape) and leave in ONE collective: out, with room for world * k * B rows, then holds
    [world, k, B, ...].  The NCCL kernel needs SMs, which the persistent kernels of the forward occupy completely, and it
    spins while its peers catch up -- so every collective both waits for a gap between two kernels and holds SMs the next
    kernel's CTAs want; issuing it every k-th step divides that cost by k (measured at 2 GPUs: 12.06 -> see DESIGN.md).
    `wait()` flushes what is staged (a shorter [world, n, B, ...] in the leading elements of the last `out`) and makes the
    caller's stream wait for all collectives issued so far (call it before reading `out`)."""

    def __init__(self, device, every: int = 1):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.every = max(int(every), 1)
        self._stage = [None, None]          # two staging buffers: one fills while the other is being sent
        self._sent = [None, None]           # event: the collective reading staging buffer i has finished
        self._flip = 0
        self._k = 0
        self._out = None

    def gather(self, out: torch.Tensor, local: torch.Tensor) -> None:
        world = rank_world()[1]
        if self.every == 1:
            if world == 1:
                out.copy_(local)
                return
            cur = torch.cuda.current_stream(self.device)
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(out, local)
            local.record_stream(self.stream)       # `local` may be freed by the caller while the collective still reads it
            return
        assert out.numel() >= world * self.every * local.numel() and out.is_contiguous(), "out must hold world * every * local elements"
        cur = torch.cuda.current_stream(self.device)
        st = self._stage[self._flip]
        if st is None or st.shape[1:] != local.shape or st.dtype != local.dtype:
            st = self._stage[self._flip] = torch.empty((self.every,) + tuple(local.shape), dtype=local.dtype, device=self.device)
        if self._k == 0 and self._sent[self._flip] is not None:
            cur.wait_event(self._sent[self._flip])          # the previous collective out of this staging buffer has read it
        st[self._k].copy_(local)
        self._k += 1
        self._out = out
        if self._k == self.every:
            self._flush()

    def _flush(self) -> None:
        n = self._k
        if n == 0:
            return
        world = rank_world()[1]
        st = self._stage[self._flip][:n]
        dst = self._out.view(-1)[: world * st.numel()].view((world,) + tuple(st.shape))
        cur = torch.cuda.current_stream(self.device)
        if world == 1:
            dst[0].copy_(st)
        else:
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(dst, st)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._sent[self._flip] = ev
        self._flip ^= 1
        self._k = 0

    def wait(self) -> None:
        if self.every > 1:
            self._flush()
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


class PeerPushGather:
    """The same gather without a collective kernel: every rank WRITES its staged block straight into the result buffer of
    every peer with plain device-to-device copies over NVLink (copy engines; no SM is taken from the persistent kernels of
    the forward, which is what made the NCCL all-gather cost 7 % at 2 GPUs).  The result buffers are symmetric memory
    (`torch.distributed._symmetric_memory`: one allocation per rank, mapped into every peer of the box), allocated by
    `alloc()`.  Interface of SideStreamGather: `gather(out, local)` stages one step, every `every`-th call pushes the staged
    steps on a side stream; `wait()` pushes what is left and makes the caller's stream wait for this rank's copies.  The
    peers' copies into MY buffer are ordered by the next collective of the caller (bench.py: the all-reduce of the counters,
    which every rank enqueues after its own `wait()`): when that collective has completed here, every peer had reached it,
    i.e. had finished its pushes.  Raises at construction if symmetric memory cannot be set up (the caller then uses
    SideStreamGather)."""

    def __init__(self, device, every: int = 1, group=None):
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.every = max(int(every), 1)
        self.stream = torch.cuda.Stream(device=self.device)
        self._stage = [None, None]
        self._sent = [None, None]
        self._flip = 0
        self._k = 0
        self._peers = None          # views of every peer's result buffer
        self._out = None

    def alloc(self, shape, dtype=torch.float32) -> torch.Tensor:
        """Result buffer [world, every, ...] in symmetric memory; collective (every rank calls it with the same shape)."""
        out = self._symm.empty(tuple(shape), dtype=dtype, device=self.device)
        hdl = self._symm.rendezvous(out, self.group)
        self._peers = [out if r == self.rank else hdl.get_buffer(r, tuple(shape), dtype) for r in range(self.world)]
        self._hdl = hdl
        self._out = out
        return out

    def gather(self, out: torch.Tensor, local: torch.Tensor) -> None:
        assert self._out is not None and out.data_ptr() == self._out.data_ptr(), "out must come from alloc()"
        cur = torch.cuda.current_stream(self.device)
        st = self._stage[self._flip]
        if st is None or st.shape[1:] != local.shape or st.dtype != local.dtype:
            st = self._stage[self._flip] = torch.empty((self.every,) + tuple(local.shape), dtype=local.dtype, device=self.device)
        if self._k == 0 and self._sent[self._flip] is not None:
            cur.wait_event(self._sent[self._flip])
        st[self._k].copy_(local)
        self._k += 1
        if self._k == self.every:
            self._flush()

    def _flush(self) -> None:
        n = self._k
        if n == 0:
            return
        st = self._stage[self._flip][:n]
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for r in range(self.world):
                dst = self._peers[r].view(self.world, -1)[self.rank, : st.numel()]
                dst.copy_(st.reshape(-1), non_blocking=True)          # contiguous same-dtype copy: cudaMemcpyAsync, peer-mapped destination
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._sent[self._flip] = ev
        self._flip ^= 1
        self._k = 0

    def wait(self) -> None:
        self._flush()
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


def reduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """Sum int64 counters (21x21 confusion matrix, AP sum / count, top-1 hits ...) over all ranks, in place."""
    if rank_world()[1] > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if rank_world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
