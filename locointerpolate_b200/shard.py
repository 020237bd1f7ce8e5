"""Query-batch sharding over the GPUs of one box: one process per GPU, the flat model replicated,
the batch split into contiguous ranges, and ONE exchange step at the end --

  * scalar results (8 B/query), leaf indices and status bytes: all_gather of the per-shard slices;
  * error check: all_reduce(MAX) of the per-leaf error maxima and of the split flags.

The maxima are reduced as the int64 BIT PATTERNS of the non-negative doubles, so that the integer order is the
numeric order and a NaN (bit pattern above every finite value) wins, as decideSplit's `!(max(diff - thr) <= 0)` demands
(LCAdaptiveGrid.cpp:598) -- the same convention as the device-side atomicMax (loco_kernels.cu error_reduce_kernel).
Mesh-valued results (240 KB - 1.2 MB per query) never cross GPUs; only their per-query checksums do.

`torch.distributed` is plumbing: NCCL over NVLink on the GPU box, gloo in the CPU tests.  The evaluator is a
callable so that the partition/exchange logic can be exercised without a GPU (tests/test_shard.py).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of rank `rank` among `world` ranks: the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _gather_ragged(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of contiguous slices whose lengths follow shard_range (differ by at most one)."""
    rank, world = _world(group)
    if world == 1:
        return local
    longest = shard_range(n_total, 0, world)
    width = longest[1] - longest[0]
    tail = local.shape[1:]
    padded = torch.zeros((width, *tail), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width, *tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        b, e = shard_range(n_total, r, world)
        parts.append(out[r * width: r * width + (e - b)])
    return torch.cat(parts)


def sharded_eval(evaluate: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], q: torch.Tensor, group=None):
    """Every rank holds the full query batch q [Q, N]; rank r evaluates its contiguous shard with `evaluate`
    (-> out [n] or [n, D], leaf [n] int32, status [n] uint8) and the shards are all-gathered.
    Returns (out, leaf, status) for the whole batch on every rank."""
    rank, world = _world(group)
    Q = q.shape[0]
    b, e = shard_range(Q, rank, world)
    out, leaf, status = evaluate(q[b:e])
    return _gather_ragged(out, Q, group), _gather_ragged(leaf, Q, group), _gather_ragged(status, Q, group)


def reduce_error_maxima(err_max: torch.Tensor, split: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Combine per-rank per-leaf error maxima (float64, >= 0 or NaN) and split flags (uint8) over all ranks."""
    rank, world = _world(group)
    if world == 1:
        return err_max, split
    bits = err_max.contiguous().view(torch.int64).clone()
    dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
    flags = split.to(torch.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return bits.view(torch.float64), flags.to(torch.uint8)


def sharded_error_check(check: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], q: torch.Tensor,
                        truth: torch.Tensor, group=None):
    """decideSplit's data-parallel part over a sharded candidate batch: rank r checks its shard with `check`
    (-> err_max [L] float64, split [L] uint8) and the per-leaf results are max/or-reduced."""
    rank, world = _world(group)
    b, e = shard_range(q.shape[0], rank, world)
    err, split = check(q[b:e], truth[b:e])
    return reduce_error_maxima(err, split, group)


class ShardedModel:
    """A replica of the flat model on this rank's GPU plus the exchange step.  q/truth are CUDA tensors."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    def _eval_local(self, q: torch.Tensor):
        n, D = q.shape[0], self.model.D
        out = torch.empty((n,) if D == 1 else (n, D), dtype=torch.float64, device=q.device)
        leaf = torch.empty(n, dtype=torch.int32, device=q.device)
        status = torch.empty(n, dtype=torch.uint8, device=q.device)
        if n:
            self.model.eval_device(q.contiguous().data_ptr(), n, out.data_ptr(), leaf.data_ptr(), status.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
        return out, leaf, status

    def eval(self, q: torch.Tensor):
        return sharded_eval(self._eval_local, q, self.group)

    def _check_local(self, threshold: float):
        def run(q: torch.Tensor, truth: torch.Tensor):
            L = self.model.L
            err = torch.zeros(L, dtype=torch.float64, device=q.device)
            split = torch.zeros(L, dtype=torch.uint8, device=q.device)
            self.model.error_check_device(q.contiguous().data_ptr(), truth.contiguous().data_ptr(), q.shape[0], threshold, err.data_ptr(),
                                          split.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
            return err, split
        return run

    def error_check(self, q: torch.Tensor, truth: torch.Tensor, threshold: float):
        return sharded_error_check(self._check_local(threshold), q, truth, self.group)


# ------------------------------------------------------------------------------------------------------------------
# Streaming exchange for scalar results: batch b (one shard per rank) is consumed on rank b % world
# ------------------------------------------------------------------------------------------------------------------
# measurement switches (bench experiments only: they break the exchange's guarantees)
import os as _os
_EXPERIMENT_NO_COPY = bool(_os.environ.get("LOCO_EX_NOCOPY"))
_EXPERIMENT_NO_SIGNAL = bool(_os.environ.get("LOCO_EX_NOSIGNAL"))


class _ArrayView:
    """CUDA array interface over a raw device pointer (a peer-mappable buffer from the C ABI), for torch.as_tensor"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class RotatingExchange:
    """The path's one exchange step for a STREAM of batches: every rank evaluates its shard of every batch,
    the shards of batch b are gathered on rank b % world (the consumer of that batch).

    Why rotating consumers and a skewed schedule.  One GPU produces 8 B x ~27 G results/s = 216 GB/s; eight produce 1.7 TB/s,
    twice what ONE GPU's NVLink ingress takes, so gathering every batch on one GPU cannot scale, and even with rotating
    consumers all ranks would push batch b to the same GPU at the same moment.  Rank r therefore ships batch b at step
    b + r: in every step the `world` transfers form a PERMUTATION (rank r -> rank (s - r) % world), each GPU sends one shard
    and receives one shard per step, and the ingress load is the GPU's own production rate.

    Transport "ipc" (GPU): the consumer's gather slots are device buffers exported through the C ABI
    (loco_peer_buffer_create / _open, CUDA IPC); a shard travels as ONE asynchronous peer copy on a side stream
    (loco_peer_copy_async): copy engines over NVLink, no SM kernel competes with the evaluation.  A 4-byte all_reduce per step
    on the same side stream is the completion signal: when the signal of step s has passed on a rank, every rank's copy of
    step s has landed.  Transport "p2p" (fallback, and the CPU/gloo tests): the same permutation as batch_isend_irecv.

    Buffers: world + 1 local result slots (a shard waits at most world - 1 steps for its turn), 2 gather slots per rank
    (a rank consumes every world-th batch).  Usage per batch b, in order on the evaluation stream:
        out = ex.out_buffer(b); <evaluate shard into out>; ex.submit(b)      ...      ex.drain(); ex.gathered(b) on rank b % world
    """

    def __init__(self, shard_elems: int, dtype=torch.float64, device="cpu", group=None, transport: str = "auto"):
        self.rank, self.world = _world(group)
        self.group, self.n, self.dtype = group, int(shard_elems), dtype
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.itemsize = torch.empty(0, dtype=dtype).element_size()
        self.NB = self.world + 1
        self.local = [torch.empty(self.n, dtype=dtype, device=self.device) for _ in range(self.NB)]
        self.transport = transport
        if transport == "auto":
            self.transport = "ipc" if self.cuda and self.world > 1 else "p2p"
        self._peer_ptr = None        # [rank] -> device pointer of that rank's gather area, mapped into this process
        self._own_ptr = None
        self._opened = []
        slot_elems = self.world * self.n
        if self.transport == "ipc":
            try:
                self._setup_ipc(2 * slot_elems)
            except Exception as e:   # CUDA IPC unavailable in this container: same schedule over NCCL point-to-point
                self._ipc_error = f"{type(e).__name__}: {e}"[:200]
                self.transport = "p2p"
            ok = torch.tensor([1 if self.transport == "ipc" else 0], device=self.device)
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # all ranks or none
            if int(ok.item()) == 0:
                self._teardown_ipc()
                self.transport = "p2p"
        if self.transport != "ipc":
            self._area = torch.empty(2 * slot_elems, dtype=dtype, device=self.device)
        self.slots = [self._area[k * slot_elems:(k + 1) * slot_elems].view(self.world, self.n) for k in range(2)]
        if self.cuda:
            self.copy_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._ev_done = [torch.cuda.Event() for _ in range(self.NB)]     # evaluation of the batch in local slot k enqueued/finished
            self._ev_sent = [None] * self.NB                                  # its copy has left the slot
            self._ev_step = {}                                                # step -> event after that step's signal
        self._sig = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._next_step = 0
        self._last_batch = -1
        self.time_copies = False
        self.copy_events = []   # (start, end, bytes) of timed peer copies while self.time_copies

    # ---- transports
    def _setup_ipc(self, area_elems: int):
        from . import capi
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._dev = dev
        nbytes = area_elems * self.itemsize
        self._own_ptr, handle = capi.peer_buffer_create(dev, nbytes)
        typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[self.dtype]
        self._area = torch.as_tensor(_ArrayView(self._own_ptr, area_elems, typestr), device=self.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        self._peer_ptr = [None] * self.world
        for r in range(self.world):
            if r == self.rank:
                self._peer_ptr[r] = self._own_ptr
            else:
                self._peer_ptr[r] = capi.peer_buffer_open(dev, handles[r])
                self._opened.append(self._peer_ptr[r])

    def _teardown_ipc(self):
        if self._own_ptr is None:
            return
        from . import capi
        for p in self._opened:
            try:
                capi.peer_buffer_close(self._dev, p)
            except Exception:
                pass
        self._opened = []
        self._area = None
        self.slots = None
        try:
            capi.peer_buffer_free(self._dev, self._own_ptr)
        except Exception:
            pass
        self._own_ptr = None

    def close(self):
        if self.cuda:
            torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)   # nobody unmaps / frees while a peer may still write
        self._teardown_ipc()

    # ---- schedule
    def consumer(self, b: int) -> int:
        return b % self.world

    def _slot_row(self, b: int, r: int) -> torch.Tensor:
        return self.slots[(b // self.world) % 2][r]

    def out_buffer(self, b: int) -> torch.Tensor:
        """where the evaluation of batch b must write this rank's shard (call on the evaluation stream)"""
        if self.consumer(b) == self.rank:
            # straight into the own row of the gather slot: the slot's previous batch (b - 2*world) was complete -- and its
            # consumer's reads were enqueued -- before the signal of step b - world - 1
            if self.cuda and (b - self.world - 1) in self._ev_step:
                torch.cuda.current_stream(self.device).wait_event(self._ev_step[b - self.world - 1])
            return self._slot_row(b, self.rank)
        k = b % self.NB
        if self.cuda and self._ev_sent[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._ev_sent[k])   # the slot's previous shard has been shipped
        return self.local[k]

    def _ship(self, s: int):
        """the transfers of step s: this rank sends batch s - rank (if it exists), and receives from rank (s - rank) % world"""
        bd = s - self.rank
        send = 0 <= bd <= self._last_batch and self.consumer(bd) != self.rank
        if self.transport == "ipc":
            if send:
                from . import capi
                c, k = self.consumer(bd), bd % self.NB
                dst = self._peer_ptr[c] + (((bd // self.world) % 2) * self.world + self.rank) * self.n * self.itemsize
                self.copy_stream.wait_event(self._ev_done[k])
                with torch.cuda.stream(self.copy_stream):
                    t0 = t1 = None
                    if self.time_copies:
                        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        t0.record()
                    if not _EXPERIMENT_NO_COPY:
                        capi.peer_copy_async(dst, self.local[k].data_ptr(), self.n * self.itemsize, self.copy_stream.cuda_stream)
                    if t0 is not None:
                        t1.record()
                        self.copy_events.append((t0, t1, self.n * self.itemsize))
                    self._ev_sent[k] = torch.cuda.Event()
                    self._ev_sent[k].record()
        else:
            ops = []
            if send:
                k = bd % self.NB
                if self.cuda:
                    self.copy_stream.wait_event(self._ev_done[k])
                ops.append(dist.P2POp(dist.isend, self.local[k], self.consumer(bd), group=self.group))
            src = (s - self.rank) % self.world
            bs = s - src   # the batch rank `src` ships in step s
            if src != self.rank and 0 <= bs <= self._last_batch:
                ops.append(dist.P2POp(dist.irecv, self._slot_row(bs, src), src, group=self.group))
            if ops:
                if self.cuda:
                    with torch.cuda.stream(self.copy_stream):
                        for w in dist.batch_isend_irecv(ops):
                            w.wait()   # stream-ordered on the side stream, the host does not block
                        if send:
                            self._ev_sent[bd % self.NB] = torch.cuda.Event()
                            self._ev_sent[bd % self.NB].record()
                else:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()

    def _signal(self, s: int):
        if self.world == 1:
            return
        if self.cuda:
            with torch.cuda.stream(self.copy_stream):
                if self.transport == "ipc" and not _EXPERIMENT_NO_SIGNAL:
                    dist.all_reduce(self._sig, group=self.group)   # stream-ordered: passes when every rank's copies of step s landed
                ev = torch.cuda.Event()
                ev.record()
            self._ev_step[s] = ev
            self._ev_step.pop(s - 2 * self.world - 2, None)
        # p2p: the receive itself is the completion (its work item was waited on in stream order)

    def submit(self, b: int):
        """the evaluation of batch b (into out_buffer(b)) has been enqueued on the current stream"""
        assert b == self._last_batch + 1 == self._next_step, "batches must be submitted in order, one per step"
        self._last_batch = b
        if self.cuda and self.consumer(b) != self.rank:
            self._ev_done[b % self.NB].record(torch.cuda.current_stream(self.device))
        elif self.cuda:
            # own shard written in place: later transfers into the same slot are ordered behind it by the step signals; the
            # side stream must still not signal completion of this batch before the evaluation finished
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.copy_stream.wait_event(ev)
        if self.world > 1:
            self._ship(b)
            self._signal(b)
        self._next_step = b + 1

    def drain(self):
        """ship what the skew still holds back (steps last+1 .. last+world-1) and make the current stream wait for everything"""
        if self.world > 1:
            for s in range(self._next_step, self._last_batch + self.world):
                self._ship(s)
                self._signal(s)
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
            self._ev_step.clear()
        # the epoch is over (every shard of every submitted batch has landed): batch numbering restarts at 0
        self._next_step, self._last_batch = 0, -1

    def warm(self):
        """touch every peer once (first-use connection / mapping set-up must not fall into a timed region)"""
        if self.world == 1:
            return
        if self.transport == "ipc":
            from . import capi
            with torch.cuda.stream(self.copy_stream):
                for r in range(self.world):
                    if r != self.rank:
                        nb = min(self.n, 1024) * self.itemsize
                        capi.peer_copy_async(self._peer_ptr[r] + self.rank * self.n * self.itemsize, self.local[0].data_ptr(), nb, self.copy_stream.cuda_stream)
                dist.all_reduce(self._sig, group=self.group)
        else:
            tiny_s = torch.zeros(8, dtype=self.dtype, device=self.device)
            tiny_r = torch.zeros(8, dtype=self.dtype, device=self.device)
            for shift in range(1, self.world):
                ops = [dist.P2POp(dist.isend, tiny_s, (self.rank + shift) % self.world, group=self.group),
                       dist.P2POp(dist.irecv, tiny_r, (self.rank - shift) % self.world, group=self.group)]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        if self.cuda:
            torch.cuda.synchronize(self.device)

    def gathered(self, b: int) -> torch.Tensor:
        """[world, shard_elems] results of batch b on its consumer (valid after drain(), or once step b + world - 1 was signalled)"""
        assert self.consumer(b) == self.rank
        return self.slots[(b // self.world) % 2]
