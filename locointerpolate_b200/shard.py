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
