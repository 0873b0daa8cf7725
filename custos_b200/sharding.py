"""Multi-GPU layout of the hot path: one process per GPU, contiguous slices, no collective on
the element-wise data path; only reduction partials are exchanged.

The reference is single-device (its one cross-device test is ignored: src/devices/cuda/cuda.rs:187-202),
so this is new functionality with a deliberately small contract:

  * `shard_bounds(n, elem_bytes, world, rank)` — the slice [begin, end) a rank owns; slice starts are
    multiples of 16 bytes so every rank runs the 128-bit kernels (cb_shard_range in the C ABI);
  * a sharded sum = each rank's deterministic two-pass partial (cb_sum) -> all-gather of ONE scalar per
    rank -> fold in RANK ORDER on every rank.  All ranks end with identical bits and the order never
    depends on arrival time.  `ShardedReducer` does the exchange through cb_comm (NCCL) when it has a
    communicator, or through a torch.distributed process group (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from .raw import shard_range


def shard_bounds(n: int, elem_bytes: int, world: int, rank: int) -> tuple[int, int]:
    return shard_range(n, elem_bytes, world, rank)


def fold_rank_order(partials: Sequence, acc_dtype) -> np.generic:
    """p0 + p1 + ... + p_{R-1}, left to right, in the accumulation dtype (f32 for f32/f16 data).
    The device does the same in fold_ranks_kernel (csrc/kernels.cu)."""
    acc = acc_dtype(partials[0])
    for p in partials[1:]:
        acc = acc_dtype(acc + acc_dtype(p))
    return acc


class ShardedReducer:
    """Rank-ordered combine of per-rank partials over a torch.distributed process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def bounds(self, n: int, elem_bytes: int) -> tuple[int, int]:
        return shard_bounds(n, elem_bytes, self.world, self.rank)

    def sum(self, local_partial, acc_dtype=np.float32, device: Optional[str] = None):
        """`local_partial`: this rank's partial (a Python/NumPy scalar).  Returns the global sum."""
        import torch
        tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int64: torch.int64}[acc_dtype]
        mine = torch.tensor([local_partial], dtype=tdt, device=device or "cpu")
        gathered = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(gathered, mine, group=self.group)
        return fold_rank_order([g.cpu().numpy()[0] for g in gathered], acc_dtype)

    def mean(self, local_partial, n_global: int, acc_dtype=np.float32, device: Optional[str] = None):
        total = self.sum(local_partial, acc_dtype, device)
        if acc_dtype is np.int64:
            return np.int64(int(total) // n_global)
        return acc_dtype(total / acc_dtype(n_global))


def sharded_apply(world: int, rank: int, n: int, elem_bytes: int, run_slice: Callable[[int, int], None]) -> tuple[int, int]:
    """Element-wise work needs no communication: run the caller's kernel on this rank's slice."""
    begin, end = shard_bounds(n, elem_bytes, world, rank)
    if end > begin:
        run_slice(begin, end)
    return begin, end
