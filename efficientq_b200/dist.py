"""Data-parallel plumbing: one process per GPU, calibration volumes sharded by rank.

The reference is single-process.  Layers are sequential and
samples are independent inside a layer, so rank r keeps N/G volumes (and their FP
targets and activations) resident for the whole run and only these sums cross
NVLink, all as NCCL all-reduce(SUM):

  * sum(b*x), sum(b*b) per fixed-point pass of the activation scale search
    (reference layer_helper.py:51,59)                               -- 2 doubles
  * sum(y), sum(y^2), n  and  sum(att), n  for rho_scale
    (EfficientQConv.py:45-49,61)                                    -- 5 doubles
  * A0 || B0 partial normal-equation sums (solver.py:305-312)       -- once per layer
  * the per-iteration squared error (EfficientQConv.py:121)         -- 1 double x 200
  * class voxel counts for the attention map (ptqer.py:172-188)     -- once per run

Everything else (factorisation, weight projection, dual update) is replicated: the
inputs are identical after the all-reduce and the kernels are deterministic, so the
ranks stay in lock-step without further exchange.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as td


class DistCtx:
    def __init__(self, group=None):
        self.enabled = td.is_available() and td.is_initialized()
        self.group = group
        self.world = td.get_world_size(group) if self.enabled else 1
        self.rank = td.get_rank(group) if self.enabled else 0

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        """In-place SUM all-reduce (no-op for a single process); returns ``t``."""
        if self.world > 1:
            td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self) -> None:
        if self.world > 1:
            td.barrier(group=self.group)

    def shard(self, n_total: int) -> Tuple[int, int]:
        """[begin, end) of the calibration volumes owned by this rank (contiguous, balanced)."""
        return shard_range(n_total, self.rank, self.world)


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend: str = "nccl") -> DistCtx:
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not td.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group(backend=backend)
    return DistCtx()
