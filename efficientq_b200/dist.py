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
    def __init__(self, group=None, single: bool = False):
        """``single=True``: a one-process context even when a process group is initialised (e.g. rank 0 repeating a
        job unsharded for comparison) -- no collective is ever issued through it."""
        self.enabled = td.is_available() and td.is_initialized() and not single
        self.group = group
        self.world = td.get_world_size(group) if self.enabled else 1
        self.rank = td.get_rank(group) if self.enabled else 0
        self._peer = None

    def peer_link(self, device: torch.device):
        """The NVLink peer link of this run (created on first use; None for a single process, for
        non-CUDA process groups, with EFFQ_PEER=0, or when CUDA IPC is unavailable on ANY rank: the
        decision is collective -- a rank that fell back to NCCL alone would wait in a different
        collective than its peers)."""
        if self._peer is None and self.world > 1 and device.type == "cuda" and \
                os.environ.get("EFFQ_PEER", "1") != "0" and td.get_backend(self.group) == "nccl":
            link = PeerLink(self, device)
            if link.ok:
                self._peer = link
            else:
                import warnings
                warnings.warn(f"NVLink peer link unavailable ({link.error!r} on this rank or a failure on a peer); "
                              "using NCCL all-reduces")
                link.close()
                self._peer = False
        return self._peer or None

    def close(self) -> None:
        """Release the peer mappings (end of do_ptq)."""
        if self._peer:
            self._peer.close()
        self._peer = None

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

    def all_gather_object(self, obj) -> list:
        """One small picklable object from every rank, in rank order (plumbing for per-subject metric rows)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        td.all_gather_object(out, obj, group=self.group)
        return out

    def shard(self, n_total: int) -> Tuple[int, int]:
        """[begin, end) of the calibration volumes owned by this rank (contiguous, balanced)."""
        return shard_range(n_total, self.rank, self.world)


class PeerLink:
    """NVLink peer-memory link between the ranks of one node for the in-kernel exchanges of
    csrc/peer.cuh (2 doubles per activation-search pass, 1 double per ADMM iteration: thousands of
    latency-bound all-reduces per layer that are cheaper inside the kernels than as NCCL calls).

    Every rank allocates one 4 KB slot buffer, exports it through CUDA IPC, all-gathers the 64-byte
    handles over the process group and maps its peers' buffers.  ``comm_ptr`` is what the C-ABI
    takes as ``const effq_peer_comm*``.  Construction fails (EffqError) when IPC / peer access is
    not available; the caller then keeps the NCCL form of the same exchanges."""

    def __init__(self, ctx: "DistCtx", device: torch.device):
        import ctypes as C
        from . import capi
        lib = capi.load()
        self._lib = lib
        self.local = C.c_void_p()
        self._opened = []
        self.error = None
        self.comm = capi.PeerComm()
        self.comm.rank, self.comm.world = ctx.rank, ctx.world
        # Every step that can fail locally is followed by the SAME collectives on every rank; `ok` is the
        # all-reduced (MIN) success flag, so either all ranks use the link or none does.
        handle = C.create_string_buffer(64)
        try:
            if ctx.world > 8:
                raise capi.EffqError("PeerLink: at most 8 ranks (one node)")
            capi.check(lib.effq_peer_alloc(C.byref(self.local), handle), "effq_peer_alloc")
        except Exception as exc:  # noqa: BLE001
            self.error = exc
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(device)
        gathered = [torch.empty_like(mine) for _ in range(ctx.world)]
        td.all_gather(gathered, mine, group=ctx.group)
        flag = torch.tensor([0.0 if self.error else 1.0], device=device)
        td.all_reduce(flag, op=td.ReduceOp.MIN, group=ctx.group)
        if float(flag.item()) == 1.0:
            try:
                for r, h in enumerate(gathered):
                    if r == ctx.rank:
                        self.comm.slots[r] = self.local.value
                        continue
                    p = C.c_void_p()
                    capi.check(lib.effq_peer_open(bytes(h.cpu().numpy().tobytes()), C.byref(p)), "effq_peer_open")
                    self._opened.append(p)
                    self.comm.slots[r] = p.value
            except Exception as exc:  # noqa: BLE001
                self.error = exc
        flag = torch.tensor([0.0 if self.error else 1.0], device=device)
        td.all_reduce(flag, op=td.ReduceOp.MIN, group=ctx.group)     # also the "everyone has mapped everything" barrier
        self.ok = float(flag.item()) == 1.0
        self.comm_ptr = C.byref(self.comm)

    def close(self) -> None:
        for p in self._opened:
            self._lib.effq_peer_close(p)
        self._opened = []
        if self.local:
            self._lib.effq_peer_free(self.local)
            self.local = None


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
        kw = {}
        if os.environ.get("EFFQ_DIST_TIMEOUT_S"):               # tests: fail fast on a mismatched collective
            import datetime
            kw["timeout"] = datetime.timedelta(seconds=int(os.environ["EFFQ_DIST_TIMEOUT_S"]))
        td.init_process_group(backend=backend, **kw)
    return DistCtx()
