"""Data-parallel plumbing for the photometric engine: frames of a batch are sharded across ranks (one process per GPU,
torch.distributed over NCCL/NVLink); per Adam step there is
  * one tiny reduction of the forward slab between the forward and backward halves (the loss normaliser
    sum(alpha_aa > 0) and the reg_diffuse mean / max are batch-global, tracker.py:439,548-549), and
  * one sum-allreduce of the gradient slab (shared parameters + dense per-frame rows) and of the dense texture gradient.
The reference has no distributed code; this is new.  Everything here is backend-agnostic so the host
logic is testable on CPU with gloo (tests/test_parallel_gloo.py)."""
from __future__ import annotations

import torch
import torch.distributed as dist

SLAB_SUM = slice(0, 3)      # [0] sum|err|  [1] n(alpha_aa>0)  [2] sum var(diffuse)
SLAB_MAX = slice(3, 4)      # [3] max diffuse
# [4] arg-max index, [5] local max, [6] n background pixels stay rank-local


def reduce_forward_slab(local: torch.Tensor, out: torch.Tensor, group=None) -> None:
    """out <- cross-rank reduction of the forward slab produced by vhap_energy_forward (layout above)."""
    out.copy_(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    # ONE latency-bound collective in the middle of the step instead of a sum- and a max-all-reduce: gather the four scalars
    # of every rank and reduce them locally
    world = dist.get_world_size(group)
    flat = torch.empty(world * 4, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, local[0:4].contiguous(), group=group)
    gathered = flat.view(world, 4)
    out[SLAB_SUM] = gathered[:, 0:3].sum(0)
    out[SLAB_MAX] = gathered[:, 3:4].max(0).values


def allreduce_sum(t: torch.Tensor, group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def shard_frames(n_frames: int, rank: int, world: int):
    """Contiguous frame shard of rank `rank` (frames are independent units of the path)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


class DataParallelStep:
    """Wraps an Engine: step(batch) = zero_grad, forward, slab reduce, backward, grad allreduce, Adam."""

    def __init__(self, engine, group=None):
        self.e, self.group = engine, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    def body(self, batch):
        """the step without host-side counters: what gets captured into a CUDA graph (Engine.graph_begin(body=...))"""
        e = self.e
        e.zero_grad()
        gB = batch.B * self.world
        red = (lambda a, b: reduce_forward_slab(a, b, self.group)) if self.world > 1 else None
        e.energy(batch, backward=True, training=True, global_B=gB, reduce_fn=red)
        e.adam_step(allreduce_fn=(lambda t: allreduce_sum(t, self.group)) if self.world > 1 else None)

    def graph_begin(self, batches, pipelined=True):
        """Engine.graph_begin with this group's collectives captured into the step graphs (pipelined: the texture all-reduce of step k
        runs at the start of step k+1, hidden behind FLAME / rasteriser / pools)"""
        red = (lambda a, b: reduce_forward_slab(a, b, self.group)) if self.world > 1 else None
        allr = (lambda t: allreduce_sum(t, self.group)) if self.world > 1 else None
        self.e.graph_begin(batches, reduce_fn=red, allreduce_fn=allr, world=self.world, pipelined=pipelined)

    def step(self, batch):
        e = self.e
        e.zero_grad()
        gB = batch.B * self.world
        red = (lambda a, b: reduce_forward_slab(a, b, self.group)) if self.world > 1 else None
        losses = e.energy(batch, backward=True, training=True, global_B=gB, reduce_fn=red)
        e.adam_step(allreduce_fn=(lambda t: allreduce_sum(t, self.group)) if self.world > 1 else None)
        e.global_step += 1
        return losses
