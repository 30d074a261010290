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
    """Contiguous frame shard of rank `rank` (frames are independent units of the path): the first n_frames % world
    ranks hold one frame more.  Every rank needs at least one frame (a rank with an empty batch has nothing to launch)."""
    if n_frames < world:
        raise ValueError(f"{n_frames} frames cannot be sharded over {world} ranks: every rank needs at least one frame")
    per, extra = divmod(n_frames, world)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def global_batch_size(local_B: int, device=None, group=None) -> int:
    """sum of the per-rank batch sizes: the batch-global normalisers (landmark mean, per-frame regulariser means, (2 G B - 1) of the joint
    prior, n_pix of reg_diffuse) need the TRUE global batch, which is not local_B * world when the shards are uneven"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(local_B)
    t = torch.tensor([int(local_B)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


class TexShardComm:
    """collectives of the sharded texture update (Engine.tex_update): rank r owns the rows [r T/world, (r+1) T/world) of the texture.
    On its own communicator: the texture chain runs on a side stream beside the step's other collectives."""

    def __init__(self, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def owned(self):
        return [self.rank]

    def reduce_scatter(self, g_rm: torch.Tensor, bands: dict) -> None:
        dist.reduce_scatter_tensor(bands[self.rank], g_rm, op=dist.ReduceOp.SUM, group=self.group)

    def all_gather(self, ex_rm: torch.Tensor, bands: dict) -> None:
        dist.all_gather_into_tensor(ex_rm, bands[self.rank], group=self.group)


class TexPeerComm:
    """peer-memory texture update (csrc/dp_tex.cu): the exchange buffers are torch symmetric memory (one allocation per rank, mapped by
    every rank; with NVSwitch multicast the in-switch reduction / replication of NVLS is used), the kernels and barriers are the engine's
    own -- no collective call on the path.  Needs Engine.dp_connect (the barrier mailboxes) first."""
    peer = True

    def __init__(self, engine, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        nt = 3 * engine.T * engine.T
        g = group if group is not None else dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(g.group_name)
        except Exception:
            pass                                             # newer torch: enabled implicitly
        self.g_rm = symm_mem.empty(nt, dtype=torch.float32, device=engine.dev)
        self.ex_rm = symm_mem.empty(nt, dtype=torch.float32, device=engine.dev)
        self.g_rm.zero_(); self.ex_rm.zero_()
        hg, hx = symm_mem.rendezvous(self.g_rm, g), symm_mem.rendezvous(self.ex_rm, g)
        self.handles = (hg, hx)
        mg, mx = int(getattr(hg, "multicast_ptr", 0) or 0), int(getattr(hx, "multicast_ptr", 0) or 0)
        self.multicast = bool(mg and mx)
        engine.dp_tex_connect(list(hg.buffer_ptrs), mg if self.multicast else 0, list(hx.buffer_ptrs), mx if self.multicast else 0)

    def owned(self):
        return [self.rank]


class LocalShardComm:
    """the same interface inside ONE process that owns all `world` bands (sum over one rank = copy): exercises the band arithmetic of the
    sharded update without a second GPU (tests)"""

    def __init__(self, world: int):
        self.world, self.rank, self.group = world, 0, None

    def owned(self):
        return list(range(self.world))

    def reduce_scatter(self, g_rm, bands):
        n = g_rm.numel() // self.world
        for i, b in bands.items():
            b.copy_(g_rm[i * n:(i + 1) * n])

    def all_gather(self, ex_rm, bands):
        n = ex_rm.numel() // self.world
        for i, b in bands.items():
            ex_rm[i * n:(i + 1) * n].copy_(b)


class DataParallelStep:
    """Wraps an Engine: step(batch) = zero_grad, forward, slab reduce, backward, grad allreduce, Adam."""

    def __init__(self, engine, group=None, texture="auto", slab="peer"):
        """texture: "peer" (fold -> in-switch band reduction over NVSwitch multicast -> Adam on 1/world of the rows -> multicast store ->
        rebuild, all in the engine's own kernels over symmetric memory), "shard" (the same dataflow with NCCL reduce-scatter / all-gather,
        TexShardComm on its own communicator), "allreduce" (round-1 baseline: dense all-reduce of the regularised gradient, full-texture Adam
        on every rank) or "auto" (default: "peer", falling back to "shard" when symmetric memory cannot be set up; the choice is in
        .texture_mode).
        slab: "peer" (default: the mid-step batch-global scalars travel through CUDA-IPC mailboxes written / read by the engine's own
        kernels over NVLink, Engine.dp_connect) or "nccl" (round-1 baseline: an all-gather + host-side reduction between the halves)"""
        self.e, self.group = engine, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._gB = {}
        self.tex_comm = None
        self.peer_slab = False
        if self.world > 1 and slab == "peer" and hasattr(engine, "dp_connect"):
            def gather(b):
                out = [None] * self.world
                dist.all_gather_object(out, b, group=group)
                return out
            engine.dp_connect(dist.get_rank(group), self.world, gather)
            self.peer_slab = True
        self.texture_mode = "single" if self.world == 1 else texture
        # "auto" = the peer-memory path at every world size: measured (profiles/r02_scale.md, final state) 0.856 vs 0.861 ms/step at 2 ranks
        # (dense all-reduce + full-texture Adam), 0.850 vs 0.989 at 8
        if self.world > 1 and texture in ("peer", "auto"):
            # peer-memory update (NVLS when the fabric has multicast): needs the barrier mailboxes of the peer slab exchange
            try:
                if not self.peer_slab:
                    raise RuntimeError("texture='peer' needs slab='peer' (the barrier mailboxes)")
                self.tex_comm = TexPeerComm(engine, group)
                self.texture_mode = "peer+nvls" if self.tex_comm.multicast else "peer"
            except Exception as ex:
                if texture == "peer":
                    raise
                self.texture_mode = f"shard (peer path unavailable: {type(ex).__name__}: {str(ex)[:120]})"
                texture = "shard"
        if self.world > 1 and texture == "shard":
            ranks = dist.get_process_group_ranks(group) if group is not None else None
            self.tex_comm = TexShardComm(dist.new_group(ranks=ranks))

    def global_B(self, batch) -> int:
        """true global batch size of the step this batch belongs to (one host-side all-reduce per staged batch object, cached; must be
        called by all ranks in the same order -- step() / graph_begin() do)"""
        if batch.B < 1:
            raise ValueError("data parallel: every rank needs at least one frame per step")
        k = id(batch)
        if k not in self._gB:
            dev = self.e.dev if getattr(self.e, "dev", None) is not None and dist.is_initialized() and dist.get_backend(self.group) == "nccl" else None
            self._gB[k] = global_batch_size(batch.B, dev, self.group)
        return self._gB[k]

    def body(self, batch):
        """the step without host-side counters: what gets captured into a CUDA graph (Engine.graph_begin(body=...))"""
        e = self.e
        e.zero_grad()
        gB = self.global_B(batch)
        red = (lambda a, b: reduce_forward_slab(a, b, self.group)) if (self.world > 1 and not self.peer_slab) else None
        e.energy(batch, backward=True, training=True, global_B=gB, reduce_fn=red)
        e.adam_step(allreduce_fn=(lambda t: allreduce_sum(t, self.group)) if self.world > 1 else None, tex_comm=self.tex_comm)

    def graph_begin(self, batches, pipelined=True):
        """Engine.graph_begin with this group's collectives captured into the step graphs (pipelined: the texture all-reduce of step k
        runs at the start of step k+1, hidden behind FLAME / rasteriser / pools)"""
        red = (lambda a, b: reduce_forward_slab(a, b, self.group)) if (self.world > 1 and not self.peer_slab) else None
        allr = (lambda t: allreduce_sum(t, self.group)) if self.world > 1 else None
        gBs = [self.global_B(b) for b in batches]            # host-side collectives: before the capture
        self.e.graph_begin(batches, reduce_fn=red, allreduce_fn=allr, world=self.world, pipelined=pipelined, global_Bs=gBs, tex_comm=self.tex_comm)

    def step(self, batch):
        e = self.e
        e.zero_grad()
        gB = self.global_B(batch)
        red = (lambda a, b: reduce_forward_slab(a, b, self.group)) if (self.world > 1 and not self.peer_slab) else None
        losses = e.energy(batch, backward=True, training=True, global_B=gB, reduce_fn=red)
        e.adam_step(allreduce_fn=(lambda t: allreduce_sum(t, self.group)) if self.world > 1 else None, tex_comm=self.tex_comm)
        e.global_step += 1
        return losses
