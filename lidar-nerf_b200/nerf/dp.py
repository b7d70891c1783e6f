"""Data-parallel plumbing of the training step: rays shard across ranks with no data-path
collective; the ONE exchange per optimiser step is reduce-scatter(fp32 gradient) -> Adam on this rank's shard ->
all-gather(fp16 parameters), with the 1/world mean folded into Adam's gradient scale.

This module is the NCCL form of that exchange (and the layout both forms share); when torch symmetric memory is
available the engine replaces the three calls by ONE peer-memory kernel over NVLink (lnb_dp_adam_exchange,
nerf/engine.py).  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_gradient_(flat_grad):
    """In-place SUM over ranks of the flat gradient vector (no-op on one rank)."""
    if is_distributed():
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def grad_scale(loss_scale):
    """Factor Adam applies to the summed gradient: undo the static loss scale and average over ranks."""
    return 1.0 / (loss_scale * world_size())


def shard_bounds(n, r, world):
    """Contiguous [lo, hi) slice of an n-element vector owned by rank r (for reduce-scatter style updates)."""
    base, rem = divmod(n, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def rank_seed(base_seed):
    """Each rank draws its own rays: seed = base + rank."""
    return base_seed + rank()


class ShardedExchange:
    """Reduce-scatter -> (caller updates its shard) -> all-gather, over a flat vector padded to world * align.

    The data-parallel step then moves 4 B/param (fp32 gradient reduce-scatter) + 2 B/param (fp16 parameter all-gather)
    per rank instead of 8 B/param for a gradient all-reduce, and the optimiser touches 1/world of the Adam state.
    """

    def __init__(self, n, align=8, local_only=False):
        """local_only: a one-rank layout regardless of the process group (parameters owned and synchronised by someone
        else, e.g. a DistributedDataParallel wrapper around the nn.Module path)."""
        self.world, self.rank = (1, 0) if local_only else (world_size(), rank())
        unit = self.world * align
        self.n = n
        self.n_padded = (n + unit - 1) // unit * unit
        self.shard = self.n_padded // self.world
        self.lo = self.rank * self.shard
        self.hi = self.lo + self.shard

    def reduce_scatter(self, full_padded, out_shard):
        """out_shard <- sum over ranks of full_padded[lo:hi]."""
        if self.world == 1:
            out_shard.copy_(full_padded[self.lo:self.hi])
        else:
            dist.reduce_scatter_tensor(out_shard, full_padded, op=dist.ReduceOp.SUM)
        return out_shard

    def all_gather(self, full_padded, shard):
        """full_padded <- concatenation over ranks of `shard`."""
        if self.world == 1:
            full_padded[self.lo:self.hi].copy_(shard)
        else:
            dist.all_gather_into_tensor(full_padded, shard)
        return full_padded


def all_ranks_agree(go, device=None):
    """True only if EVERY rank passes True (all-reduce MIN; identity on one rank).  For loops whose exit test is rank-local
    (a wall-clock budget, a data-dependent stop) but whose body contains collectives: every rank must leave the loop in
    the same iteration, or the ranks that go on wait for partners that never come."""
    if not is_distributed():
        return bool(go)
    import torch
    flag = torch.tensor([1 if go else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(int(flag.item()))

