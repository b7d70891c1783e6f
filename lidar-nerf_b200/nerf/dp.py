"""Data-parallel plumbing of the training step: rays shard across ranks with no data-path
collective; the ONE exchange per optimiser step is a sum all-reduce of the flat fp32 gradient (hash table + both
MLPs, 54.8 MB) over NCCL/NVLink, with the 1/world mean folded into Adam's gradient scale.  Backend-agnostic
(NCCL on GPUs, gloo in the CPU tests)."""
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
