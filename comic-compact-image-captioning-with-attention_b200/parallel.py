"""One process per GPU: image sharding for inference (no collective) and the gradient
all-reduce for training.  The reference is single-GPU (CUDA_VISIBLE_DEVICES only,
src/train.py:172-173); this is the B200 scale-out of SURVEY.md §8e.

`torch.distributed` is plumbing: NCCL over NVLink 5 / NVSwitch on GPUs, gloo in the CPU
tests.  Inference shards are independent units; training exchanges ONE flat fp32 gradient
buffer per step (17.2 MB for COMIC-256 decoder mode).
"""
from __future__ import annotations


def world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [lo, hi) slice of `n_items` independent units for `rank`."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(flat):
    """In-place SUM all-reduce of the flat gradient buffer; returns the world size so the caller
    can fold 1/world into the optimiser (Trainer.apply_gradients -> comic_adam_step grad_scale)."""
    rank, ws = world()
    if ws > 1:
        import torch.distributed as dist
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return ws


def gather_objects(obj):
    """Host-side gather of per-rank results (captions) on every rank, in rank order."""
    rank, ws = world()
    if ws == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * ws
    dist.all_gather_object(out, obj)
    return out
