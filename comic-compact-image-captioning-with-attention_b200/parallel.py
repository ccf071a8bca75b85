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


def gpu_local_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off (NVML's ideal CPU affinity), restricted to the
    CPUs this process may use; empty set when NVML cannot tell."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        h = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(os.cpu_count() or 1, 1) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        return cpus & set(os.sched_getaffinity(0))
    except Exception:
        return set()


def bind_to_gpu_numa(device_index):
    """Pin this process (one per GPU) to the CPUs next to its GPU BEFORE it allocates pinned host
    buffers, so the per-batch H2D / D2H copies of the inference loop stay on the GPU's own socket.
    With 8 ranks streaming 0.5 GB per 23 ms batch each, remote-socket pinned memory is what the
    end-to-end rate loses first.  Returns the CPU set used (empty = left unchanged)."""
    import os
    cpus = gpu_local_cpus(device_index)
    if cpus and cpus != set(os.sched_getaffinity(0)):
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            return set()
    return cpus
