"""world_size-2 `gloo` tests (CPU) of the multi-process host logic: balanced image sharding
with host-side gather (inference: no collective on the data path) and the flat-gradient
all-reduce + mean that precedes the optimiser step (training)."""
import os
import socket

import numpy as np

import comic_b200  # noqa: F401
from comic_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 512, 513):
        for ws in (1, 2, 3, 8):
            parts = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        assert parallel.world() == (rank, ws)
        # training: per-rank gradient of a quadratic on its shard; sum all-reduce then 1/world
        rng = np.random.default_rng(0)
        data = rng.standard_normal((10, 16)).astype(np.float32)
        lo, hi = parallel.shard_range(10, rank, ws)
        g = torch.from_numpy(data[lo:hi].mean(axis=0).copy())
        n = parallel.allreduce_sum_(g)
        g = g / n
        # inference: every rank "captions" its own shard, results gathered on the host
        caps = ['img%d' % i for i in range(lo, hi)]
        allcaps = [c for part in parallel.gather_objects(caps) for c in part]
        q.put((rank, g.numpy(), allcaps))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allreduce_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    data = rng.standard_normal((10, 16)).astype(np.float32)
    want = (data[:5].mean(0) + data[5:].mean(0)) / 2
    for rank, g, allcaps in res:
        np.testing.assert_allclose(g, want, rtol=1e-6, atol=1e-6)
        assert allcaps == ['img%d' % i for i in range(10)]


def test_gpu_numa_binding_is_a_noop_without_nvml_or_gpu():
    """`parallel.bind_to_gpu_numa` (called by bench.py for N > 1 before pinned buffers are allocated) may only ever
    NARROW the CPU set to NVML's ideal CPUs for the GPU; without a GPU / NVML it must leave the process untouched."""
    import os
    from comic_b200 import parallel as par
    before = set(os.sched_getaffinity(0))
    local = par.gpu_local_cpus(0)
    assert isinstance(local, set) and local <= before
    used = par.bind_to_gpu_numa(0)
    after = set(os.sched_getaffinity(0))
    assert used == local
    assert after == (local if local else before)
    os.sched_setaffinity(0, before)
