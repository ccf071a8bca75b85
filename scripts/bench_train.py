#!/usr/bin/env python
"""Training throughput (BASELINE.json configs 3, 4 and 5) on N GPUs, one process per GPU.

    python scripts/bench_train.py --mode decoder --batch 32 --steps 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/bench_train.py --mode scst --batch 10 --steps 10

decoder: frozen-CNN encoder forward + teacher-forced fwd/bwd (T = 41 radix steps, every row
full length, seeded Philox dropout) + NCCL all-reduce of the flat gradient + Adam.
cnn_finetune: the same with the InceptionV1 forward-with-tape + backward (conv kernels + BN betas train).
scst: greedy + beam-7 sampling (40 steps), host CIDEr-D/BLEU reward, weighted-XE step.
Prints one JSON line on rank 0 (device-timed, max over ranks).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def run_mode(mode, batch=32, steps=20, warmup=3, precision='split'):
    """One training configuration on the ranks of the current process group (NCCL initialised by the caller when
    WORLD_SIZE > 1).  Returns the result dict on every rank (times are max over ranks)."""
    import torch
    import torch.distributed as dist
    import comic_b200  # noqa: F401
    from comic_b200 import configuration as conf, weights as wts, scst as S
    from comic_b200.train import Trainer
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    c = conf.make_config(train_mode=mode, batch_size_train=batch, max_step=100000)
    W = wts.init_weights(c, seed=c.rand_seed, cnn_init='he')
    tr = Trainer(c, W)
    eng = tr.engine
    eng.set_precision(precision)
    B = batch
    g = torch.Generator().manual_seed(100 + rank)
    images = torch.empty((B, 224, 224, 3)).uniform_(-1, 1, generator=g).to(eng.device)
    rng = np.random.default_rng(rank)
    if mode in ('decoder', 'cnn_finetune'):
        caps = np.concatenate([np.full((B, 1), 256), rng.integers(0, 256, size=(B, 40)), np.full((B, 1), 257)],
                              axis=1).astype(np.int32)                     # L = 42 -> T = 41, mask all ones

        def step(i):
            return tr.step(images, caps, None, seed=1000 + i)
    else:
        refs = [[' '.join('w%d' % w for w in rng.integers(0, 997, size=10)) for _ in range(5)] for _ in range(B)]
        df = {'document_frequency': S.compute_doc_freq(refs), 'ref_len': B}
        scorer = S.CaptionScorer(df, dict(ciderD=c.scst_weight_ciderD, bleu=c.scst_weight_bleu))

        def step(i):
            return S.scst_step(tr, scorer, images, refs, seed=1000 + i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for i in range(warmup):
        out = step(i)
    l0, r0 = eng.launch_count(), tr.replayed_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        out = step(warmup + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    host_launches = eng.launch_count() - l0
    launches = host_launches + (tr.replayed_launches - r0)     # kernels replayed from the step's CUDA graph included
    # the gradient exchange alone: NCCL sum all-reduce of the flat fp32 gradient buffer
    ar_ms = 0.0
    if world > 1:
        from comic_b200.parallel import allreduce_sum_
        scratch = tr.grads.clone()
        for _ in range(3):
            allreduce_sum_(scratch)
        barrier()
        ev0.record()
        for _ in range(10):
            allreduce_sum_(scratch)
        ev1.record()
        barrier()
        ar_ms = ev0.elapsed_time(ev1) / 10
    t = torch.tensor([ms, ar_ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = [float(x) for x in t.tolist()]
    res = {
        'metric': 'train steps/sec COMIC-256 train_mode=%s' % mode, 'value': steps / (ms * 1e-3),
        'unit': 'steps/s', 'n_gpus': world, 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms / steps, 'examples_per_sec': steps * B * world / (ms * 1e-3),
        'scaling': 'weak', 'dtype': 'f32' if precision == 'f32' else 'f32, GEMMs bf16x3 on tcgen05', 'data': 'synthetic',
        'config': {'workload': 'COMIC-256 %s, batch %d/GPU' % (mode, B), 'precision': precision,
                   'allreduce_bytes': int(tr.n_flat * 4)},
        'allreduce_ms': ar_ms,
        'loss': [float(x) for x in out['loss'].cpu().tolist()],
        'gpu_launches_per_step': int(launches // max(steps, 1)),
        'host_launches_per_step': int(host_launches // max(steps, 1)) + (1 if tr.replayed_launches > r0 else 0),
        'cuda_graph': bool(tr.replayed_launches > r0)}
    del tr, eng
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='decoder', choices=['decoder', 'cnn_finetune', 'scst'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--precision', default='split', choices=['f32', 'split', 'fast'])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')      # those levels print a version banner on stdout: keep it to the one JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    res = run_mode(args.mode, args.batch, args.steps, args.warmup, args.precision)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
