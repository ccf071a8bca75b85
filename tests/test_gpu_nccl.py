"""Two-rank NCCL check of the data-parallel training step: needs 2 GPUs
(`gpurun --gpus 2`), skipped otherwise.  Each rank computes its own gradient on its own batch; after
`Trainer.apply_gradients`' all-reduce the buffer on every rank holds the SUM, and Adam's 1 / world scale makes it
the mean of the two per-rank gradients (each of which is checked against fp64 autograd in test_gpu_train.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
    from _common import comic_config, make_weights, fake_features
    from comic_b200.train import Trainer
    from comic_b200.parallel import allreduce_sum_
    c = comic_config(train_mode='decoder', max_step=100)
    W = make_weights(c, include_cnn=False)
    tr = Trainer(c, W, with_cnn=False)
    eng = tr.engine
    grads = []
    for r in range(world):                                   # every rank computes BOTH per-rank gradients locally ...
        im, fm = fake_features(3, seed=40 + r)
        rng = np.random.default_rng(50 + r)
        caps = np.concatenate([np.full((3, 1), 256), rng.integers(0, 256, size=(3, 6)), np.full((3, 1), 257)], axis=1).astype(np.int32)
        tr.forward_backward(eng.to_dev(fm), eng.to_dev(im), caps)
        grads.append(tr.grads.clone())
    tr.grads.copy_(grads[rank])                              # ... and contributes its own to the exchange
    n = allreduce_sum_(tr.grads)
    assert n == world
    expect = grads[0] + grads[1]
    err = float((tr.grads - expect).abs().max().item() / (expect.abs().max().item() + 1e-30))
    before = tr.params.clone()
    tr.apply_gradients(lr=1e-3)                              # second all-reduce inside: exercise the real call path too
    moved = float((tr.params - before).abs().max().item())
    with open(os.path.join(out_dir, 'rank%d.txt' % rank), 'w') as f:
        f.write('%r %r' % (err, moved))
    dist.destroy_process_group()


def test_two_rank_allreduced_gradient_is_the_sum_of_the_per_rank_gradients(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29641, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        err, moved = [float(x) for x in open(tmp_path / ('rank%d.txt' % r)).read().split()]
        assert err < 1e-6, 'rank %d: all-reduced gradient differs from the sum of the per-rank gradients (%g)' % (r, err)
        assert moved > 0
