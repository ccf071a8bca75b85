#!/usr/bin/env python
"""Per-phase timing of the persistent decode loop (clock64 stamps of CTA 0 and the first selection CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import comic_b200  # noqa
from comic_b200 import configuration as conf, weights as wts
from comic_b200.engine import Engine

B, k, T = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 3, 60
c = conf.make_config()
W = wts.init_weights(c, seed=1, cnn_init='he', include_cnn=False)
eng = Engine(c)
eng.bind_weights(W, with_cnn=False)
eng.set_option('persistent_trace', int(sys.argv[2]) if len(sys.argv) > 2 else 1)
rng = np.random.default_rng(0)
fm = eng.to_dev(np.maximum(rng.standard_normal((B, 196, 832)), 0).astype(np.float32))
im = eng.to_dev(np.maximum(rng.standard_normal((B, 1024)), 0).astype(np.float32))
keys, values = eng.project_fm(fm)
c0, h0 = eng.rnn_init(im)
for _ in range(3):
    r = eng.decode_beam(keys, values, c0, h0, k, 0.0, T)
torch.cuda.synchronize()
tr = eng.decode_trace(T).astype(np.float64)
order = [0, 9, 10, 11, 12, 1, 2, 3, 4, 5, 6, 7, 8]
d = np.diff(tr[5:, 0, order], axis=1) / 1.965e3      # us, CTA 0
names = ['A1 ptrs', 'A1 x load', 'A1 gemm', 'bar A1', 'A2 cell', 'bar A2', 'B work', 'bar B', 'C1 work', 'bar C1', 'C2 work', 'bar C2']
print('CTA 0, mean us over steps 5..:')
for n, v in zip(names, d.mean(0)):
    print('  %-10s %6.2f' % (n, v))
print('  total    %6.2f' % d.sum(1).mean())
s = tr[5:, 1, :]
print('selection CTA: arrive->done (stamps 5..7) %.2f us; step %.2f us' %
      (((s[:, 7] - s[:, 5]) / 1.965e3).mean(), ((s[:, 8] - s[:, 0]) / 1.965e3).mean()))
