"""Decode loop alone (no encoder) at the benchmark shape, for ncu captures of the decoder-step kernels:
   ncu --set full -k regex:gemm_bf16x3 -s 20 -c 2 python scripts/decode_only.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import __graft_entry__ as G

G.build()
from _common import comic_config, word_config, make_weights, fake_features
from comic_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
c = word_config(n_words=10000) if (len(sys.argv) > 3 and sys.argv[3] == 'word') else comic_config()   # BASELINE config 2 / 1
W = make_weights(c, include_cnn=False)
eng = Engine(c)
eng.bind_weights(W, with_cnn=False)
im, fm = fake_features(B, seed=3)
keys, values = eng.project_fm(eng.to_dev(fm))
c0, h0 = eng.rnn_init(eng.to_dev(im))
for _ in range(2):
    r = eng.decode_beam(keys, values, c0, h0, 3, 0.0, steps, want_attn=False)
torch.cuda.synchronize()
print('ok', int(r['T'].item()))
