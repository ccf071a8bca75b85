"""cProfile of the SCST step (host side) at the reference's batch 10 / beam 7: where the non-GPU time goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import numpy as np
import torch
import __graft_entry__ as G

G.build()
import comic_b200  # noqa
from comic_b200 import configuration as conf, weights as wts, scst as S
from comic_b200.train import Trainer

c = conf.make_config(train_mode='scst', max_step=1000)
W = wts.init_weights(c, seed=1, cnn_init='he')
tr = Trainer(c, W)
eng = tr.engine
B = 10
rng = np.random.default_rng(0)
images = torch.empty((B, 224, 224, 3)).uniform_(-1, 1).to(eng.device)
refs = [[' '.join('w%d' % w for w in rng.integers(0, 997, size=10)) for _ in range(5)] for _ in range(B)]
df = {'document_frequency': S.compute_doc_freq(refs), 'ref_len': B}
scorer = S.CaptionScorer(df, dict(ciderD=c.scst_weight_ciderD, bleu=c.scst_weight_bleu))
for i in range(5):
    S.scst_step(tr, scorer, images, refs, seed=i)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    S.scst_step(tr, scorer, images, refs, seed=10 + i)
torch.cuda.synchronize()
pr.disable()
print('ms per step', (time.perf_counter() - t0) * 50)
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
