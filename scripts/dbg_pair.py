"""Diagnostic: CTA-pair (cta_group::2) GEMM kernel vs the single-CTA kernel and fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import numpy as np
import torch
from _common import comic_config
from comic_b200.engine import Engine
c = comic_config()
eng = Engine(c)
eng.set_precision('split')
eng.set_option('gemm_pair_min_tiles', 1)
torch.manual_seed(0)
for (M, N, K) in [(256, 256, 64), (512, 256, 256), (300, 192, 576), (1536, 2048, 1280), (1000, 128, 200), (20000, 624, 832)]:
    A = torch.randn(M, K, device=eng.device)
    Bm = torch.randn(K, N, device=eng.device)
    ref = (A.double() @ Bm.double())
    eng.set_option('gemm_pair', 0)
    c0 = eng.gemm(A, Bm)
    eng.set_option('gemm_pair', 1)
    c1 = eng.gemm(A, Bm)
    torch.cuda.synchronize()
    e0 = float((c0.double() - ref).abs().max() / ref.abs().max())
    e1 = float((c1.double() - ref).abs().max() / ref.abs().max())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for mode, (a, b) in ((0, (0, 1)), (1, (2, 3))):
        eng.set_option('gemm_pair', mode)
        for _ in range(3):
            eng.gemm(A, Bm)
        ev[a].record()
        for _ in range(10):
            eng.gemm(A, Bm)
        ev[b].record()
    torch.cuda.synchronize()
    print('M=%d N=%d K=%d  err single %.2e pair %.2e  identical %s  us single %.1f pair %.1f' % (
        M, N, K, e0, e1, bool(torch.equal(c0, c1)), ev[0].elapsed_time(ev[1]) * 100, ev[2].elapsed_time(ev[3]) * 100))
