"""Diagnostic: stem conv variants (0 fp32 gather, 1 s2d planes from L2, 2 s2d planes from a smem halo) vs FFMA."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import numpy as np
from _common import comic_config, make_weights, images
from comic_b200.engine import Engine
c = comic_config()
W = make_weights(c)
eng = Engine(c)
eng.bind_weights(W)
img = eng.to_dev(images(5, seed=5))
eng.set_precision('f32')
ref = eng.encode(img)[1].cpu().numpy().astype(np.float64)
eng.set_precision('split')
out = {}
for m in (0, 1, 2, 2):
    eng.set_option('stem_s2d', m)
    out.setdefault(m, []).append(eng.encode(img)[1].cpu().numpy().astype(np.float64))
for m in (0, 1, 2):
    d = np.abs(out[m][0] - ref)
    print('stem_s2d=%d vs FFMA: max-normalised %.3e  mean abs %.3e' % (m, d.max() / np.abs(ref).max(), d.mean()))
d = np.abs(out[1][0] - out[2][0])
print('1 vs 2: max %.3e frac_nonzero %.4f; 2 rerun identical %s' % (d.max() / np.abs(ref).max(), (d > 0).mean(),
                                                                 np.array_equal(out[2][0], out[2][1])))
