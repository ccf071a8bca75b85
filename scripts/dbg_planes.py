"""Diagnostic: bf16-plane encoder path vs fp32-activation tensor path, element statistics."""
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
eng.set_precision('split')
for B in (3, 70):
    img = eng.to_dev(images(B, seed=5))
    out = {}
    for p in (0, 1, 1):
        eng.set_option('enc_planes', p)
        emb, fm, m5c = eng.encode(img, want_mixed5c=True)
        out.setdefault(p, []).append((fm.cpu().numpy().astype(np.float64), m5c.cpu().numpy().astype(np.float64)))
    eng.set_precision('f32')
    emb, fm, m5c = eng.encode(img, want_mixed5c=True)
    ref = fm.cpu().numpy().astype(np.float64)
    eng.set_precision('split')
    a, b = out[0][0][0], out[1][0][0]
    d = np.abs(a - b)
    print('B=%d planes-vs-fp32act: max %.3e (rel to max %.3e) mean %.3e frac_nonzero %.4f; rerun identical: %s' % (
        B, d.max(), d.max() / np.abs(a).max(), d.mean(), (d > 0).mean(), np.array_equal(out[1][0][0], out[1][1][0])))
    print('   vs FFMA: fp32act %.3e planes %.3e (max-normalised)' % (np.abs(a - ref).max() / np.abs(ref).max(),
                                                                     np.abs(b - ref).max() / np.abs(ref).max()))
