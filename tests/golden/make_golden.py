"""Generates the committed golden fixtures in this directory from the NumPy oracle.

    python tests/golden/make_golden.py            # rewrites *.npz

The reference itself (TF 1.9 / Python 2.7) cannot run in this container, so the
fixtures are outputs of `oracle/` on seeded inputs, not of the reference; they
freeze the oracle (tests/test_oracle_known_answers.py) and give the CUDA path a
target that does not move (tests/test_gpu_parity.py::test_golden_*).  The only
piece of the reference that imports under Python 3 is its CIDEr-D scorer
(common/scst/cider_ruotianluo/pyciderevalcap/ciderD); `ciderd_case()` records
ITS outputs (generated here with /root/reference importable) for the SCST reward.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import comic_b200  # noqa: E402,F401
from comic_b200 import configuration as conf  # noqa: E402
from comic_b200 import weights as wts  # noqa: E402
import comic_oracle as O  # noqa: E402
import inception_v1_oracle as I  # noqa: E402


def golden_weights(c, include_cnn):
    return wts.perturb_for_parity(wts.init_weights(c, seed=20240101, cnn_init='he', include_cnn=include_cnn))


def decoder_inputs(B, seed):
    rng = np.random.default_rng(seed)
    fm = np.maximum(rng.standard_normal((B, 196, 832)), 0).astype(np.float32)
    im = np.maximum(rng.standard_normal((B, 1024)), 0).astype(np.float32)
    return im, fm


def decoder_beam_case():
    """COMIC-256 beam-3, 4 images, 10 radix steps."""
    c = conf.make_config()
    W = golden_weights(c, False)
    im, fm = decoder_inputs(4, 77)
    r = O.beam_search_decode(O.Decoder(W, c), im, fm, 3, 0.0, 10)
    _, _, am = O.post_process_beam(r, 8, 3)
    return dict(predicted_ids=r['predicted_ids'], parent_ids=r['parent_ids'], step_ids=r['step_ids'],
                lengths=r['lengths'], scores=r['scores'], attn_top=am.astype(np.float32))


def encoder_case():
    """2 images through InceptionV1: im_embed, a strided sample of Mixed_4f and its sum."""
    c = conf.make_config()
    W = golden_weights(c, True)
    rng = np.random.default_rng(5)
    img = rng.uniform(-1, 1, (2, 224, 224, 3)).astype(np.float32)
    emb, fm, _ = I.encoder(img, W, c)
    return dict(im_embed=emb, fm_sample=fm[:, ::7, ::13].copy(), fm_sum=np.float64(fm.astype(np.float64).sum()))


def main():
    np.savez_compressed(os.path.join(HERE, 'decoder_beam_comic256.npz'), **decoder_beam_case())
    np.savez_compressed(os.path.join(HERE, 'encoder_2img.npz'), **encoder_case())
    print('wrote fixtures to', HERE)


if __name__ == '__main__':
    main()
