"""Shared helpers for the parity tests: seeded configs / weights / inputs."""
import numpy as np

import comic_b200  # noqa: F401  (import shim)
from comic_b200 import configuration as conf
from comic_b200 import weights as wts


def comic_config(**kw):
    return conf.make_config(**kw)


def word_config(n_words=1000, **kw):
    return conf.make_config(token_type='word', cnn_fm_projection='none', attn_num_heads=1,
                            n_words=n_words, **kw)


def make_weights(c, seed=1234, include_cnn=True):
    return wts.perturb_for_parity(wts.init_weights(c, seed=seed, cnn_init='he', include_cnn=include_cnn))


def images(B, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (B, 224, 224, 3)).astype(np.float32)


def fake_features(B, C=832, M=196, seed=3):
    """Encoder-free decoder inputs with realistic magnitudes (post-ReLU)."""
    rng = np.random.default_rng(seed)
    fm = np.maximum(rng.standard_normal((B, M, C)), 0).astype(np.float32)
    im = np.maximum(rng.standard_normal((B, 1024)), 0).astype(np.float32)
    return im, fm


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
