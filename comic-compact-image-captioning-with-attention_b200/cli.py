"""Command-line surface of the reference's `src/train.py` / `src/infer.py` on the CUDA engine.

The drop-in contract keeps the reference's flags: `create_train_parser()` mirrors
`train.py:create_parser` (:25-164) and `create_infer_parser()` mirrors `infer.py:create_parser`
(:23-74) flag for flag -- same names, types, defaults and choices (checked against the reference
sources in tests/test_host.py when they are mounted).  `config_from_train_args` assembles the
`Config` the way `train.py:167-300` does (legacy overrides, per-mode overrides, 'none' -> None,
fixed kwargs, run -> seed).  The reference's dataset readers are out of scope, so the
two `main`s drive the hot path on seeded synthetic batches of the configured shapes:

    python -m comic_b200.cli train --train_mode cnn_finetune --synthetic_steps 20
    python -m comic_b200.cli infer --infer_beam_size 3 --batch_size_infer 512 --synthetic_batches 8
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import configuration as conf

_CHOICES = dict(
    train_mode=['decoder', 'cnn_finetune', 'scst'], token_type=['radix', 'word', 'char'],
    cnn_fm_projection=['none', 'independent', 'tied'], rnn_name=['LSTM', 'LN_LSTM', 'GRU'],
    rnn_init_method=['project_hidden', 'first_input'], attn_alignment_method=['add_LN', 'add', 'dot'],
    attn_probability_fn=['softmax', 'sigmoid'], initialiser=['xavier', 'he', 'none'], optimiser=['adam', 'sgd'],
    infer_set=['test', 'valid', 'coco_test', 'coco_valid'])

# flags whose command-line default differs in TYPE from the assembled config value
_RAW_DEFAULTS = dict(cnn_input_size='224,224')


def _add(parser, name, default):
    kw = dict(default=_RAW_DEFAULTS.get(name, default))
    d = kw['default']
    kw['type'] = str if d is None else type(d)          # argparse `type=bool` as in the reference
    if name in _CHOICES:
        kw['choices'] = _CHOICES[name]
    parser.add_argument('--' + name, **kw)


def create_train_parser():
    """src/train.py:25-164."""
    p = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, default in conf.TRAIN_DEFAULTS.items():
        _add(p, name, default)
    return p


def create_infer_parser():
    """src/infer.py:23-74 (`dataset_dir`, `gpu`, `per_process_gpu_memory_fraction` included)."""
    p = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter)
    for name, default in conf.INFER_DEFAULTS.items():
        _add(p, name, default)
    _add(p, 'dataset_dir', '')
    _add(p, 'gpu', '0')
    _add(p, 'per_process_gpu_memory_fraction', 0.75)
    return p


def config_from_train_args(args, **extra):
    """train.py:167-300: parsed flags -> Config (synthetic vocabulary unless `itow` / `wtoi` are given)."""
    kw = dict(vars(args))
    for k in ('synthetic_steps', 'synthetic_seed', 'n_words'):
        kw.pop(k, None)
    # the reference FORCES these per mode (src/train.py:245-247, 258-261), whatever the command line says
    if kw['train_mode'] in ('cnn_finetune', 'scst'):
        if kw.get('legacy'):
            raise NotImplementedError                    # src/train.py:242, 253
        for k in ('lr_start', 'max_epoch') + (('batch_size_train',) if kw['train_mode'] == 'scst' else ()):
            kw.pop(k, None)
    kw.update(extra)
    return conf.make_config(**kw)


def train_main(argv=None):
    p = create_train_parser()
    p.add_argument('--synthetic_steps', type=int, default=10)
    p.add_argument('--synthetic_seed', type=int, default=0)
    p.add_argument('--n_words', type=int, default=10000)
    args = p.parse_args(argv)
    import torch
    from . import scst as S
    from . import weights as wts
    from .train import Trainer
    c = config_from_train_args(args, n_words=args.n_words, max_step=max(args.synthetic_steps, 1))
    W = wts.init_weights(c, seed=c.rand_seed, cnn_init='he')
    tr = Trainer(c, W)
    B = c.batch_size_train
    g = torch.Generator().manual_seed(args.synthetic_seed)
    images = torch.empty((B, 224, 224, 3)).uniform_(-1, 1, generator=g).to(tr.engine.device)
    rng = np.random.default_rng(args.synthetic_seed)
    if c.train_mode == 'scst':
        refs = [[' '.join('w%d' % w for w in rng.integers(0, 997, size=10)) for _ in range(5)] for _ in range(B)]
        df = {'document_frequency': S.compute_doc_freq(refs), 'ref_len': B}
        scorer = S.CaptionScorer(df, dict(ciderD=c.scst_weight_ciderD, bleu=c.scst_weight_bleu))
        step = lambda i: S.scst_step(tr, scorer, images, refs, seed=1000 + i)
    else:
        go, eos = (c.radix_base, c.radix_base + 1) if c.token_type == 'radix' else (c.wtoi['<GO>'], c.wtoi['<EOS>'])
        hi = c.radix_base if c.token_type == 'radix' else len(c.itow) - 3
        caps = np.concatenate([np.full((B, 1), go), rng.integers(0, hi, size=(B, 40)), np.full((B, 1), eos)],
                              axis=1).astype(np.int32)
        step = lambda i: tr.step(images, caps, None, seed=1000 + i)
    t0 = time.perf_counter()
    for i in range(args.synthetic_steps):
        out = step(i)
        print('step %4d  loss %.4f  (xe %.4f  map %.4f  reg %.4f)' % ((tr.global_step,) + tuple(out['loss'].cpu().tolist())))
    torch.cuda.synchronize()
    print('%.2f steps/s' % (args.synthetic_steps / (time.perf_counter() - t0)))
    return 0


def infer_main(argv=None):
    p = create_infer_parser()
    for name in ('token_type', 'cnn_fm_projection', 'attn_num_heads', 'rnn_size', 'legacy'):
        _add(p, name, conf.TRAIN_DEFAULTS[name])
    p.add_argument('--synthetic_batches', type=int, default=4)
    p.add_argument('--n_words', type=int, default=10000)
    args = p.parse_args(argv)
    import torch
    from . import weights as wts
    from .model import CaptionModel
    kw = {k: v for k, v in vars(args).items() if k not in ('synthetic_batches', 'n_words')}
    c = conf.make_config(n_words=args.n_words, **kw)
    W = wts.init_weights(c, seed=c.rand_seed, cnn_init='he')
    m = CaptionModel(c, 'infer', weights=W)
    B = c.batch_size_infer
    host = [torch.empty((B, 224, 224, 3)).uniform_(-1, 1, generator=torch.Generator().manual_seed(s)).pin_memory()
            for s in range(2)]
    t0 = time.perf_counter()
    n = 0
    for preds, attn in m.run_stream(host[i % 2] for i in range(args.synthetic_batches)):
        n += preds.shape[0]
    dt = time.perf_counter() - t0
    print('%d captions in %.3f s: %.1f captions/s (first batch includes warm-up)' % (n, dt, n / dt))   # infer_fn.py:176-184
    return 0


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ('train', 'infer'):
        print(__doc__)
        return 2
    return train_main(argv[1:]) if argv[0] == 'train' else infer_main(argv[1:])


if __name__ == '__main__':
    sys.exit(main())
