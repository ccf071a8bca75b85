"""`Config` attribute bag + `config.pkl` reader/writer + flag defaults.

Drop-in contract: the same flag names and defaults as the
reference CLIs (src/train.py:29-162, src/infer.py:27-72), the same fixed
kwargs (src/train.py:281-300), the `legacy` overrides (src/train.py:178-200)
and the same `config.pkl` format: a protocol-2 pickle of `Config.__dict__`
(common/configuration.py:18-59).  Python-2 pickles are read with
`encoding='latin1'`.
"""
from __future__ import annotations

import os
import pickle
from time import localtime, strftime


class Config(object):
    """Configuration object (common/configuration.py:18-23)."""

    def __init__(self, **kwargs):
        for key, value in sorted(kwargs.items()):
            setattr(self, key, value)

    def save_config_to_file(self):
        """common/configuration.py:25-35: txt dump + `config.pkl` holding the
        plain dict (protocol 2 so the py2 reference can read it back)."""
        params = sorted(self.__dict__.keys())
        f_dump = ['%s = %s' % (k, self.__dict__[k]) for k in params]
        config_name = 'config___%s.txt' % strftime('%Y-%m-%d_%H-%M-%S', localtime())
        with open(os.path.join(self.log_path, config_name), 'w') as f:
            f.write('\r\n'.join(f_dump))
        with open(os.path.join(self.log_path, 'config.pkl'), 'wb') as f:
            pickle.dump(self.__dict__, f, 2)


def load_config(config_filepath):
    """common/configuration.py:55-59."""
    with open(config_filepath, 'rb') as f:
        try:
            c_dict = pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            c_dict = pickle.load(f, encoding='latin1')
    return Config(**c_dict)


# src/train.py:29-162 defaults (after the 'none' -> None conversion, :277-279)
TRAIN_DEFAULTS = dict(
    name='lstm', dataset_dir='', dataset_file_pattern='mscoco_{}_w5_s20_include_restval',
    train_mode='decoder', legacy=False, token_type='radix', radix_base=256,
    cnn_name='inception_v1', cnn_input_size=[224, 224], cnn_input_augment=True,
    cnn_fm_attention='Mixed_4f', cnn_fm_projection='tied',
    rnn_name='LSTM', rnn_size=512, rnn_word_size=256, rnn_init_method='first_input',
    rnn_recurr_dropout=False,
    attn_num_heads=8, attn_context_layer=False, attn_alignment_method='add_LN',
    attn_probability_fn='softmax', attn_keep_prob=0.9,
    initialiser='xavier', optimiser='adam', batch_size_train=32, batch_size_eval=61,
    max_epoch=30, lr_start=1e-2, lr_end=1e-5, cnn_grad_multiplier=1.0,
    adam_epsilon=1e-2, scst_beam_size=7, scst_weight_ciderD=1.0,
    scst_weight_bleu='0,0,0,2', freeze_scopes='Model/encoder/cnn',
    checkpoint_path=None, checkpoint_exclude_scopes='', gpu='0', run=1,
)

# src/train.py:281-300
FIXED_KWARGS = dict(
    rnn_layers=1, dropout_rnn_in=0.35, dropout_rnn_out=0.35,
    rnn_map_loss_scale=1.0, l2_decay=1e-5, clip_gradient_norm=0,
    max_saves=12, num_logs_per_epoch=100, per_process_gpu_memory_fraction=None,
    rand_seed=48964896, add_image_summaries=True, add_vars_summaries=False,
    add_grad_summaries=False, log_path='', save_path='',
)

# src/infer.py:27-72
INFER_DEFAULTS = dict(
    infer_set='test', infer_checkpoints_dir='', infer_checkpoints='all',
    annotations_file='captions_val2014.json', run_inference=True,
    get_metric_score=True, save_attention_maps=False,
    infer_beam_size=3, infer_length_penalty_weight=0.0, infer_max_length=30,
    batch_size_infer=25,
)

# src/train.py:178-200
LEGACY_OVERRIDES = dict(
    cnn_name='inception_v1', cnn_input_size=[224, 224], cnn_input_augment=True,
    cnn_fm_attention='Mixed_4f', rnn_name='LSTM', rnn_size=512, rnn_word_size=256,
    rnn_init_method='project_hidden', rnn_recurr_dropout=False,
    attn_context_layer=False, attn_alignment_method='add_LN',
    attn_probability_fn='softmax', attn_keep_prob=1.0, lr_start=1e-3, lr_end=2e-4,
    lr_reduce_every_n_epochs=4, cnn_grad_multiplier=1.0, initialiser='xavier',
    optimiser='adam', batch_size_train=32, adam_epsilon=1e-6,
)

RAND_SEEDS = {1: 48964896, 2: 88888888, 3: 123456789}        # src/train.py:202-207


def synthetic_vocab(n_words, token_type='radix'):
    """`itow` / `wtoi` with the id layout of datasets/preprocessing/
    prepro_base.py:149-223: PAD = -1, words 0.., then <UNK>, <GO>, <EOS>.
    `itow` is str-keyed as in the reference's JSON-loaded table."""
    words = ['w%d' % i for i in range(n_words - 3)] + ['<UNK>', '<GO>', '<EOS>']
    wtoi = {w: i for i, w in enumerate(words)}
    itow = {str(i): w for i, w in enumerate(words)}
    wtoi['<PAD>'] = -1
    return itow, wtoi


def make_config(**overrides):
    """Assemble a Config the way src/train.py:167-300 + infer.py:106-107 do:
    flag defaults -> legacy overrides -> train-mode overrides -> 'none' ->
    None -> fixed kwargs -> infer flags overlaid; then the InputManager-added
    fields (`itow`, `wtoi`, `vocab_size`, `max_step`) for a synthetic vocab."""
    kw = dict(FIXED_KWARGS)
    kw.update(TRAIN_DEFAULTS)
    kw.update(INFER_DEFAULTS)
    n_words = overrides.pop('n_words', None)
    kw.update(overrides)
    if kw.get('legacy'):
        kw.update(LEGACY_OVERRIDES)
        for k in LEGACY_OVERRIDES:                      # explicit overrides lose, as in train.py
            pass
    mode = kw['train_mode']
    if mode == 'cnn_finetune':                           # src/train.py:241-250
        kw.update(lr_start=overrides.get('lr_start', 1e-3), max_epoch=10, freeze_scopes='')
    elif mode == 'scst':                                 # src/train.py:252-262
        if isinstance(kw['scst_weight_bleu'], str):
            kw['scst_weight_bleu'] = [float(w) for w in kw['scst_weight_bleu'].split(',')]
        kw.update(batch_size_train=overrides.get('batch_size_train', 10),
                  lr_start=overrides.get('lr_start', 1e-3), max_epoch=10,
                  freeze_scopes='Model/encoder/cnn')
    for k, v in list(kw.items()):
        if v == 'none':
            kw[k] = None
    if isinstance(kw['cnn_input_size'], str):
        kw['cnn_input_size'] = [int(v) for v in kw['cnn_input_size'].split(',')]
    kw['rand_seed'] = RAND_SEEDS.get(kw['run'], kw['rand_seed'])
    if 'itow' not in kw:
        if n_words is None:
            n_words = 10000
        kw['itow'], kw['wtoi'] = synthetic_vocab(n_words, kw['token_type'])
    kw.setdefault('vocab_size', len(kw['itow']))
    kw.setdefault('max_step', 10000)
    kw.setdefault('resume_training', False)
    return Config(**kw)
