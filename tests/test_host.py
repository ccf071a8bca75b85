"""Host-side logic (CPU): config.pkl contract, flag defaults against the
reference's argparse (when /root/reference is mounted), radix helpers."""
import os
import pickle
import re

import pytest

import comic_b200  # noqa: F401
from comic_b200 import configuration as conf

REF = '/root/reference/src'


def test_config_pickle_round_trip(tmp_path):
    c = conf.make_config(log_path=str(tmp_path))
    c.save_config_to_file()
    p = os.path.join(str(tmp_path), 'config.pkl')
    with open(p, 'rb') as f:
        raw = pickle.load(f)
    assert isinstance(raw, dict) and raw['rnn_size'] == 512 and raw['cnn_fm_projection'] == 'tied'
    c2 = conf.load_config(p)
    assert c2.__dict__ == c.__dict__
    # protocol 2, as common/configuration.py:34-35 writes it
    assert open(p, 'rb').read(2) == b'\x80\x02'


def test_python2_pickle_is_readable(tmp_path):
    """A py2 cPickle protocol-2 dict with a non-ASCII byte string loads via the latin1 retry."""
    blob = b'\x80\x02}q\x00(U\x04nameq\x01U\x04caf\xe9q\x02U\x08rnn_sizeq\x03M\x00\x02u.'
    p = tmp_path / 'config.pkl'
    p.write_bytes(blob)
    c = conf.load_config(str(p))
    assert c.rnn_size == 512


def test_north_star_fields_present():
    c = conf.make_config()
    for f in ('token_type', 'attn_num_heads', 'cnn_fm_projection', 'rnn_size', 'infer_beam_size', 'scst_beam_size',
              'legacy', 'radix_base', 'rnn_word_size', 'attn_keep_prob', 'dropout_rnn_in', 'dropout_rnn_out',
              'l2_decay', 'adam_epsilon', 'lr_start', 'lr_end', 'itow', 'wtoi', 'vocab_size', 'max_step'):
        assert hasattr(c, f), f
    assert c.cnn_input_size == [224, 224]


def test_none_string_becomes_none_and_mode_overrides():
    c = conf.make_config(cnn_fm_projection='none', token_type='word', attn_num_heads=1)
    assert c.cnn_fm_projection is None
    s = conf.make_config(train_mode='scst')
    assert s.batch_size_train == 10 and s.lr_start == 1e-3 and s.scst_weight_bleu == [0.0, 0.0, 0.0, 2.0]
    f = conf.make_config(train_mode='cnn_finetune')
    assert f.freeze_scopes == '' and f.lr_start == 1e-3 and f.max_epoch == 10
    leg = conf.make_config(legacy=True)
    assert leg.rnn_init_method == 'project_hidden' and leg.attn_keep_prob == 1.0 and leg.adam_epsilon == 1e-6


def _argparse_defaults(path):
    """(flag -> default literal) scraped from the reference's add_argument calls."""
    src = open(path).read()
    out = {}
    for m in re.finditer(r"add_argument\(\s*'--(\w+)'(.*?)\)\s*\n", src, flags=re.S):
        d = re.search(r"default=([^,\n]+)", m.group(2))
        if d:
            out[m.group(1)] = d.group(1).strip()
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not mounted')
def test_flag_defaults_match_reference_cli():
    """Every flag of src/train.py:29-162 / src/infer.py:27-72 exists with the same default."""
    import ast
    for fname, table in (('train.py', conf.TRAIN_DEFAULTS), ('infer.py', conf.INFER_DEFAULTS)):
        ref = _argparse_defaults(os.path.join(REF, fname))
        assert len(ref) > 10
        for flag, lit in ref.items():
            if flag in ('gpu', 'per_process_gpu_memory_fraction', 'dataset_dir', 'infer_checkpoints_dir'):
                continue
            assert flag in table, '%s: flag --%s missing' % (fname, flag)
            try:
                val = ast.literal_eval(lit)
            except Exception:
                continue
            ours = table[flag]
            if flag == 'cnn_input_size':
                val = [int(v) for v in val.split(',')]
            if val == 'none':
                val = None if ours is None else val
            if isinstance(val, str) and val == '' and ours in ('', None):
                continue
            assert ours == val, '%s --%s: %r != reference %r' % (fname, flag, ours, val)


def test_number_to_base_and_max_iterations():
    from comic_b200.model import number_to_base
    assert number_to_base(0, 256) == [0]
    assert number_to_base(255, 256) == [255]
    assert number_to_base(256, 256) == [1, 0]
    assert number_to_base(9999, 256) == [39, 15]
    with pytest.raises(ValueError):
        number_to_base(5, 1)


def test_synthetic_vocab_layout():
    """datasets/preprocessing/prepro_base.py:149-223: words 0.., <UNK>, <GO>, <EOS>; PAD = -1."""
    itow, wtoi = conf.synthetic_vocab(1000)
    assert wtoi['<EOS>'] == 999 and wtoi['<GO>'] == 998 and wtoi['<UNK>'] == 997 and wtoi['<PAD>'] == -1
    assert itow['0'] == 'w0' and len(itow) == 1000
