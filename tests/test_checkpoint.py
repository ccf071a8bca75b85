"""TF V2 checkpoint reader / writer and the reference's restore rules (comic_b200/checkpoint.py), CPU only.
No TensorFlow-written file exists in this environment: the parser is checked against the format's fixed points (CRC-32C
test vector, LevelDB footer magic and block trailer), against hand-assembled tables (prefix compression, restarts,
snappy blocks) and against the writer."""
import os
import struct

import numpy as np
import pytest

import comic_b200  # noqa: F401
from comic_b200 import checkpoint as ck
from comic_b200 import configuration as conf, weights as wts


def test_crc32c_and_mask_known_answers():
    assert ck.crc32c(b'123456789') == 0xE3069283                       # the CRC-32C (Castagnoli) check value
    assert ck.crc32c(b'') == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4: 32 bytes of zeros
    assert ck.mask_crc(0) == 0xa282ead8


def test_table_round_trip_multi_block_and_prefix_compression(tmp_path):
    keys = [b''] + sorted(('Model/decoder/rnn_decoder/var_%04d/kernel' % i).encode() for i in range(300))
    items = [(k, (b'v' + k[::-1]) * (1 + i % 3)) for i, k in enumerate(keys)]
    p = str(tmp_path / 't.index')
    ck.write_table(p, items, block_size=512)
    assert ck.read_table(p) == items
    raw = open(p, 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == ck.TABLE_MAGIC
    bad = bytearray(raw)
    bad[10] ^= 0x40
    open(p, 'wb').write(bad)
    with pytest.raises(ValueError):
        ck.read_table(p)
    with pytest.raises(ValueError):
        ck.write_table(p, [(b'b', b'1'), (b'a', b'2')])


def test_snappy_blocks_are_read(tmp_path):
    """A table whose data block is snappy-compressed (literal + overlapping copy elements)."""
    block = ck._BlockBuilder()
    block.add(b'', b'hdr')
    block.add(b'abcabcabcabc', b'xyz' * 20)
    plain = block.finish()
    # hand-made snappy stream of `plain`: all literals, except one run expressed as a copy when possible
    comp = ck._put_varint(len(plain))
    pos = 0
    while pos < len(plain):
        chunk = plain[pos:pos + 60]
        comp += bytes([(len(chunk) - 1) << 2]) + chunk
        pos += len(chunk)
    assert ck.snappy_uncompress(comp) == plain
    assert ck.snappy_uncompress(ck._put_varint(10) + bytes([(1 - 1) << 2]) + b'a' + bytes([((9 - 4) << 2) | 1, 1])) == b'a' * 10
    out = bytearray()
    off = len(out)
    out += comp + b'\x01' + struct.pack('<I', ck.mask_crc(ck.crc32c(comp + b'\x01')))
    handle = ck._put_varint(off) + ck._put_varint(len(comp))
    idx = ck._BlockBuilder(restart_interval=1)
    idx.add(b'abcabcabcabc', handle)
    def emit(b):
        o = len(out)
        out.extend(b + b'\x00' + struct.pack('<I', ck.mask_crc(ck.crc32c(b + b'\x00'))))
        return ck._put_varint(o) + ck._put_varint(len(b))
    meta = emit(ck._BlockBuilder().finish())
    ih = emit(idx.finish())
    footer = meta + ih
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', ck.TABLE_MAGIC)
    p = str(tmp_path / 's.index')
    open(p, 'wb').write(out)
    assert ck.read_table(p) == [(b'', b'hdr'), (b'abcabcabcabc', b'xyz' * 20)]


def _model(**kw):
    c = conf.make_config(**kw)
    return c, wts.init_weights(c, seed=5, cnn_init='he')


def test_v2_round_trip_of_a_w_table(tmp_path):
    c, W = _model()
    prefix = str(tmp_path / 'model_compact-123')
    extra = {'global_step': np.array(123, np.int64), 'beta1_power': np.array(0.5, np.float32)}
    ck.write_v2(prefix, dict(W, **extra), with_data_crc=False)      # (pure-Python CRC-32C of 40 MB is slow; checked below)
    shapes = ck.variable_shapes(prefix)
    assert shapes['Model/decoder/rnn_decoder/embedding_map'] == [258, 256] and shapes['global_step'] == []
    got = ck.read_v2(prefix, verify_data=False)
    assert set(got) == set(W) | set(extra)
    for k in W:
        np.testing.assert_array_equal(got[k], np.asarray(W[k]))
        assert got[k].dtype == np.asarray(W[k]).dtype
    small = ck.read_v2(prefix, names=['Model/decoder/rnn_decoder/softmax_temperature', 'global_step'])
    assert int(small['global_step']) == 123
    assert ck.latest_checkpoint(str(tmp_path)) == prefix
    # a flipped data byte is caught by the per-tensor CRC (small checkpoint written with CRCs)
    prefix = str(tmp_path / 'small' / 'model-1')
    ck.write_v2(prefix, {k: W[k] for k in W if 'rnn_decoder' in k and 'kernel' not in k and 'embedding' not in k})
    ck.read_v2(prefix, verify_data=True)
    d = prefix + '.data-00000-of-00001'
    raw = bytearray(open(d, 'rb').read())
    e = ck._parse_entry(dict(ck.read_table(prefix + '.index'))[b'Model/decoder/rnn_decoder/softmax_temperature'])
    raw[e['offset']] ^= 1
    open(d, 'wb').write(raw)
    with pytest.raises(ValueError):
        ck.read_v2(prefix, names=['Model/decoder/rnn_decoder/softmax_temperature'], verify_data=True)


def test_restore_rules_follow_the_reference(tmp_path):
    """src/model_base.py:422-490: resume / fine-tune with exclude scopes / CNN-only from a slim-named checkpoint."""
    c, W0 = _model()
    _, W1 = _model()
    rng = np.random.default_rng(0)
    trained = {k: (np.asarray(v) + rng.standard_normal(np.shape(v)).astype(np.float32) * 0.01).astype(np.float32) for k, v in W1.items()}
    prefix = str(tmp_path / 'run' / 'model_compact-7')
    ck.write_v2(prefix, dict(trained, global_step=np.array(7, np.int64),
                             **{'Model/decoder/rnn_decoder/embedding_map/Adam': np.ones((258, 256), np.float32)}),
                with_data_crc=False)
    # (1) nothing configured -> scratch
    W, info = ck.restore_weights(c, W0)
    assert info['mode'] == 'scratch'
    # (2) resume: whole checkpoint, optimiser tensors handed back; a directory resolves through the state file
    c.checkpoint_path, c.resume_training, c.checkpoint_exclude_scopes = str(tmp_path / 'run'), True, ''
    W, info = ck.restore_weights(c, W0)
    assert info['mode'] == 'resume' and int(info['extra']['global_step']) == 7
    assert 'Model/decoder/rnn_decoder/embedding_map/Adam' in info['extra']
    for k in W0:
        np.testing.assert_array_equal(W[k], trained[k])
    # (3) fine-tune: exclude scopes are regular expressions searched in the variable name
    c.resume_training, c.checkpoint_exclude_scopes = False, 'output_projection, embedding_map'
    W, info = ck.restore_weights(c, W0)
    assert info['mode'] == 'model'
    k_out = 'Model/decoder/rnn_decoder/output_projection/kernel'
    np.testing.assert_array_equal(W[k_out], np.asarray(W0[k_out]))
    np.testing.assert_array_equal(W['Model/decoder/rnn_decoder/memory_layer/kernel'], trained['Model/decoder/rnn_decoder/memory_layer/kernel'])
    # (4) a slim InceptionV1 checkpoint (names without the Model/encoder/cnn/ prefix) restores the CNN only
    slim = {k[len(ck.CNN_SCOPE):]: v for k, v in trained.items() if k.startswith(ck.CNN_SCOPE)}
    sp = str(tmp_path / 'slim' / 'inception_v1.ckpt')
    ck.write_v2(sp, slim, with_data_crc=False)
    c.checkpoint_path, c.checkpoint_exclude_scopes = sp, ''
    W, info = ck.restore_weights(c, W0)
    assert info['mode'] == 'cnn' and all(n.startswith(ck.CNN_SCOPE) for n in info['restored'])
    kc = ck.CNN_SCOPE + 'InceptionV1/Conv2d_1a_7x7/weights'
    np.testing.assert_array_equal(W[kc], trained[kc])
    np.testing.assert_array_equal(W[k_out], np.asarray(W0[k_out]))
    with pytest.raises(IOError):
        ck.restore_weights(c, W0, checkpoint_path=str(tmp_path / 'nowhere'))
