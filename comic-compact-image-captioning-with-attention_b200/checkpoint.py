"""TensorFlow V2 checkpoints (`<prefix>.index` + `<prefix>.data-00000-of-0000N`) without TensorFlow, and the
reference's restore rules (`ModelBase.restore_model`, src/model_base.py:422-490).

The reference restores either a whole COMIC model or only the slim InceptionV1 weights from such a checkpoint through
`tf.train.Saver` / `tf.train.NewCheckpointReader`; here the files are parsed directly into the W-table (a dict TF variable
name -> numpy array) that `Engine.bind_weights` / `CaptionModel.restore_model` take.

Format (tensorflow/core/util/tensor_bundle, as of TF r1.9 -- restated from the public sources, no TensorFlow and no
checkpoint produced by it is available in this environment, so this parser is validated against the writer below and
against the format's fixed points only: PARITY UNPINNED against a TF-written file):
  * `.index` is a LevelDB-format SSTable (tensorflow/core/lib/io/table*): data blocks of prefix-compressed
    (shared, non_shared, value_len varint32 triples + restart array) key -> value entries, each block followed by a
    1-byte compression type (0 none, 1 snappy) and a masked CRC-32C of block + type; an index block mapping separator
    keys to block handles; a 48-byte footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}; every other key is a tensor name ->
    BundleEntryProto {1: dtype, 2: shape {2: dim {1: size}}, 3: shard_id, 4: offset, 5: size, 6: crc32c (fixed32,
    masked), 7: slices}.
  * `.data-XXXXX-of-YYYYY` holds the raw little-endian tensor bytes at (offset, size).
`write_v2` produces the same structure (multi-block index, restart interval 16) so that weights trained here can be
handed back to TensorFlow tooling; V1 checkpoints (a single SSTable of SavedTensorSlices, e.g. slim's original
`inception_v1.ckpt`) are not parsed -- convert them with `tf.train.Saver(write_version=V2)` first.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------------------------
# CRC-32C (Castagnoli), LevelDB masking
# ---------------------------------------------------------------------------------------------------------------
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()


def crc32c(data, crc=0):
    crc ^= 0xffffffff
    for b in bytes(data):
        crc = _CRC_TABLE[(crc ^ b) & 0xff] ^ (crc >> 8)
    return crc ^ 0xffffffff


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xffffffff


# ---------------------------------------------------------------------------------------------------------------
# varints / minimal protobuf
# ---------------------------------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    res, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        res |= (b & 0x7f) << shift
        if not b & 0x80:
            return res, pos
        shift += 7
        if shift > 70:
            raise ValueError('malformed varint')


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """Yields (field_number, wire_type, value) of one serialized message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield f, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, slices=0)
    for f, _wt, v in _proto_fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _w2, v2 in _proto_fields(v):
                if f2 == 2:                              # TensorShapeProto.dim
                    size = 0
                    for f3, _w3, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
        elif f == 7:
            e['slices'] += 1
    return e


def _ld(field, payload):
    return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


def _serialize_entry(dtype_enum, shape, shard_id, offset, size, crc):
    out = _put_varint((1 << 3) | 0) + _put_varint(dtype_enum)
    dims = b''.join(_ld(2, _put_varint((1 << 3) | 0) + _put_varint(int(d))) for d in shape)
    out += _ld(2, dims)
    if shard_id:
        out += _put_varint((3 << 3) | 0) + _put_varint(shard_id)
    if offset:
        out += _put_varint((4 << 3) | 0) + _put_varint(offset)
    out += _put_varint((5 << 3) | 0) + _put_varint(size)
    out += _put_varint((6 << 3) | 5) + struct.pack('<I', crc)
    return out


# ---------------------------------------------------------------------------------------------------------------
# snappy (raw format) -- only needed when an index was written with block compression
# ---------------------------------------------------------------------------------------------------------------
def snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy stream')
        for _ in range(ln):                              # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy length mismatch')
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------
# SSTable
# ---------------------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify=True):
    body = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from('<I', data, offset + size + 1)[0]
    if verify and mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
        raise ValueError('index block at %d: CRC mismatch' % offset)
    if ctype == 0:
        return body
    if ctype == 1:
        return snappy_uncompress(body)
    raise ValueError('unknown block compression type %d' % ctype)


def _block_entries(block):
    nrestart = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    with open(path, 'rb') as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError('%s: too short for an SSTable' % path)
    footer = data[-48:]
    if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
        raise ValueError('%s: bad table magic (not a V2 checkpoint index?)' % path)
    pos = 0
    _mo, pos = _get_varint(footer, pos)
    _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    out = []
    for _sep, handle in _block_entries(_read_block(data, io, isz, verify)):
        bo, p = _get_varint(handle, 0)
        bs, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, bo, bs, verify)))
    return out


class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b''
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=4096):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order.  Uncompressed blocks, as TF's
    BundleWriter writes them."""
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)
        out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        return _put_varint(off) + _put_varint(len(block))

    bb, last_key = _BlockBuilder(), None
    for key, value in items:
        if last_key is not None and key <= last_key:
            raise ValueError('keys must be strictly increasing')
        bb.add(key, value)
        last_key = key
        if len(bb.buf) >= block_size:
            index.add(last_key, emit(bb.finish()))
            bb = _BlockBuilder()
    if bb.buf or last_key is None:
        index.add(last_key if last_key is not None else b'', emit(bb.finish()))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(out)


# ---------------------------------------------------------------------------------------------------------------
# V2 checkpoints
# ---------------------------------------------------------------------------------------------------------------
def _shard_path(prefix, shard, num_shards):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


def read_v2(prefix, names=None, verify_data=False):
    """dict name -> numpy array of the tensors in checkpoint `prefix` (all, or only `names`).  The index blocks' CRCs
    are always checked; `verify_data` also checks every tensor's CRC-32C (slow in pure Python for large tensors)."""
    entries = read_table(prefix + '.index')
    if not entries or entries[0][0] != b'':
        raise ValueError('%s.index: no bundle header' % prefix)
    num_shards, endianness = 1, 0
    for f, _wt, v in _proto_fields(entries[0][1]):
        if f == 1:
            num_shards = v
        elif f == 2:
            endianness = v
    if endianness != 0:
        raise NotImplementedError('big-endian checkpoints are not supported')
    want = None if names is None else set(names)
    shards, out = {}, {}
    for key, val in entries[1:]:
        name = key.decode('utf-8')
        if want is not None and name not in want:
            continue
        e = _parse_entry(val)
        if e['slices']:
            raise NotImplementedError('%s: partitioned (sliced) variables are not supported' % name)
        if e['dtype'] not in _DTYPES:
            raise NotImplementedError('%s: dtype enum %d' % (name, e['dtype']))
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap(_shard_path(prefix, sid, num_shards), dtype=np.uint8, mode='r')
        raw = np.asarray(shards[sid][e['offset']:e['offset'] + e['size']])
        dt = np.dtype(_DTYPES[e['dtype']])
        count = int(np.prod(e['shape'])) if e['shape'] else 1
        if raw.size != count * dt.itemsize:
            raise ValueError('%s: %d bytes for shape %s %s' % (name, raw.size, e['shape'], dt))
        if verify_data and e['crc32c'] and mask_crc(crc32c(raw.tobytes())) != e['crc32c']:
            raise ValueError('%s: tensor CRC mismatch' % name)
        out[name] = raw.view(dt).reshape(e['shape']).copy()
    return out


def variable_shapes(prefix):
    """NewCheckpointReader.get_variable_to_shape_map()."""
    return {k.decode('utf-8'): _parse_entry(v)['shape'] for k, v in read_table(prefix + '.index')[1:]}


def write_v2(prefix, tensors, with_data_crc=True):
    """Writes dict name -> array as `<prefix>.index` + `<prefix>.data-00000-of-00001` (the layout tf.train.Saver(V2)
    produces for unpartitioned variables), and the `checkpoint` state file next to it."""
    items, off = [], 0
    data = bytearray()
    for name in sorted(tensors, key=lambda s: s.encode('utf-8')):
        a = np.asarray(tensors[name])                       # (np.ascontiguousarray would turn a scalar into shape [1])
        if not a.flags.c_contiguous:
            a = a.copy()
        if a.dtype not in _DTYPE_ENUM:
            raise NotImplementedError('%s: dtype %s' % (name, a.dtype))
        raw = a.tobytes()
        crc = mask_crc(crc32c(raw)) if with_data_crc else 0
        items.append((name.encode('utf-8'), _serialize_entry(_DTYPE_ENUM[a.dtype], a.shape, 0, off, len(raw), crc)))
        data += raw
        off += len(raw)
    header = _put_varint((1 << 3) | 0) + _put_varint(1)                       # num_shards = 1
    header += _ld(3, _put_varint((1 << 3) | 0) + _put_varint(1))              # version { producer: 1 }
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    write_table(prefix + '.index', [(b'', header)] + items)
    with open(_shard_path(prefix, 0, 1), 'wb') as f:
        f.write(data)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), 'checkpoint'), 'w') as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the `model_checkpoint_path` of the directory's `checkpoint` state file."""
    state = os.path.join(checkpoint_dir, 'checkpoint')
    if not os.path.isfile(state):
        return None
    with open(state) as f:
        for line in f:
            m = re.match(r'\s*model_checkpoint_path:\s*"(.*)"', line)
            if m:
                p = m.group(1)
                p = p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)
                return p if os.path.isfile(p + '.index') else None
    return None


# ---------------------------------------------------------------------------------------------------------------
# ModelBase.restore_model (src/model_base.py:422-490)
# ---------------------------------------------------------------------------------------------------------------
CNN_SCOPE = 'Model/encoder/cnn/'


def _exclude_patterns(config):
    s = getattr(config, 'checkpoint_exclude_scopes', '') or ''
    return [sc.strip() for sc in s.split(',') if sc.strip()]


def restore_weights(config, weights, trainable_names=None, checkpoint_path=None):
    """The reference's three restore cases applied to a W-table.

    weights: dict variable name -> array (the freshly initialised model, `weights.init_weights`);
    trainable_names: the model's trainable variables (default: every key of `weights` that is not a BatchNorm moving
    statistic).  Returns (new W-table, info) with info['mode'] in {'scratch', 'resume', 'model', 'cnn'} and, when
    resuming, info['extra'] = every checkpoint tensor that is not a model variable (Adam slots, beta powers,
    global_step) for the caller's optimiser.
      * no checkpoint_path                       -> training from scratch;
      * trainable variables all in the checkpoint, no exclude scopes, resume_training -> whole checkpoint ('resume');
      * trainable variables all in the checkpoint otherwise -> every `Model` variable not matched (re.search) by an
        exclude scope ('model': fine-tuning);
      * else -> only variables under Model/encoder/cnn/, looked up WITHOUT that prefix (a slim InceptionV1
        checkpoint), again minus the exclude scopes ('cnn')."""
    path = checkpoint_path if checkpoint_path is not None else getattr(config, 'checkpoint_path', None)
    W = dict(weights)
    if not path:
        return W, dict(mode='scratch', restored=[])
    if not (os.path.isfile(path + '.index') or os.path.isfile(path)):
        path = latest_checkpoint(path)
        if path is None:
            raise IOError('no checkpoint found at %s' % (checkpoint_path or config.checkpoint_path))
    if not os.path.isfile(path + '.index'):
        raise NotImplementedError('%s is a V1 checkpoint: convert it with tf.train.Saver(write_version=V2)' % path)
    ckpt_vars = set(variable_shapes(path))
    if trainable_names is None:
        trainable_names = [n for n in W if not n.endswith(('moving_mean', 'moving_variance'))]
    exc = _exclude_patterns(config)
    keep = lambda n: not any(re.search(p, n) for p in exc)

    def load(names_map):                                 # model name -> checkpoint name
        got = read_v2(path, names=set(names_map.values()))
        for mname, cname in names_map.items():
            a = got[cname]
            if tuple(a.shape) != tuple(np.shape(W[mname])):
                raise ValueError('%s: checkpoint shape %s != model shape %s' % (mname, a.shape, np.shape(W[mname])))
            W[mname] = a.astype(np.asarray(W[mname]).dtype, copy=False)
        return sorted(names_map)

    if set(trainable_names).issubset(ckpt_vars):
        if not exc and getattr(config, 'resume_training', False):
            names = {n: n for n in W if n in ckpt_vars}
            restored = load(names)
            extra = read_v2(path, names=ckpt_vars - set(names))
            return W, dict(mode='resume', restored=restored, extra=extra, path=path)
        names = {n: n for n in W if n.startswith('Model') and n in ckpt_vars and keep(n)}
        return W, dict(mode='model', restored=load(names), path=path)
    names = {n: n[len(CNN_SCOPE):] for n in W if n.startswith(CNN_SCOPE) and keep(n)}
    missing = sorted(c for c in names.values() if c not in ckpt_vars)
    if missing:
        raise KeyError('CNN variables missing from %s: %s ...' % (path, missing[:3]))
    return W, dict(mode='cnn', restored=load(names), path=path)
