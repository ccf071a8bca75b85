"""Host-side input managers (comic_b200/inputs.py), CPU only: vocabulary files, word / radix / char tokenisation,
bucketing with the semantics of tf.contrib.data.bucket_by_sequence_length as the reference calls it
(common/inputs/manager_image_caption.py:83-86, 177-183, 231-357), max_step, the validation-perplexity loop."""
import json
import os

import numpy as np
import pytest

import comic_b200  # noqa: F401
from comic_b200 import configuration as conf, inputs


def _dataset(tmp_path, n_words=300, n_train=64, n_valid=8, pattern='coco_{}_v1'):
    """A tiny caption dataset in the reference's on-disk layout."""
    itow, wtoi = conf.synthetic_vocab(n_words)
    cap = tmp_path / 'captions'
    cap.mkdir()
    json.dump(itow, open(cap / (pattern.format('itow') + '.json'), 'w'))
    json.dump(wtoi, open(cap / (pattern.format('wtoi') + '.json'), 'w'))
    rng = np.random.default_rng(3)
    words = [w for w in wtoi if not w.startswith('<')]
    for split, n in (('train', n_train), ('valid', n_valid)):
        with open(cap / (pattern.format(split) + '.txt'), 'w') as f:
            for i in range(n):
                length = int(rng.integers(4, 17))
                body = [words[int(j)] for j in rng.integers(0, len(words), length)]
                if i == 0:
                    body[0] = 'neverseen'                                   # -> <UNK>
                f.write('%s_%d.jpg,%s\n' % (split, i, ' '.join(['<GO>'] + body + ['<EOS>'])))
    with open(cap / 'filenames_test.txt', 'w') as f:
        f.write('\n'.join('test_%d.jpg' % i for i in range(10)) + '\n')
    return wtoi


def _config(tmp_path, token_type, **kw):
    c = conf.make_config(token_type=token_type, dataset_dir=str(tmp_path), dataset_file_pattern='coco_{}_v1',
                         batch_size_train=4, batch_size_eval=4, max_epoch=3, **kw)
    return c


def test_number_to_base_and_bucket_rule():
    assert inputs.number_to_base(0, 256) == [0]
    assert inputs.number_to_base(255, 256) == [255]
    assert inputs.number_to_base(256, 256) == [1, 0]
    assert inputs.number_to_base(9999, 256) == [39, 15]
    with pytest.raises(ValueError):
        inputs.number_to_base(5, 1)
    # bucket id = number of boundaries <= length; a bucket is emitted when full, padded to ITS longest member
    ex = [(i, np.arange(n, dtype=np.int32)) for i, n in enumerate([3, 11, 4, 12, 10, 20, 15, 13, 14, 2])]
    got = list(inputs.bucket_batches(ex, [11, 13, 15], 2, -1))
    assert [k for k, _ in got] == [[0, 2], [1, 3], [5, 6], [7, 8], [4, 9]]
    assert got[0][1].shape == (2, 4) and got[0][1][0, 3] == -1
    assert got[2][1].shape == (2, 20) and got[4][1].shape == (2, 10)
    assert [k for k, _ in inputs.bucket_batches(ex[:9], [11, 13, 15], 2, -1)] == [[0, 2], [1, 3], [5, 6], [7, 8]]
    rest = list(inputs.bucket_batches(ex[:9], [11, 13, 15], 2, -1, drop_remainder=False))
    assert [k for k, _ in rest][-1] == [4]


def test_word_manager_files_unk_max_step_and_shuffle(tmp_path):
    wtoi = _dataset(tmp_path)
    c = _config(tmp_path, 'word')
    m = inputs.get_input_manager(c)
    assert type(m) is inputs.InputManager and m.buckets == [11, 13, 15]
    assert c.vocab_size == len(wtoi) - 1 and c.split_sizes == {'train': 64, 'valid': 8}       # <PAD> is not in itow
    assert c.max_step == int(64 / 4 * 3)
    first = dict((os.path.basename(p), ids) for p, ids in m.examples('valid', epochs=1))
    ids0 = first['valid_0.jpg']
    assert ids0[0] == wtoi['<GO>'] and ids0[-1] == wtoi['<EOS>'] and ids0[1] == wtoi['<UNK>']
    # validation keeps file order; training is shuffled from rand_seed, differently every epoch, reproducibly
    assert [os.path.basename(p) for p, _ in m.examples('valid', epochs=1)] == ['valid_%d.jpg' % i for i in range(8)]
    order = [os.path.basename(p) for p, _ in m.examples('train', epochs=2)]
    assert sorted(order[:64]) == sorted(order[64:]) and order[:64] != order[64:]
    assert order[:64] != ['train_%d.jpg' % i for i in range(64)]
    m2 = inputs.get_input_manager(_config(tmp_path, 'word'))
    assert [os.path.basename(p) for p, _ in m2.examples('train', epochs=2)] == order
    for paths, caps in m.batches('valid', epochs=1):
        assert caps.dtype == np.int32 and caps.shape[0] == 4 and len(paths) == 4
        lens = (caps != -1).sum(1)
        assert lens.max() == caps.shape[1]
        b = [sum(1 for x in m.buckets if x <= n) for n in lens]
        assert len(set(b)) == 1                                            # one bucket per batch


def test_radix_manager_digits_and_buckets(tmp_path):
    wtoi = _dataset(tmp_path)
    c = _config(tmp_path, 'radix', radix_base=16)
    m = inputs.get_input_manager(c)
    digits = len(inputs.number_to_base(len(wtoi), 16))
    assert digits == 3 and m.buckets == [33, 39, 45]
    ids = m.encode(['<GO>', 'w17', 'w255', 'nope', '<EOS>'])
    unk = inputs.number_to_base(wtoi['<UNK>'], 16)
    assert ids.tolist() == [16, 0, 1, 1, 0, 15, 15] + [0] * (3 - len(unk)) + unk + [17]
    images = lambda paths: np.zeros((len(paths), 8, 8, 3), np.uint8)
    m = inputs.get_input_manager(_config(tmp_path, 'radix', radix_base=16), image_loader=images)
    n = 0
    for im, caps in m.batches('train', epochs=1):
        assert im.shape == (4, 8, 8, 3) and caps.shape[0] == 4
        assert ((caps[:, 0] == 16) & (caps.max(1) == 17)).all() and (caps.shape[1] - 2) % 3 == 0
        n += 1
    assert 8 <= n <= 16                                                     # 64 examples, up to 4 partly filled buckets dropped


def test_char_manager_vocabulary_and_ids(tmp_path):
    _dataset(tmp_path)
    c = _config(tmp_path, 'char')
    m = inputs.get_input_manager(c)
    w = c.wtoi
    assert m.buckets == [45, 55, 70]
    assert w['<PAD>'] == -1 and w[' '] == 0 and w['0'] == 1 and w['9'] == 10 and w['a'] == 11 and w['z'] == 36
    assert w['<GO>'] == 38 and w['<EOS>'] == 39 and c.vocab_size == 40 and c.itow[39] == '<EOS>'
    ids = m.encode(['<GO>', 'w1', 'ab', '<EOS>'])
    assert ids.tolist() == [38, w['w'], w['1'], w[' '], w['a'], w['b'], 39]


def test_inference_file_list_and_eval_loop(tmp_path):
    _dataset(tmp_path)
    c = _config(tmp_path, 'word', infer_set='test', batch_size_infer=5)
    m = inputs.get_input_manager(c, is_inference=True)
    got = list(m.infer_batches())
    assert len(got) == 2 and got[0][1] == ['test_%d.jpg' % i for i in range(5)] and c.split_sizes['infer'] == 10
    c.batch_size_infer = 4
    with pytest.raises(AssertionError):
        inputs.get_input_manager(c, is_inference=True)

    class Fake(object):
        def __init__(self):
            self.v = iter([1.0, 2.0, 3.0])

        def eval_step(self, images, caps):
            return next(self.v)
    ppl = inputs.run_eval_loop(Fake(), [(None, None)] * 3)
    assert abs(ppl - np.exp(2.0)) < 1e-12
    with pytest.raises(ValueError):
        inputs.run_eval_loop(Fake(), [])
