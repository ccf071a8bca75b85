"""The C-ABI library builds for sm_100a, loads without a GPU and exports every
symbol include/comic_b200.h declares.  No compute entry point is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'comic_b200.h')


@pytest.fixture(scope='module')
def lib_path():
    import __graft_entry__ as G
    return G.build()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(comic_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for need in ('comic_create', 'comic_destroy', 'comic_bind_weights', 'comic_encode_fwd', 'comic_project_fm',
                 'comic_rnn_init', 'comic_decode_step', 'comic_decode_greedy', 'comic_decode_beam',
                 'comic_beam_step', 'comic_gather_tree', 'comic_workspace_bytes', 'comic_last_error'):
        assert need in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), 'libcomic_b200.so does not export %s' % s


def test_python_binding_covers_the_header(lib_path):
    from comic_b200 import engine
    assert sorted(engine.SIGNATURES) == declared_symbols()
    engine.load_library(lib_path)


def test_struct_layouts_match_header():
    """ctypes mirrors of comic_cfg_t / comic_weights_t have the C sizes."""
    from comic_b200 import engine
    assert ctypes.sizeof(engine.ComicCfg) == 16 * 4
    assert ctypes.sizeof(engine.ComicWeights) == (17 + 4 * 57) * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(engine.ComicConvDesc) == 16


def test_conv_table_matches_weight_container(lib_path):
    """comic_conv_table() (host-only call) == the W-table's 57 conv descriptors."""
    from comic_b200 import engine, weights as wts
    lib = engine.load_library(lib_path)
    tab = lib.comic_conv_table()
    for i, (_scope, k, s, cin, cout) in enumerate(wts.cnn_conv_list()):
        d = tab[i]
        assert (d.k, d.stride, d.c_in, d.c_out) == (k, s, cin, cout)


def test_no_cpu_fallback():
    """The product path fails loudly without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from comic_b200 import configuration as conf
    from comic_b200.engine import Engine, ComicError
    with pytest.raises(ComicError):
        Engine(conf.make_config())


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'comic-compact-image-captioning-with-attention_b200')
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'comic_oracle' not in src and 'inception_v1_oracle' not in src, f
