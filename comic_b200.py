"""Import shim: exposes the package directory
`comic-compact-image-captioning-with-attention_b200/` (not a legal Python
identifier) under the importable name `comic_b200`.

    import comic_b200
    from comic_b200 import rops, model
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        'comic-compact-image-captioning-with-attention_b200')
_spec = importlib.util.spec_from_file_location(
    'comic_b200', os.path.join(_PKG_DIR, '__init__.py'),
    submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['comic_b200'] = _mod
_spec.loader.exec_module(_mod)
