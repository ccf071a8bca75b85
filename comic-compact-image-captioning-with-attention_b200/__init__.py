"""comic_b200 -- B200-native implementation of COMIC's caption-decoding hot path.

Host side (Python, mirrors the reference's call surface):
  configuration  Config / config.pkl / flag defaults   (common/configuration.py, src/train.py, src/infer.py)
  weights        W-table container + initialisers       
  engine         ctypes binding of libcomic_b200.so     (include/comic_b200.h)
  rops           rnn_decoder_* / attention cell mirror   (common/ops_rnn.py)
  model          CaptionModel / CaptionModel_SCST        (src/model.py, src/model_base.py)
Device side: csrc/*.cu, hand-written sm_100a CUDA behind a C ABI.
"""
from . import configuration, weights  # noqa: F401

__all__ = ['configuration', 'weights']
