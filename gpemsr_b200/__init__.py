"""gpemsr_b200: B200-native (sm_100a) kernels for GPEMSR's inference hot path.

Host-side mirrors of the reference's operators for this path (same names, argument
meaning and error behaviour) over the C ABI in ``include/gpemsr_b200.h``:

  * ``flow_warp``                       -- BasicSR ``flow_warp`` (SpyNet's warping operator)
  * ``Codebook`` (``forward`` / ``inference_lr``)   -- model/codebook.py
  * ``Decoder`` (``forward`` / ``multi_scale_feat_calculate``) -- model/decoder.py
  * ``SRTail``                          -- model/GPEMSR.py:441-455
  * ``SpyNet``                          -- basicsr/archs/spynet_arch.py (7x7 convs + flow_warp, coarse to fine)
  * ``VGG19Slice1``                     -- model/VGG.py slice1 + the patch-similarity mask of model/GPEMSR.py:344-353
  * ``GPEMSR``                          -- model/GPEMSR.py (the whole model: reference fusion, POD, ThreeDA, SR tail)
  * ``Indexer16`` / ``Indexer8`` / ``lrGenerator16`` / ``lrGenerator8`` (inference methods) -- model/indexer.py, model/vqgan_indexer.py

The CUDA library is mandatory: nothing here falls back to PyTorch or the CPU.
"""
from ._lib import GpemsrError, LIB_PATH, kernel_launches, lib  # noqa: F401
from .flow_warp import flow_warp  # noqa: F401
from .codebook import Codebook, argmax_gather, logits_argmax_gather, vq_lookup  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .sr_tail import SRTail  # noqa: F401
from .indexer import Indexer8, Indexer16, lrGenerator8, lrGenerator16  # noqa: F401
from .spynet import SpyNet  # noqa: F401
from .vgg import VGG19Slice1  # noqa: F401
from .gpemsr import GPEMSR  # noqa: F401

__all__ = ['GPEMSR', 'flow_warp', 'Decoder', 'SRTail', 'Indexer16', 'Indexer8', 'lrGenerator16', 'lrGenerator8', 'SpyNet', 'VGG19Slice1', 'Codebook', 'vq_lookup', 'logits_argmax_gather', 'argmax_gather', 'GpemsrError', 'lib', 'kernel_launches', 'LIB_PATH']
