"""Deterministic random-init weights for the fixture generator and the tests.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The generator itself lives in ``gpemsr_b200/synth_weights.py`` (bench.py's
native arm needs the same synthetic parameters and must not import ``oracle/``); this module re-exports it so that
``oracle.make_golden`` and the tests keep one spelling.
"""
from gpemsr_b200.synth_weights import *  # noqa: F401,F403
from gpemsr_b200.synth_weights import (OrderedDict, _conv, _gn, _nonlocal, _resblock, codebook_spec, decoder_spec, fill,  # noqa: F401
                                       indexer_head_spec, indexer_spec, spynet_spec, tail_spec, vgg_slice1_spec)
