"""Whole-model drop-in for the reference's entry point.

``output_GPEMSR.py:5`` does ``from model.GPEMSR import GPEMSR`` and builds the model from ``option/*.yml`` (:36-43), loads
the stage-3 checkpoint with ``strict=True`` (:52) and calls ``model(LQ)`` per window (:63-124).  ``install()`` registers a
module named ``model.GPEMSR`` whose ``GPEMSR`` (and ``POD`` / ``ThreeDA``) are the sm_100a mirrors, so that import -- and
everything after it -- runs unchanged on the CUDA library; BasicSR is no longer imported at all.  The other ``model.*``
modules of the reference stay importable (only this one entry of ``sys.modules`` is replaced).
"""
from __future__ import annotations

import importlib
import sys
import types


def install():
    """Call once before ``from model.GPEMSR import GPEMSR``.  Returns the registered module."""
    from . import gpemsr as G
    try:                                            # the reference's own (empty) model/__init__.py when it is on sys.path
        pkg = importlib.import_module('model')
    except ImportError:
        pkg = types.ModuleType('model')
        pkg.__path__ = []
        sys.modules['model'] = pkg
    mod = types.ModuleType('model.GPEMSR')
    mod.__doc__ = 'gpemsr_b200 drop-in for model/GPEMSR.py (sm_100a kernels, inference only)'
    mod.GPEMSR, mod.POD, mod.ThreeDA = G.GPEMSR, G.POD, G.ThreeDA
    sys.modules['model.GPEMSR'] = mod
    setattr(pkg, 'GPEMSR', mod)
    return mod
