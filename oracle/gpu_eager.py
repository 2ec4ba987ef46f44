"""The reference's own device path: PyTorch eager on the GPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference runs on ``torch.device("cuda")`` (output_GPEMSR.py:44), so where
the CUDA and CPU library kernels of PyTorch round differently (ATen's CUDA true-divide by a Python scalar is a reciprocal
multiply; cuDNN / cuBLAS sum in another order than MKL) the parity oracle is THIS path, not the CPU run.  It is the same
functional restatement as everywhere (``oracle/gpemsr_model.py``, pinned bit-exact to the reference on CPU) evaluated on
``cuda`` tensors with the numerics pinned the way SURVEY.md section 7 step 1 asks:

    torch.backends.cudnn.allow_tf32 = False, torch.backends.cuda.matmul.allow_tf32 = False, cudnn.deterministic = True

``tf32=True`` selects PyTorch's defaults instead (cuDNN convolutions may use TF32, matmul stays fp32): what a user of the
reference gets out of the box, and the GPU baseline ``bench.py`` times beside the native path.
"""
from __future__ import annotations

import contextlib

import torch

from . import gpemsr_model as GM


@contextlib.contextmanager
def numerics(tf32=False):
    b = torch.backends
    old = (b.cudnn.allow_tf32, b.cuda.matmul.allow_tf32, b.cudnn.deterministic, b.cudnn.benchmark)
    try:
        if tf32:                      # PyTorch defaults
            b.cudnn.allow_tf32, b.cuda.matmul.allow_tf32, b.cudnn.deterministic, b.cudnn.benchmark = True, False, False, False
        else:
            b.cudnn.allow_tf32, b.cuda.matmul.allow_tf32, b.cudnn.deterministic, b.cudnn.benchmark = False, False, True, False
        yield
    finally:
        b.cudnn.allow_tf32, b.cuda.matmul.allow_tf32, b.cudnn.deterministic, b.cudnn.benchmark = old


def to_device(sd, device='cuda'):
    return {k: v.to(device) for k, v in sd.items()}


@torch.no_grad()
def forward(x, sd_dev, scale, tf32=False, **kw):
    """``GPEMSR.forward`` (model/GPEMSR.py:323-456) in PyTorch eager on the device of ``sd_dev``; returns device tensors."""
    dev = next(iter(sd_dev.values())).device
    with numerics(tf32):
        return GM.forward(x.to(dev), sd_dev, scale, **kw)


def flow_warp(x, flow, interp_mode='bilinear', padding_mode='zeros', align_corners=True):
    """BasicSR ``flow_warp`` evaluated by ATen's CUDA kernels (what SpyNet.process executes in the reference's run)."""
    from .flow_warp import flow_warp_torch
    assert x.is_cuda and flow.is_cuda
    with numerics(False):
        return flow_warp_torch(x, flow, interp_mode, padding_mode, align_corners)
