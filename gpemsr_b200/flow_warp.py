"""Host-side mirror of BasicSR's ``flow_warp`` (the operator SpyNet calls), on the sm_100a kernel.

Same name, argument meaning and error behaviour as
``basicsr.archs.arch_util.flow_warp(x, flow, interp_mode='bilinear', padding_mode='zeros',
align_corners=True)``, reached by the reference from ``model/GPEMSR.py:99-100`` via
``SpyNet.process`` (which passes ``'bilinear', 'border'``).
"""
from __future__ import annotations

import torch

from . import _lib

_PAD = {'zeros': 0, 'border': 1}
# how `2 * v / max(size - 1, 1)` is rounded: 'cuda' = the reference's own device path (ATen's CUDA true-divide kernel turns a
# Python-scalar divisor into a multiplication by the fp32 reciprocal), 'cpu' = ATen's CPU kernel (a true division).  Measured
# on the B200 against F.grid_sample on both devices: profiles/r02_flow_h3.json.
_COORD = {'cpu': 0, 'cuda': 1}
DEFAULT_COORD_FORM = 'cuda'


def flow_warp(x, flow, interp_mode='bilinear', padding_mode='zeros', align_corners=True, coord_form=None):
    """x: f32[n, c, h, w]; flow: f32[n, h, w, 2] (dx, dy in pixels) -> f32[n, c, h, w].
    coord_form (not a BasicSR argument): which device's rounding of the coordinate normalisation to reproduce."""
    assert x.size()[-2:] == flow.size()[1:3]      # same assertion as BasicSR
    if interp_mode != 'bilinear':
        raise NotImplementedError(f'interp_mode={interp_mode!r}: only bilinear is built (the reference uses no other)')
    if padding_mode not in _PAD:
        raise NotImplementedError(f'padding_mode={padding_mode!r}: only zeros/border are built')
    if not x.is_cuda:
        raise _lib.GpemsrError(-3, 'flow_warp needs CUDA tensors: there is no CPU fallback')
    if x.dtype != torch.float32 or flow.dtype != torch.float32:
        raise TypeError('flow_warp: fp32 only (the reference path is fp32)')
    n, c, h, w = x.shape
    if flow.shape[0] != n or flow.shape[3] != 2:
        raise ValueError(f'flow must be [n, h, w, 2], got {tuple(flow.shape)}')
    x = x.contiguous()
    flow = flow.contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.lib().gpemsr_flow_warp_ex(_lib.ptr(x), _lib.ptr(flow), n, c, h, w, _PAD[padding_mode],
                                              int(bool(align_corners)), _COORD[coord_form or DEFAULT_COORD_FORM], _lib.ptr(out),
                                              _lib.stream_ptr()))
    return out
