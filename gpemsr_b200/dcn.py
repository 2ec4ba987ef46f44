"""Host-side mirror of BasicSR's ``DCNv2Pack`` (basicsr/archs/arch_util.py v1.4.2; used by ``POD`` in model/GPEMSR.py:79-94,
123-150) on the sm_100a kernels -- SURVEY.md 8(f)-4.

``forward(x, feat)``: ``conv_offset(feat)`` (3x3 conv -> 3 * groups * 9 channels) gives per-(group, tap) offsets and mask
logits; the modulated deformable convolution (torchvision.ops.deform_conv2d) then runs as a deformable-im2col gather
(one 8-channel cell per deformable group) followed by ONE tensor-core GEMM over the 9 * C gathered channels.  Parameter
names are BasicSR's (``weight``, ``bias``, ``conv_offset.{weight,bias}``).  Built for the reference's configuration:
3x3, stride 1, padding 1, dilation 1, groups 1, in_channels == 8 * deformable_groups (nf = 64, groups = 8).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from . import igemm as G


class DCNv2Pack(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, groups=1,
                 deformable_groups=8, bias=True, precision='fp32'):
        super().__init__()
        if (kernel_size, stride, padding, dilation, groups) != (3, 1, 1, 1, 1) or in_channels != 8 * deformable_groups or not bias:
            raise _lib.GpemsrError(-6, 'DCNv2Pack: built for 3x3 / stride 1 / padding 1 / dilation 1 / groups 1 with '
                                       'in_channels == 8 * deformable_groups (the reference uses 64 channels, 8 groups)')
        self.in_channels, self.out_channels, self.deformable_groups = in_channels, out_channels, deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels * 9)
        nn.init.uniform_(self.weight, -bound, bound)
        self.conv_offset = nn.Conv2d(in_channels, deformable_groups * 3 * 9, 3, 1, 1, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)
        assert precision in ('fp32', 'bf16')
        self.split = 3 if precision == 'fp32' else 1
        self._plans = {}

    def _plan(self, n, h, w, device):
        key = (n, h, w, device.index)
        P = self._plans.get(key)
        if P is None:
            g = G.Geom(n, h, w, True)
            c = self.in_channels
            P = dict(g=g, err=G.err_flag(device), wts={},
                     x=G.Act(g, c, device, f32=True, planes=False), feat=G.Act(g, c, device, f32=False, split=self.split),
                     col=G.Act(g, 9 * c, device, f32=False, split=self.split))
            self._plans[key] = P
        return P

    def _weights(self, cache, name):
        """(conv_offset weights, main weights as a 1x1 GEMM over the gathered channels), re-packed when the parameters change."""
        c = self.in_channels
        w_off = G.cached(cache, name + '.off', (self.conv_offset.weight,),
                         lambda: G.Weights(self.conv_offset.weight, 'conv', split=self.split))
        # [co, ci, ky, kx] -> [co, (ky*3 + kx) * C + ci]: the channel order the gather writes
        w_main = G.cached(cache, name + '.main', (self.weight,),
                          lambda: G.Weights(self.weight.detach().permute(0, 2, 3, 1).reshape(self.out_channels, 9 * c).contiguous(),
                                            'linear', split=self.split))
        return w_off, w_main

    @torch.no_grad()
    def forward(self, x, feat):
        if not x.is_cuda:
            raise _lib.GpemsrError(-3, 'DCNv2Pack needs CUDA tensors: there is no CPU fallback')
        G.poll_error(x.device)
        n, c, h, w = x.shape
        P = self._plan(n, h, w, x.device)
        G.pack_nchw(x.float(), P['x'])
        G.pack_nchw(feat.float(), P['feat'])
        om = torch.empty(n, 3 * self.deformable_groups * 9, h, w, dtype=torch.float32, device=x.device)
        w_off, w_main = self._weights(P['wts'], 'dcn')
        G.igemm(P['feat'], w_off, P['err'], split=self.split, bias=self.conv_offset.bias.detach(), out_nchw=om,
                nchw_c=om.shape[1])
        g = P['g'].c
        col = P['col']
        _lib.check(_lib.lib().gpemsr_deform_im2col(_lib.ptr(P['x'].f32), C.byref(g), c, self.deformable_groups, _lib.ptr(om),
                                                   _lib.ptr(col.hi), _lib.ptr(col.lo), C.byref(g), _lib.stream_ptr()))
        out = torch.empty(n, self.out_channels, h, w, dtype=torch.float32, device=x.device)
        G.igemm(col, w_main, P['err'], split=self.split, bias=self.bias.detach(), out_nchw=out, nchw_c=self.out_channels)
        self._last_err = P['err']
        G.post_error_check(x.device)
        return out

    def run_acts(self, P, name, x_f32, g, feat, out, err, act=G.ACT_NONE, slope=0.0, c_off=0, out_f32=True, out_planes=True):
        """The same three launches on tensors already in the internal format (used by ``gpemsr_b200.GPEMSR``'s POD): x_f32 =
        fp32 master cells of the input (64 channels, geometry g), feat = Act with the offset features' operand planes; the
        result goes to channel slot c_off of Act `out` with `act` fused.  `P` is the caller's buffer / weight cache."""
        c = self.in_channels
        w_off, w_main = self._weights(P.wts, name)
        key = f'dcn.om{g.key()}'
        om = P.bufs.get(key)
        if om is None:
            om = P.bufs[key] = torch.empty(g.n, 3 * self.deformable_groups * 9, g.h, g.w, dtype=torch.float32, device=x_f32.device)
        col = P.act(f'dcn.col{g.key()}', g, 9 * c, f32=False)
        G.igemm(feat, w_off, err, split=self.split, bias=self.conv_offset.bias.detach(), out_nchw=om, nchw_c=om.shape[1])
        gc = g.c
        _lib.check(_lib.lib().gpemsr_deform_im2col(_lib.ptr(x_f32), C.byref(gc), c, self.deformable_groups, _lib.ptr(om),
                                                   _lib.ptr(col.hi), _lib.ptr(col.lo), C.byref(gc), _lib.stream_ptr()))
        G.igemm(col, w_main, err, split=self.split, bias=self.bias.detach(), act=act, slope=slope, out=out,
                c_off=c_off, out_f32=out_f32, out_planes=out_planes)

    def check(self):
        G.check_pipeline(self._last_err)
