"""Host-side mirror of BasicSR's ``SpyNet`` (basicsr/archs/spynet_arch.py, v1.4.2; reached from model/GPEMSR.py:67, 99-100)
on the sm_100a kernels -- SURVEY.md 8(f)-3.

Six coarse-to-fine levels; each level upsamples the flow (bilinear, align_corners=True, x2), warps the support frame with
``flow_warp(..., 'bilinear', 'border')`` (row a-5), concatenates [ref, warped, flow] (8 channels) and runs the 5-layer
7 x 7 ``BasicModule``.  Here the 7 x 7 convolutions are 49-tap implicit GEMMs on a padded geometry with a 3-pixel zero ring,
the concatenation is fused with the operand packing, the normalisation / resizes / flow scaling are one bilinear kernel,
and the level's residual ``+ up`` is that same kernel in accumulate mode.  Parameter and buffer names are BasicSR's
(``basic_module.{l}.basic_module.{0,2,4,6,8}.*``, ``mean``, ``std``), so ``spynet_sintel_final-3d2a1287.pth`` loads with
``strict=True``.  Inference only; batched over (ref, supp) pairs.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import igemm as G
from .flow_warp import flow_warp


class BasicModule(nn.Module):                                  # spynet_arch.py BasicModule
    def __init__(self):
        super().__init__()
        chans = [8, 32, 64, 32, 16, 2]
        mods = []
        for i in range(5):
            mods.append(nn.Conv2d(chans[i], chans[i + 1], 7, 1, 3))
            if i < 4:
                mods.append(nn.ReLU(inplace=False))
        self.basic_module = nn.Sequential(*mods)


def _f32(v):
    return float(np.float32(v))


def resize_bilinear(x, ho, wo, align_corners, c_out=None, scale=None, rep=(0, 0), sub=None, div=None, mul=None, out=None,
                    accumulate=False, out_nhwc=None):
    """F.interpolate(x, bilinear) (+ per-channel (v - sub) / div * mul, channel broadcast, replicate tail, accumulate)."""
    n, c_in, h, w = x.shape
    c_out = c_in if c_out is None else c_out
    if align_corners:                                           # ATen area_pixel_compute_scale, computed in fp32
        nh, nw = rep[0] or ho, rep[1] or wo                     # natural output size (before a replicate tail)
        rh = _f32(np.float32(h - 1) / np.float32(nh - 1)) if nh > 1 else 0.0
        rw = _f32(np.float32(w - 1) / np.float32(nw - 1)) if nw > 1 else 0.0
    elif scale is not None:                                     # scale_factor given: ATen uses 1 / scale_factor
        rh = rw = _f32(np.float32(1.0) / np.float32(scale))
    else:
        rh, rw = _f32(np.float32(h) / np.float32(ho)), _f32(np.float32(w) / np.float32(wo))
    if out is None and out_nhwc is None:
        out = torch.empty(n, c_out, ho, wo, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().gpemsr_resize_bilinear(_lib.ptr(x), n, c_in, h, w, c_out, ho, wo, int(align_corners), rh, rw,
                                                 int(rep[0]), int(rep[1]), _lib.ptr(sub), _lib.ptr(div), _lib.ptr(mul),
                                                 int(accumulate), _lib.ptr(out), _lib.ptr(out_nhwc), _lib.stream_ptr()))
    return out


def avg_pool2(x):
    n, c, h, w = x.shape
    out = torch.empty(n, c, h // 2, w // 2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().gpemsr_avg_pool2(_lib.ptr(x), n * c, h, w, _lib.ptr(out), _lib.stream_ptr()))
    return out


class SpyNet(nn.Module):
    """Drop-in for ``basicsr.archs.spynet_arch.SpyNet`` (``load_path`` is accepted and ignored: load the state dict)."""

    def __init__(self, load_path=None, precision='fp32'):
        super().__init__()
        self.basic_module = nn.ModuleList([BasicModule() for _ in range(6)])
        self.register_buffer('mean', torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        assert precision in ('fp32', 'bf16')
        self.split = 3 if precision == 'fp32' else 1
        self._plans = {}

    # ------------------------------------------------------------------ CUDA path
    def _plan(self, n, h, w, device):
        key = (n, h, w, device.index)
        P = self._plans.get(key)
        if P is None:
            P = dict(err=G.err_flag(device), levels=[], two=torch.full((2,), 2.0, device=device), cache={})
            for lv in range(6):
                s = 5 - lv
                g = G.Geom(n, h >> s, w >> s, padded=3)
                acts = [G.Act(g, c, device, f32=False, split=self.split) for c in (8, 32, 64, 32, 16)]
                P['levels'].append(dict(g=g, acts=acts))
            self._plans[key] = P
        # packed forms follow the live parameters / buffers (igemm.cached)
        P['mean'], P['std'] = G.cached(P['cache'], 'norm', (self.mean, self.std), lambda: (
            self.mean.reshape(3).float().contiguous(), self.std.reshape(3).float().contiguous()))
        for lv in range(6):
            mods = self.basic_module[lv].basic_module
            P['levels'][lv]['wts'] = [G.cached(P['cache'], (lv, i), (mods[2 * i].weight,),
                                               lambda m=mods[2 * i]: G.Weights(m.weight, 'conv', split=self.split)) for i in range(5)]
        return P

    @torch.no_grad()
    def process(self, ref, supp):
        """spynet_arch.py SpyNet.process: ref, supp f32 [n, 3 or 1, h, w] with h, w multiples of 32 -> flow f32 [n, 2, h, w]."""
        n, _, h, w = ref.shape
        if h % 32 or w % 32:
            raise ValueError('SpyNet.process needs sizes that are multiples of 32 (SpyNet.forward resizes to them)')
        P = self._plan(n, h, w, ref.device)
        # (x - mean) / std, a one-channel frame broadcast to the three channels the network was trained on
        pyr_r = [resize_bilinear(ref, h, w, False, c_out=3, sub=P['mean'], div=P['std'])]
        pyr_s = [resize_bilinear(supp, h, w, False, c_out=3, sub=P['mean'], div=P['std'])]
        for _ in range(5):
            pyr_r.insert(0, avg_pool2(pyr_r[0]))
            pyr_s.insert(0, avg_pool2(pyr_s[0]))
        flow = torch.zeros(n, 2, pyr_r[0].shape[2] // 2, pyr_r[0].shape[3] // 2, dtype=torch.float32, device=ref.device)
        for lv in range(6):
            L = P['levels'][lv]
            hl, wl = pyr_r[lv].shape[2], pyr_r[lv].shape[3]
            nat = (2 * flow.shape[2], 2 * flow.shape[3])                       # natural size of the x2 upsampling
            rep = (nat[0] if nat[0] != hl else 0, nat[1] if nat[1] != wl else 0)      # F.pad(..., 'replicate') case
            up = torch.empty(n, 2, hl, wl, dtype=torch.float32, device=ref.device)
            up_nhwc = torch.empty(n, hl, wl, 2, dtype=torch.float32, device=ref.device)
            resize_bilinear(flow, hl, wl, True, rep=rep, mul=P['two'], out=up, out_nhwc=up_nhwc)
            warped = flow_warp(pyr_s[lv], up_nhwc, 'bilinear', 'border')
            x = L['acts'][0]
            g = L['g'].c
            _lib.check(_lib.lib().gpemsr_pack_concat3(_lib.ptr(pyr_r[lv]), 3, _lib.ptr(warped), 3, _lib.ptr(up), 2, C.byref(g),
                                                      _lib.ptr(x.hi), _lib.ptr(x.lo), _lib.stream_ptr()))
            mods = self.basic_module[lv].basic_module
            for i in range(4):
                G.igemm(x, L['wts'][i], P['err'], split=self.split, bias=mods[2 * i].bias.detach(), act=G.ACT_RELU,
                        out=L['acts'][i + 1], out_f32=False)
                x = L['acts'][i + 1]
            new_flow = torch.empty(n, 2, hl, wl, dtype=torch.float32, device=ref.device)
            G.igemm(x, L['wts'][4], P['err'], split=self.split, bias=mods[8].bias.detach(), out_nchw=new_flow, nchw_c=2)
            # flow = basic_module(...) + upsampled_flow : the same upsampling kernel in accumulate mode
            resize_bilinear(flow, hl, wl, True, rep=rep, mul=P['two'], out=new_flow, accumulate=True)
            flow = new_flow
        self._last_err = P['err']
        return flow

    @torch.no_grad()
    def forward(self, ref, supp):
        """spynet_arch.py SpyNet.forward: resize to multiples of 32, process, resize the flow back and rescale it."""
        if not ref.is_cuda:
            raise _lib.GpemsrError(-3, 'SpyNet needs CUDA tensors: there is no CPU fallback')
        G.poll_error(ref.device)
        ref, supp = ref.float().contiguous(), supp.float().contiguous()
        n, c, h, w = ref.shape
        wf, hf = int(math.floor(math.ceil(w / 32.0) * 32.0)), int(math.floor(math.ceil(h / 32.0) * 32.0))
        r = resize_bilinear(ref, hf, wf, False)
        s = resize_bilinear(supp, hf, wf, False)
        flow = self.process(r, s)
        key = ('scale', h, w, ref.device.index)
        scale = self._plans.get(key)
        if scale is None:                                      # flow[:, 0] *= w / wf ; flow[:, 1] *= h / hf
            scale = self._plans[key] = torch.tensor([float(w) / float(wf), float(h) / float(hf)], dtype=torch.float32,
                                                    device=ref.device)
        out = resize_bilinear(flow, h, w, False, mul=scale)
        G.post_error_check(ref.device)
        return out

    def check(self):
        G.check_pipeline(self._last_err)
