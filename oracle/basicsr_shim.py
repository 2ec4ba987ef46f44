"""Stand-in for the five BasicSR symbols the reference imports.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``model/GPEMSR.py:4,7,8`` imports
``basicsr.archs.arch_util`` (``make_layer``, ``DCNv2Pack``, ``ResidualBlockNoBN``)
and ``basicsr.archs.spynet_arch.SpyNet``; BasicSR (PyPI ``basicsr``, version not
pinned by the reference, v1.4.2 restated) is absent from this image, so the
unmodified reference ``model/GPEMSR.py`` can only be *run* (to produce the golden
vectors in tests/golden/) with these definitions installed under the module
names it expects.  Only the published behaviour is restated; parameter names
match upstream so reference checkpoints would load.  Parity at this boundary is
unpinned (no copy of BasicSR is available offline).

``install()`` registers the modules in ``sys.modules`` and neutralises the
hard-coded checkpoint loads (``model/VGG.py:11-12``, ``model/GPEMSR.py:275-284``,
SpyNet's ``load_path``) so a random-init model can be constructed.
"""
from __future__ import annotations

import math
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from .flow_warp import flow_warp_torch


def make_layer(block, n, **kw):
    return nn.Sequential(*[block(**kw) for _ in range(n)])


class ResidualBlockNoBN(nn.Module):
    def __init__(self, num_feat=64, res_scale=1, pytorch_init=False):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return x + self.conv2(self.relu(self.conv1(x))) * self.res_scale


class DCNv2Pack(nn.Module):
    """Modulated deformable conv whose offsets/masks come from a second feature map."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=True):
        super().__init__()
        k = kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, k, k))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels * k * k)
        nn.init.uniform_(self.weight, -bound, bound)
        self.conv_offset = nn.Conv2d(in_channels, deformable_groups * 3 * k * k, k, stride, padding,
                                     dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x, feat):
        import torchvision
        o1, o2, m = torch.chunk(self.conv_offset(feat), 3, dim=1)
        return torchvision.ops.deform_conv2d(x, torch.cat((o1, o2), 1), self.weight, self.bias,
                                             self.stride, self.padding, self.dilation, torch.sigmoid(m))


class _SpyLevel(nn.Module):
    def __init__(self):
        super().__init__()
        chans = [8, 32, 64, 32, 16, 2]
        mods = []
        for i in range(5):
            mods.append(nn.Conv2d(chans[i], chans[i + 1], 7, 1, 3))
            if i < 4:
                mods.append(nn.ReLU(inplace=False))
        self.basic_module = nn.Sequential(*mods)

    def forward(self, t):
        return self.basic_module(t)


class SpyNet(nn.Module):
    """Six-level coarse-to-fine flow estimator; each level warps the support image
    with ``flow_warp(..., 'bilinear', 'border')`` -- the hot-path operator a-5."""

    flow_warp = staticmethod(flow_warp_torch)     # tests swap this to capture/replace the operator

    def __init__(self, load_path=None):
        super().__init__()
        self.basic_module = nn.ModuleList([_SpyLevel() for _ in range(6)])
        self.register_buffer('mean', torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def process(self, ref, supp):
        ref = [(ref - self.mean) / self.std]
        supp = [(supp - self.mean) / self.std]
        for _ in range(5):
            ref.insert(0, F.avg_pool2d(ref[0], 2, 2, count_include_pad=False))
            supp.insert(0, F.avg_pool2d(supp[0], 2, 2, count_include_pad=False))
        flow = ref[0].new_zeros([ref[0].size(0), 2, ref[0].size(2) // 2, ref[0].size(3) // 2])
        for level in range(len(ref)):
            up = F.interpolate(flow, scale_factor=2, mode='bilinear', align_corners=True) * 2.0
            if up.size(2) != ref[level].size(2):
                up = F.pad(up, [0, 0, 0, 1], mode='replicate')
            if up.size(3) != ref[level].size(3):
                up = F.pad(up, [0, 1, 0, 0], mode='replicate')
            warped = type(self).flow_warp(supp[level], up.permute(0, 2, 3, 1),
                                          interp_mode='bilinear', padding_mode='border')
            flow = self.basic_module[level](torch.cat([ref[level], warped, up], 1)) + up
        return flow

    def forward(self, ref, supp):
        h, w = ref.size(2), ref.size(3)
        wf = int(math.floor(math.ceil(w / 32.0) * 32.0))
        hf = int(math.floor(math.ceil(h / 32.0) * 32.0))
        ref = F.interpolate(ref, size=(hf, wf), mode='bilinear', align_corners=False)
        supp = F.interpolate(supp, size=(hf, wf), mode='bilinear', align_corners=False)
        flow = F.interpolate(self.process(ref, supp), size=(h, w), mode='bilinear', align_corners=False)
        flow[:, 0, :, :] *= float(w) / float(wf)
        flow[:, 1, :, :] *= float(h) / float(hf)
        return flow


def install(reference_root):
    """Make ``import model.GPEMSR`` work for the tree at ``reference_root``
    (``.../GPEMSR-CREMI/GPEMSR``).  Returns the imported ``model.GPEMSR`` module."""
    pkg = types.ModuleType('basicsr')
    archs = types.ModuleType('basicsr.archs')
    au = types.ModuleType('basicsr.archs.arch_util')
    sp = types.ModuleType('basicsr.archs.spynet_arch')
    au.make_layer, au.ResidualBlockNoBN, au.DCNv2Pack, au.flow_warp = make_layer, ResidualBlockNoBN, DCNv2Pack, flow_warp_torch
    sp.SpyNet = SpyNet
    pkg.archs, archs.arch_util, archs.spynet_arch = archs, au, sp
    sys.modules.update({'basicsr': pkg, 'basicsr.archs': archs,
                        'basicsr.archs.arch_util': au, 'basicsr.archs.spynet_arch': sp})
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    # F9: hard-coded torch.load()s -> random init.  torch.load returns an empty
    # dict; the loads that are strict=True get a permissive load_state_dict.
    import importlib
    real_load = torch.load
    real_lsd = nn.Module.load_state_dict
    torch.load = lambda *a, **k: {}
    nn.Module.load_state_dict = lambda self, sd, strict=True, **k: real_lsd(self, sd, strict=False) if not sd else real_lsd(self, sd, strict=strict, **k)
    gp = importlib.import_module('model.GPEMSR')
    gp._oracle_restore = lambda: (setattr(torch, 'load', real_load),
                                  setattr(nn.Module, 'load_state_dict', real_lsd))
    return gp
