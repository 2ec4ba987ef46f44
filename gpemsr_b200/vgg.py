"""The VGG19 relu1_2 patch-similarity mask of ``GPEMSR.forward`` (model/GPEMSR.py:343-353 / 395-403, model/VGG.py) on the
sm_100a kernels -- SURVEY.md 8(f)-1.

The reference runs the WHOLE VGG19 on two HR-sized batches and keeps only ``relu1_2`` (slice1 = conv3x3(3->64) + ReLU +
conv3x3(64->64) + ReLU), cuts both feature maps into 16 x 16 patches, L2-normalises every patch vector (64*256 values) and
takes their dot product: one cosine similarity per patch.  Here

  * only slice1 is evaluated (the other four slices never reach an output of the network);
  * the three identical input channels (``img.expand(-1, 3, -1, -1)``) are folded into one (weights summed on the host);
  * conv1_1 runs on CUDA cores straight from the NCHW image into operand planes, conv1_2 on the tensor cores;
  * the patch reduction rides in conv1_2's epilogue for the second image, so only ONE of the two 64-channel HR feature
    maps ever reaches HBM and neither is ever unfolded.

Parameter names are those of ``VGG19.slice1`` (``slice1.0.*``, ``slice1.2.*``): ``load_reference_state_dict`` takes the
reference module's ``state_dict()`` (or torchvision's ``features.{0,2}.*``) and ignores the slices that are not needed.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from . import igemm as G


class VGG19Slice1(nn.Module):
    def __init__(self, precision='fp32'):
        super().__init__()
        self.slice1 = nn.Sequential()
        self.slice1.add_module('0', nn.Conv2d(3, 64, 3, 1, 1))
        self.slice1.add_module('1', nn.ReLU(inplace=True))
        self.slice1.add_module('2', nn.Conv2d(64, 64, 3, 1, 1))
        self.slice1.add_module('3', nn.ReLU(inplace=True))
        for p in self.parameters():
            p.requires_grad = False
        assert precision in ('fp32', 'bf16')
        self.split = 3 if precision == 'fp32' else 1
        self._plans = {}

    def load_reference_state_dict(self, sd):
        pick = {}
        for k, v in sd.items():
            for src, dst in (('slice1.0.', 'slice1.0.'), ('slice1.2.', 'slice1.2.'), ('features.0.', 'slice1.0.'), ('features.2.', 'slice1.2.')):
                if k.startswith(src):
                    pick[dst + k[len(src):]] = v
        return self.load_state_dict(pick, strict=True)

    # ------------------------------------------------------------------ CUDA path
    def _plan(self, n, h, w, device):
        key = (n, h, w, device.index)
        P = self._plans.get(key)
        if P is None:
            g = G.Geom(n, h, w, True)
            # fp32-faithful: the first image's features are kept as fp32 cells; one-pass bf16: as a bf16 plane (what the bf16
            # branch resolves anyway) -- half the bytes of the largest tensor of the forward, written once and read once
            P = dict(g=g, c1=G.Act(g, 64, device, f32=False, split=self.split),
                     ref=G.Act(g, 64, device, f32=self.split == 3, planes=self.split == 1, split=1), err=G.err_flag(device), wts={})
            self._plans[key] = P
        c0, c2 = self.slice1[0], self.slice1[2]                      # packed forms follow the live parameters (igemm.cached)
        # the 3 input channels are copies of one image: their weights are summed
        P['w1'], P['b1'] = G.cached(P['wts'], 'c1', (c0.weight, c0.bias), lambda: (
            c0.weight.detach().float().sum(1).reshape(64, 9).contiguous(), c0.bias.detach().float().contiguous()))
        P['w2'] = G.cached(P['wts'], 'c2', (c2.weight,), lambda: G.Weights(c2.weight, 'conv', split=self.split))
        return P

    def _conv1(self, P, img):
        g = P['g'].c
        c1 = P['c1']
        _lib.check(_lib.lib().gpemsr_conv3x3_c1_relu(_lib.ptr(img), _lib.ptr(P['w1']), _lib.ptr(P['b1']), 64, C.byref(g),
                                                     _lib.ptr(c1.hi), _lib.ptr(c1.lo), _lib.stream_ptr()))
        return c1

    @torch.no_grad()
    def relu1_2(self, img):
        """``VGG19(img.expand(-1, 3, -1, -1)).relu1_2`` for a one-channel batch: f32 [n, 64, h, w] (tests / debugging)."""
        img = self._check(img)
        n, _, h, w = img.shape
        P = self._plan(n, h, w, img.device)
        out = torch.empty(n, 64, h, w, dtype=torch.float32, device=img.device)
        G.igemm(self._conv1(P, img), P['w2'], P['err'], split=self.split, bias=self.slice1[2].bias.detach(), act=G.ACT_RELU,
                out_nchw=out, nchw_c=64)
        return out

    @torch.no_grad()
    def patch_similarity(self, ref_img, other_img, ksize=16):
        """cosine similarity of the relu1_2 features of the two one-channel image batches per ksize x ksize patch:
        ``sum(normalize(patches(vgg(ref))) * normalize(patches(vgg(other))), dim=1)`` reshaped to [n, 1, h/k, w/k]
        (model/GPEMSR.py:345-353)."""
        ref_img, other_img = self._check(ref_img), self._check(other_img)
        if ref_img.shape != other_img.shape:
            raise ValueError('patch_similarity: the two batches must have the same shape')
        n, _, h, w = ref_img.shape
        if h % ksize or w % ksize:
            raise _lib.GpemsrError(-6, f'patch_similarity: {h}x{w} is not a multiple of the {ksize}x{ksize} patch (the '
                                       'reflection-padded "same" case of extract_image_patches is not built)')
        P = self._plan(n, h, w, ref_img.device)
        b2 = self.slice1[2].bias.detach()
        G.igemm(self._conv1(P, ref_img), P['w2'], P['err'], split=self.split, bias=b2, act=G.ACT_RELU, out=P['ref'],
                out_planes=self.split == 1, out_f32=self.split == 3)
        npatch = n * (h // ksize) * (w // ksize)
        sums = P.get('sums')
        if sums is None:
            sums = P['sums'] = torch.empty(npatch, 3, dtype=torch.float32, device=ref_img.device)
        G.igemm(self._conv1(P, other_img), P['w2'], P['err'], split=self.split, bias=b2, act=G.ACT_RELU, o_geom=P['g'],
                patch_other=P['ref'].f32 if self.split == 3 else P['ref'].hi, patch_other_bf16=self.split == 1, patch_sums=sums,
                patch_size=ksize)
        mask = torch.empty(n, 1, h // ksize, w // ksize, dtype=torch.float32, device=ref_img.device)
        _lib.check(_lib.lib().gpemsr_patch_cosine(_lib.ptr(sums), npatch, 1e-12, _lib.ptr(mask), _lib.stream_ptr()))
        self._last_err = P['err']
        G.post_error_check(ref_img.device)
        return mask

    @torch.no_grad()
    def similarity_mask(self, ref_img, x_lr, scale):
        """The mask input of ``refmaskconv1`` (model/GPEMSR.py:344-353): ref_img f32 [n, 1, s*H, s*W] (the VQ decoder's
        image), x_lr f32 [n, 1, H, W] -> f32 [n, 1, s*H/16, s*W/16]; ``up_lr`` is the bilinear x`scale` upsampling
        (align_corners=False) of the LR frames."""
        x_lr = self._check(x_lr)
        n, _, h, w = x_lr.shape
        up = torch.zeros(n, 1, scale * h, scale * w, dtype=torch.float32, device=x_lr.device)
        _lib.check(_lib.lib().gpemsr_add_bilinear_base(_lib.ptr(x_lr), n, h, w, scale, _lib.ptr(up), _lib.stream_ptr()))
        return self.patch_similarity(ref_img, up)

    def check(self):
        G.check_pipeline(self._last_err)

    @staticmethod
    def _check(t):
        if not t.is_cuda:
            raise _lib.GpemsrError(-3, 'VGG19Slice1 needs CUDA tensors: there is no CPU fallback')
        if t.dim() != 4 or t.shape[1] != 1:
            raise ValueError('expected a one-channel batch [n, 1, h, w] (the reference expands it to 3 identical channels)')
        return t.float().contiguous()
