"""Host-side mirror of the SR tail of ``GPEMSR.forward`` (model/GPEMSR.py:441-455, layers :302-318) on the
sm_100a implicit-GEMM kernels:

    recon_trunk (10 x ResidualBlockNoBN) -> [upconvK -> PixelShuffle(2) -> LeakyReLU(0.1)] x 3 (x8) / x 4 (x16)
    -> HRconv + LeakyReLU -> conv_last -> + bilinear_upsample(x_center, scale, align_corners=False)

Parameter names are the reference's (``recon_trunk.{i}.conv{1,2}``, ``upconv{1..4}``, ``HRconv``, ``conv_last``) so the
matching slice of a stage-3 checkpoint loads with ``strict=True``.  Bias, ReLU / LeakyReLU, the residual add and
PixelShuffle are fused into the GEMM epilogues; conv_last writes NCHW directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import igemm as G

LRELU_SLOPE = 0.1        # model/GPEMSR.py:321


class ResidualBlockNoBN(nn.Module):          # parameter holder (BasicSR ResidualBlockNoBN: conv1, conv2, res_scale=1)
    def __init__(self, num_feat=64):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)


class SRTail(nn.Module):
    def __init__(self, nf=64, back_RBs=10, scale=8, precision='fp32'):
        super().__init__()
        assert scale in (8, 16)
        self.nf, self.scale = nf, scale
        self.precision = G.Precision(precision)
        self.recon_trunk = nn.Sequential(*[ResidualBlockNoBN(nf) for _ in range(back_RBs)])
        self.upconv1 = nn.Conv2d(nf, nf * 4, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(nf, 64 * 4, 3, 1, 1, bias=True)
        self.upconv3 = nn.Conv2d(64, 64 * 4, 3, 1, 1, bias=True)
        if scale == 16:
            self.upconv4 = nn.Conv2d(64, 64 * 4, 3, 1, 1, bias=True)
        self.HRconv = nn.Conv2d(64, 64, 3, 1, 1, bias=True)
        self.conv_last = nn.Conv2d(64, 1, 3, 1, 1, bias=True)
        self._plans = {}

    @torch.no_grad()
    def forward(self, fea, x_center):
        """fea f32[B, nf, H, W] (ThreeDA output), x_center f32[B, 1, H, W] -> f32[B, 1, scale*H, scale*W]."""
        if not fea.is_cuda:
            from ._lib import GpemsrError
            raise GpemsrError(-3, 'SRTail needs CUDA tensors: there is no CPU fallback')
        G.poll_error(fea.device)
        n, c, h, w = fea.shape
        key = (n, h, w, fea.device.index)
        P = self._plans.get(key)
        if P is None:
            from .decoder import _Plan
            P = _Plan(self, n, h, w, fea.device)
            self._plans[key] = P
        g = G.Geom(n, h, w, True)
        cur = P.act('in', g, c, f32=True)
        G.pack_nchw(fea.float(), cur)
        out = self._tail(P, cur, x_center)
        G.post_error_check(fea.device)
        return out

    def _tail(self, P, cur, x_center):
        """The tail from an activation already in the internal format (fp32 master + planes): used by ``forward`` and by the
        whole-model mirror ``gpemsr_b200.GPEMSR``."""
        g = cur.geom
        n = g.n
        t = P.act('trunk.t', g, self.nf, f32=False)
        pp = [P.act('trunk.a', g, self.nf, f32=True), P.act('trunk.b', g, self.nf, f32=True)]
        for i, rb in enumerate(self.recon_trunk):          # x + conv2(relu(conv1(x)))
            G.igemm(cur, P.weights(f'tail.rt{i}.1', rb.conv1.weight, 'conv'), P.err, split=P.sp('tail.rt'), bias=rb.conv1.bias.detach(),
                    act=G.ACT_RELU, out=t, out_f32=False)
            nxt = pp[i % 2]
            G.igemm(t, P.weights(f'tail.rt{i}.2', rb.conv2.weight, 'conv'), P.err, split=P.sp('tail.rt'), bias=rb.conv2.bias.detach(),
                    residual=cur.f32, out=nxt)
            cur = nxt
        ups = [self.upconv1, self.upconv2, self.upconv3] + ([self.upconv4] if self.scale == 16 else [])
        for i, conv in enumerate(ups):                     # lrelu(pixel_shuffle(conv(x)))
            og = G.Geom(n, cur.geom.h * 2, cur.geom.w * 2, True)
            out = P.act(f'up{i}', og, 64, f32=False)
            G.igemm(cur, P.weights(f'tail.up{i}', conv.weight, 'conv'), P.err, split=P.sp(f'tail.up{i}'), bias=conv.bias.detach(), act=G.ACT_LRELU,
                    slope=LRELU_SLOPE, out=out, up=2, pixel_shuffle=True, out_f32=False)
            cur = out
        hr = P.act('hr', cur.geom, 64, f32=False)
        G.igemm(cur, P.weights('tail.hr', self.HRconv.weight, 'conv'), P.err, split=P.sp('tail.hr'), bias=self.HRconv.bias.detach(),
                act=G.ACT_LRELU, slope=LRELU_SLOPE, out=hr, out_f32=False)
        out = torch.empty(n, 1, cur.geom.h, cur.geom.w, dtype=torch.float32, device=x_center.device)
        # conv_last (64 -> 1) + the bilinear base image (:450-455): nine taps as the columns of one 1x1 GEMM, then a nine-point
        # shifted sum on CUDA cores (an N = 16 tensor-pipe tile per tap ran at 1.4 % of peak)
        sp = P.sp('tail.last')
        wt = P.derived('tail.last', (self.conv_last.weight,), lambda: G.Weights(G.taps_as_columns(self.conv_last.weight), 'conv', split=sp))
        key = f'tail.last.taps{cur.geom.key()}'
        taps = P.bufs.get(key)
        if taps is None:
            taps = P.bufs[key] = G.TapCells(cur.geom, 1, x_center.device)
        G.conv3x3_few_outputs(hr, wt, taps, P.err, sp, self.conv_last.bias.detach(), out, 1, base=x_center.float().contiguous(),
                              base_scale=self.scale)
        self._last_plan = P
        return out

    def check(self):
        G.check_pipeline(self._last_plan.err)
