"""Host-side mirrors of the reference ``Indexer16`` / ``Indexer8`` (model/indexer.py) and of the stage-2 generators'
inference methods (model/vqgan_indexer.py) on the sm_100a kernels.

Module tree, constructor arguments and parameter names are the reference's (``input_layer.0``, ``feat_extract.{i}...``,
``output_layer.{i}``, ``embedding``), so ``stage2_x{8,16}.pth`` loads with ``strict=True``.  The nn.Modules only HOLD
parameters; ``forward`` (logits, model/indexer.py:51-55 / 98-102) and ``features`` run the CUDA path.  Inference only.

``Indexer*.forward`` materialises the [B, H, W, 1024] logits like the reference.  The generators below never do: they
feed ``features`` to ``Codebook.inference_from_feat`` (Linear + softmax + top-1 + gather fused).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import igemm as G
from .codebook import Codebook
from .decoder import Decoder, DownBlock, NonLocalBlock, ResidualBlock, UpBlock, _BlockNet


class _Indexer(_BlockNet):
    down_at = None                     # index of the scale whose last block is a DownBlock (model/indexer.py:27 / 77)
    tail_up = False                    # Indexer16 appends ResidualBlock + UpBlock for 4-entry channel lists (:31-34)

    def __init__(self, args, precision='fp32'):
        super().__init__()
        self.args = args
        self.channel_list = args['channel_list']
        self.input_layer = nn.Sequential(nn.Conv2d(args['im_channel'], self.channel_list[0], 3, 1, 1), nn.ReLU(inplace=True))
        self.num_res_blocks = args['num_resblock_per_scale']
        self.num_output_resblck = args['num_output_resblck']
        self.latent_dim = args['latent_dim']
        self.use_non_local = args['use_non_local']
        self._init_runner(precision)

        layers = []
        for i in range(len(self.channel_list) - 1):
            cin, cout = self.channel_list[i], self.channel_list[i + 1]
            for _ in range(self.num_res_blocks - 1):
                layers.append(ResidualBlock(cin, cin))
            layers.append(DownBlock(cin, cout) if i == self.down_at else ResidualBlock(cin, cout))
        if self.tail_up and len(self.channel_list) == 4:
            for _ in range(self.num_res_blocks - 1):
                layers.append(ResidualBlock(self.channel_list[-1], self.channel_list[-1]))
            layers.append(UpBlock(self.channel_list[-1], self.channel_list[-1]))
        if self.use_non_local:
            layers.append(NonLocalBlock(self.channel_list[-1]))
        self.feat_extract = nn.Sequential(*layers)

        layers = [ResidualBlock(self.channel_list[-1], self.channel_list[-1]) for _ in range(self.num_output_resblck)]
        layers.append(nn.Conv2d(self.channel_list[-1], self.latent_dim, 1))
        self.output_layer = nn.Sequential(*layers)
        self.embedding = nn.Linear(self.latent_dim, 1024)

    # ------------------------------------------------------------------ CUDA path
    def _trunk(self, x):
        """output_layer(feat_extract(input_layer(x))) in the internal format (model/indexer.py:52 / 99)."""
        P = self._plan_for(x)
        G.poll_error(x.device)
        n, c, h, w = x.shape
        g = G.Geom(n, h, w, True)
        xin = P.act('in', g, c, f32=False)
        G.pack_nchw(x.float(), xin)
        cur = P.act('input_layer.0.out', g, self.channel_list[0], f32=True)
        self._conv(P, 'input_layer.0', self.input_layer[0], xin, cur, act=G.ACT_RELU)
        for li, mod in enumerate(self.feat_extract):
            name = f'feat_extract.{li}'
            if isinstance(mod, ResidualBlock):
                cur = self._res_block(P, name, mod, cur)
            elif isinstance(mod, DownBlock):
                cur = self._down_block(P, name, mod, cur)
            elif isinstance(mod, UpBlock):
                cur = self._up_block(P, name, mod, cur, need_f32=True)
            else:
                cur = self._non_local(P, name, mod, cur)
        for i in range(self.num_output_resblck):
            cur = self._res_block(P, f'output_layer.{i}', self.output_layer[i], cur)
        return P, cur

    @torch.no_grad()
    def features(self, x):
        """NCHW fp32 [B, latent_dim, h, w]: the tensor the reference feeds to ``embedding`` after its permute."""
        P, cur = self._trunk(x)
        g = cur.geom
        feat = torch.empty(g.n, self.latent_dim, g.h, g.w, dtype=torch.float32, device=x.device)
        last = self.output_layer[self.num_output_resblck]
        G.igemm(cur, P.weights('output_layer.last', last.weight, 'conv'), P.err, split=P.sp('output_layer.last'), bias=last.bias.detach(),
                out_nchw=feat, nchw_c=self.latent_dim)
        G.post_error_check(x.device)
        return feat

    @torch.no_grad()
    def forward(self, x):
        """Logits [B, h, w, 1024] (model/indexer.py:51-55 / 98-102)."""
        P, cur = self._trunk(x)
        g = cur.geom
        last = self.output_layer[self.num_output_resblck]
        feat = P.act('output_layer.last.out', g, self.latent_dim, f32=False)
        G.igemm(cur, P.weights('output_layer.last', last.weight, 'conv'), P.err, split=P.sp('output_layer.last'), bias=last.bias.detach(),
                out=feat, out_f32=False)
        # the Linear over channels: rows = padded pixel rows, columns = codes; gather the interior rows afterwards
        k = self.embedding.out_features
        rows = g.n * g.r_img
        buf = torch.empty(rows, k, dtype=torch.float32, device=x.device)
        G.igemm(feat, P.weights('embedding', self.embedding.weight, 'linear'), P.err, split=P.sp('embedding'),
                bias=self.embedding.bias.detach(), out_rowmajor=buf, ld=k)
        logits = buf.view(g.n, g.r_img, k)[:, :(g.h + 2) * (g.w + 2)].view(g.n, g.h + 2, g.w + 2, k)[:, 1:-1, 1:-1]
        G.post_error_check(x.device)
        return logits.contiguous()


class Indexer16(_Indexer):                                     # model/indexer.py:6-55
    down_at = 4
    tail_up = True


class Indexer8(_Indexer):                                      # model/indexer.py:58-102
    down_at = 3


class _LrGenerator(nn.Module):
    """Inference surface of ``lrGenerator16`` / ``lrGenerator8`` (model/vqgan_indexer.py:19-50 / 62-93): the encoder is
    training-only and is not instantiated (load reference checkpoints of the sub-modules individually, as
    model/GPEMSR.py:275-284 does)."""

    indexer_cls = None
    key = None

    def __init__(self, args, precision='fp32'):
        super().__init__()
        prec = G.Precision(precision)
        self.indexer = self.indexer_cls(args[self.key], precision=prec.sub('indexer'))
        self.decoder = Decoder(args['Decoder'], precision=prec.sub('decoder'))
        self.codebook = Codebook(args['Codebook'])

    @torch.no_grad()
    def _lookup(self, imgs):
        feat = self.indexer.features(imgs)
        return self.codebook.inference_from_feat(feat, self.indexer.embedding.weight, self.indexer.embedding.bias)

    @torch.no_grad()
    def output_ref(self, imgs):                                # vqgan_indexer.py:26-31
        return self.decoder(self._lookup(imgs))

    @torch.no_grad()
    def ref_extract(self, imgs):                               # vqgan_indexer.py:44-48
        return self.decoder.multi_scale_feat_calculate(self._lookup(imgs))

    @torch.no_grad()
    def ref_extract_into(self, imgs, sinks):
        """``ref_extract`` whose four feature maps land as operand planes in the caller's buffers (``Decoder.multi_scale_feat_into``);
        returns the decoded reference images."""
        return self.decoder.multi_scale_feat_into(self._lookup(imgs), sinks)


class lrGenerator16(_LrGenerator):
    indexer_cls, key = Indexer16, 'Indexer16'


class lrGenerator8(_LrGenerator):
    indexer_cls, key = Indexer8, 'Indexer8'
