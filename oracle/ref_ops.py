"""Functional torch-CPU restatement of the reference modules on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function takes plain
tensors plus a ``state_dict``-style mapping that uses the reference's own
parameter names, so a reference checkpoint can be fed in unchanged.  All
paths below are relative to ``GPEMSR-CREMI/GPEMSR/`` in the reference tree.

The arithmetic is fp32 on CPU with the same ATen operators the reference
calls, in the same order, so the restatement is bit-identical to the reference
modules on CPU (checked by ``oracle/make_golden.py`` and
``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GN_GROUPS = 32      # model/blocks.py:5-6  Normalize(): GroupNorm(32, C, eps=1e-6, affine)
GN_EPS = 1e-6
LRELU_SLOPE = 0.1   # model/GPEMSR.py:321


def _sub(sd, prefix):
    """View of ``sd`` restricted to keys under ``prefix`` (prefix stripped)."""
    if not prefix:
        return sd
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------- #
# a-1 / a-2: Codebook
# --------------------------------------------------------------------------- #
def codebook_forward(z, emb, beta=1.0):
    """``Codebook.forward`` -- model/codebook.py:15-32.

    z: f32[B, D, H, W]; emb: f32[K, D] (``embedding.weight``).
    Returns (z_q f32[B, D, H, W], idx int64[B*H*W], loss f32[]).
    Distances use the reference's association ``(|z|^2 + |e|^2) - 2 z.e`` (:19-21),
    argmin returns the first (lowest) index on ties (:23).
    """
    zl = z.permute(0, 2, 3, 1).contiguous()                      # :16
    zf = zl.view(-1, emb.shape[1])                               # :17
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(emb ** 2, dim=1) \
        - 2 * torch.matmul(zf, emb.t())                          # :19-21
    idx = torch.argmin(d, dim=1)                                 # :23
    zq = F.embedding(idx, emb).view(zl.shape)                    # :24
    loss = torch.mean((zq - zl) ** 2) + beta * torch.mean((zq - zl) ** 2)   # :26 (no_grad: detach is a no-op)
    zq = zl + (zq - zl)                                          # :28 straight-through, evaluated numerically
    zq = zq.permute(0, 3, 1, 2).contiguous()                     # :30
    return zq, idx, loss


def codebook_inference_lr(p, emb):
    """``Codebook.inference_lr`` -- model/codebook.py:34-43.

    p: f32[B, H, W, K] logits.  softmax -> top-1 -> embedding gather -> NCHW.
    Returns (z_q f32[B, D, H, W], idx int64[B*H*W]); the reference returns only
    z_q, the indices are exposed for the parity tests.
    """
    B, H, W, C = p.shape
    pf = p.reshape(B * H * W, C).contiguous()                    # :37
    soft = F.softmax(pf, dim=1)                                  # :38
    _, top = torch.topk(soft, 1, dim=1)                          # :39
    top = top.squeeze(1)                                         # :40
    zq = F.embedding(top, emb).view(B, H, W, -1)                 # :41
    return zq.permute(0, 3, 1, 2).contiguous(), top              # :42


def indexer_logits(feat, w, b):
    """Indexer head ``self.embedding(feat.permute(0,2,3,1))`` -- model/indexer.py:47,53 / 96,100.

    feat: f32[B, D, H, W]; w: f32[K, D]; b: f32[K] -> logits f32[B, H, W, K].
    """
    return F.linear(feat.permute(0, 2, 3, 1), w, b)


# --------------------------------------------------------------------------- #
# a-3: blocks + Decoder
# --------------------------------------------------------------------------- #
def group_norm(x, w, b):
    """``Normalize`` -- model/blocks.py:5-6."""
    return F.group_norm(x, GN_GROUPS, w, b, GN_EPS)


def residual_block(x, sd):
    """``ResidualBlock.forward`` -- model/blocks.py:8-29 (keys ``block.{0,1,3,4}``, ``channel_up``)."""
    h = F.conv2d(x, sd['block.0.weight'], sd['block.0.bias'], 1, 1)
    h = F.relu(group_norm(h, sd['block.1.weight'], sd['block.1.bias']))
    h = F.conv2d(h, sd['block.3.weight'], sd['block.3.bias'], 1, 1)
    h = F.relu(group_norm(h, sd['block.4.weight'], sd['block.4.bias']))
    if 'channel_up.weight' in sd:                                # :22-23, :26-27
        return F.conv2d(x, sd['channel_up.weight'], sd['channel_up.bias']) + h
    return x + h                                                 # :29


def up_block(x, sd):
    """``UpBlock.forward`` -- model/blocks.py:32-38: ConvTranspose2d(k3, s2, p1, op1)."""
    return F.conv_transpose2d(x, sd['upblock.weight'], sd['upblock.bias'], 2, 1, 1)


def non_local_block(x, sd):
    """``NonLocalBlock.forward`` -- model/blocks.py:61-83."""
    h = group_norm(x, sd['gn.weight'], sd['gn.bias'])            # :62
    q = F.conv2d(h, sd['q.weight'], sd['q.bias'])
    k = F.conv2d(h, sd['k.weight'], sd['k.bias'])
    v = F.conv2d(h, sd['v.weight'], sd['v.bias'])
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)                # :69-70
    k = k.reshape(b, c, hh * ww)
    v = v.reshape(b, c, hh * ww)
    attn = torch.bmm(q, k)                                       # :74  [b, T, T]
    attn = attn * (int(c) ** (-0.5))                             # :75
    attn = F.softmax(attn, dim=2)                                # :76
    attn = attn.permute(0, 2, 1)                                 # :77
    a = torch.bmm(v, attn).reshape(b, c, hh, ww)                 # :79-80
    a = F.conv2d(a, sd['proj_out.weight'], sd['proj_out.bias'])  # :81
    return x + a                                                 # :83


def _decoder_layout(sd, num_input_resblck, num_res_blocks, use_non_local, n_scales):
    """Index bookkeeping of ``Decoder.__init__`` -- model/decoder.py:15-31."""
    layers = []
    i = 0
    if use_non_local:
        layers.append(('nl', i)); i += 1
    for _ in range(n_scales):
        for _ in range(num_res_blocks):
            layers.append(('rb', i)); i += 1
        layers.append(('up', i)); i += 1
    return layers


def decoder_multi_scale(x, sd, *, num_input_resblck=3, num_res_blocks=1, use_non_local=True,
                        n_scales=4):
    """``Decoder.multi_scale_feat_calculate`` -- model/decoder.py:40-57.

    Returns the list the reference returns: the input of every UpBlock except
    that the reference's index test (:46/:51) selects the ResidualBlock output
    that precedes each UpBlock, then the final image.
    """
    h = F.conv2d(x, sd['input_layer.0.weight'], sd['input_layer.0.bias'])          # :16
    for i in range(num_input_resblck):                                             # :17-18
        h = residual_block(h, _sub(sd, f'input_layer.{i + 1}.'))
    layers = _decoder_layout(sd, num_input_resblck, num_res_blocks, use_non_local, n_scales)
    feats = []
    start = 0
    if use_non_local:                                                              # :42-43
        h = non_local_block(h, _sub(sd, 'feat_extract.0.'))
        start = 1
    for j, (kind, li) in enumerate(layers[start:]):                                # :44-47 / :49-52
        p = _sub(sd, f'feat_extract.{li}.')
        h = residual_block(h, p) if kind == 'rb' else up_block(h, p)
        if (j - num_res_blocks + 1) % (num_res_blocks + 1) == 0:
            feats.append(h)
    # NB: with use_non_local the reference loop (:44) stops one layer early
    # (range(len-1) over feat_extract[i+1]) -- that still visits every layer.
    feats.append(F.conv2d(h, sd['output_layer.weight'], sd['output_layer.bias'], 1, 1))   # :53
    return feats


def decoder_forward(x, sd, **kw):
    """``Decoder.forward`` -- model/decoder.py:37-38."""
    return decoder_multi_scale(x, sd, **kw)[-1]


# --------------------------------------------------------------------------- #
# a-4: SR tail of GPEMSR.forward
# --------------------------------------------------------------------------- #
def residual_block_nobn(x, sd, res_scale=1.0):
    """BasicSR ``ResidualBlockNoBN.forward`` (third-party, v1.4.2 arch_util.py):
    ``identity + conv2(relu(conv1(x))) * res_scale``; 3x3, stride 1, pad 1, bias."""
    out = F.conv2d(F.relu(F.conv2d(x, sd['conv1.weight'], sd['conv1.bias'], 1, 1)),
                   sd['conv2.weight'], sd['conv2.bias'], 1, 1)
    return x + out * res_scale


def sr_tail(fea, x_center, sd, scale, back_rbs=10):
    """``GPEMSR.forward`` lines 441-455 -- model/GPEMSR.py (layers :302-318).

    fea: f32[B, 64, H, W] (output of ThreeDA); x_center: f32[B, 1, H, W].
    """
    out = fea
    for i in range(back_rbs):                                                      # :441 recon_trunk
        out = residual_block_nobn(out, _sub(sd, f'recon_trunk.{i}.'))
    n_up = {8: 3, 16: 4}[scale]
    for i in range(1, n_up + 1):                                                   # :442-448
        out = F.conv2d(out, sd[f'upconv{i}.weight'], sd[f'upconv{i}.bias'], 1, 1)
        out = F.leaky_relu(F.pixel_shuffle(out, 2), LRELU_SLOPE)
    out = F.leaky_relu(F.conv2d(out, sd['HRconv.weight'], sd['HRconv.bias'], 1, 1), LRELU_SLOPE)   # :449
    out = F.conv2d(out, sd['conv_last.weight'], sd['conv_last.bias'], 1, 1)        # :450
    base = F.interpolate(x_center, scale_factor=scale, mode='bilinear', align_corners=False)   # :452/454
    return out + base                                                              # :455


# --------------------------------------------------------------------------- #
# the hot path chained the way lrGenerator*.ref_extract chains it
# --------------------------------------------------------------------------- #
def ref_extract_from_feat(feat, sd_indexer_head, emb, sd_decoder, **dec_kw):
    """``lrGenerator{8,16}.ref_extract`` minus the Indexer conv stack --
    model/vqgan_indexer.py:44-48 / 87-91: logits -> inference_lr -> multi-scale decoder."""
    logits = indexer_logits(feat, sd_indexer_head['embedding.weight'], sd_indexer_head['embedding.bias'])
    zq, idx = codebook_inference_lr(logits, emb)
    return decoder_multi_scale(zq, sd_decoder, **dec_kw), idx


# --------------------------------------------------------------------------- #
# the Indexer conv stack in front of a-2
# --------------------------------------------------------------------------- #
def down_block(x, sd):
    """``DownBlock.forward`` -- model/blocks.py:41-47 (Conv2d k3, s2, p1)."""
    return F.conv2d(x, sd['downblock.weight'], sd['downblock.bias'], 2, 1)


def indexer_features(x, sd):
    """``output_layer(feat_extract(input_layer(x)))`` of ``Indexer16`` / ``Indexer8`` -- model/indexer.py:51-52 / 98-99.

    The layer kinds are read off the parameter names (the module lists of :21-37 / 72-86 in order), so one function
    serves both variants and every channel list.  Returns feat f32[B, latent_dim, h, w].
    """
    h = F.relu(F.conv2d(x, sd['input_layer.0.weight'], sd['input_layer.0.bias'], 1, 1))      # :10-11
    li = 0
    while any(k.startswith(f'feat_extract.{li}.') for k in sd):
        p = _sub(sd, f'feat_extract.{li}.')
        if 'downblock.weight' in p:
            h = down_block(h, p)
        elif 'upblock.weight' in p:
            h = up_block(h, p)
        elif 'q.weight' in p:
            h = non_local_block(h, p)
        else:
            h = residual_block(h, p)
        li += 1
    i = 0
    while f'output_layer.{i}.block.0.weight' in sd:                                          # :40-43
        h = residual_block(h, _sub(sd, f'output_layer.{i}.'))
        i += 1
    return F.conv2d(h, sd[f'output_layer.{i}.weight'], sd[f'output_layer.{i}.bias'])


def indexer_forward(x, sd):
    """``Indexer*.forward`` -- model/indexer.py:51-55 / 98-102: logits f32[B, h, w, 1024]."""
    return indexer_logits(indexer_features(x, sd), sd['embedding.weight'], sd['embedding.bias'])


def ref_extract(imgs, sd_indexer, emb, sd_decoder, idx_override=None, logits_out=None, **dec_kw):
    """``lrGenerator{8,16}.ref_extract`` -- model/vqgan_indexer.py:44-48 / 87-91.

    Test hooks (not part of the reference): ``idx_override`` replaces the top-1 indices (so a comparison downstream of the
    DISCRETE lookup cannot be derailed by a near-tied pair of logits), ``logits_out`` (a list) receives the logits."""
    logits = indexer_forward(imgs, sd_indexer)
    if logits_out is not None:
        logits_out.append(logits)
    zq, idx = codebook_inference_lr(logits, emb)
    if idx_override is not None:
        B, H, W, _ = logits.shape
        idx = idx_override.view(-1)
        zq = F.embedding(idx, emb).view(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
    return decoder_multi_scale(zq, sd_decoder, **dec_kw), idx


# --------------------------------------------------------------------------- #
# VGG19 relu1_2 patch-similarity mask
# --------------------------------------------------------------------------- #
def vgg_relu1_2(x3, sd):
    """``VGG19.forward(X).relu1_2`` -- model/VGG.py:21-22, 39-40 (slice1 = conv, ReLU, conv, ReLU)."""
    h = F.relu(F.conv2d(x3, sd['slice1.0.weight'], sd['slice1.0.bias'], 1, 1))
    return F.relu(F.conv2d(h, sd['slice1.2.weight'], sd['slice1.2.bias'], 1, 1))


def image_patches(x, k):
    """``extract_image_patches(x, [k, k], [k, k], [1, 1], 'same')`` for sizes divisible by k (no padding is added then) --
    model/GPEMSR.py:14-60: [N, C*k*k, L]."""
    assert x.shape[2] % k == 0 and x.shape[3] % k == 0
    return F.unfold(x, kernel_size=k, dilation=1, padding=0, stride=k)


def patch_similarity(ref_img, other_img, sd, k=16):
    """model/GPEMSR.py:345-353: relu1_2 of both (expanded to 3 channels), 16x16 patches, normalize, dot -> [N, 1, h/k, w/k]."""
    n, _, h, w = ref_img.shape
    a = F.normalize(image_patches(vgg_relu1_2(ref_img.expand(-1, 3, -1, -1), sd), k), dim=1)
    b = F.normalize(image_patches(vgg_relu1_2(other_img.expand(-1, 3, -1, -1), sd), k), dim=1)
    mask = torch.sum(a.contiguous() * b.contiguous(), dim=1, keepdim=True)
    return mask.view(n, 1, h // k, w // k)


def similarity_mask(ref_img, x_lr, sd, scale, k=16):
    """model/GPEMSR.py:344-353 (16to1) / 395-403 (8to1): up_lr = bilinear x`scale` of the LR frames, then the patch similarity."""
    up_lr = F.interpolate(x_lr, scale_factor=scale, mode='bilinear', align_corners=False)
    return patch_similarity(ref_img, up_lr, sd, k)
