"""Functional torch-CPU restatement of the whole ``GPEMSR.forward`` (model/GPEMSR.py:323-456) with ``POD`` (:64-150) and
``ThreeDA`` (:153-234).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``forward(x, sd, scale)`` takes the LR window ``x`` f32[B, N, 1, H, W] and a
``state_dict`` with the reference model's own parameter names (``GPEMSR(...).state_dict()``), and returns what the reference
returns: ``(out f32[B, 1, sH, sW], ref_img f32[B, N, 1, sH, sW])``.  With ``taps`` (a dict) it also records the intermediate
tensors the GPU parity tests compare stage by stage.  Pinned: ``oracle/make_golden.py`` runs the unmodified reference
``model/GPEMSR.py`` (through ``basicsr_shim``) on the same inputs and parameters; ``tests/test_oracle_golden.py`` checks this
restatement reproduces the committed outputs bit for bit.  The BasicSR pieces (SpyNet, DCNv2Pack, ResidualBlockNoBN, and
flow_warp inside SpyNet) are the restatements of ``basicsr_shim`` -- parity unpinned at that boundary, as everywhere.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import basicsr_shim
from . import ref_ops as R
from .ref_ops import LRELU_SLOPE, _sub


def _lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


def _conv(x, sd, name, stride=1, padding=1):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], stride, padding)


def _convT(x, sd, name):
    return F.conv_transpose2d(x, sd[name + '.weight'], sd[name + '.bias'], 2, 1, 1)


def _trunk(x, sd, prefix):
    i = 0
    while f'{prefix}.{i}.conv1.weight' in sd:
        x = R.residual_block_nobn(x, _sub(sd, f'{prefix}.{i}.'))
        i += 1
    return x


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)


def _dcn(x, feat, sd, name):
    """BasicSR DCNv2Pack.forward (arch_util.py v1.4.2): conv_offset -> chunk(3) -> cat(o1, o2), sigmoid(mask) -> deform_conv2d."""
    import torchvision
    o1, o2, m = torch.chunk(_conv(feat, sd, name + '.conv_offset'), 3, dim=1)
    return torchvision.ops.deform_conv2d(x, torch.cat((o1, o2), 1), sd[name + '.weight'], sd[name + '.bias'], 1, 1, 1,
                                         torch.sigmoid(m))


def _spynet(sd, prefix):
    net = basicsr_shim.SpyNet()
    own = net.state_dict()
    net.load_state_dict({k: sd[prefix + k] for k in own}, strict=True)
    return net.eval().to(sd[prefix + next(iter(own))].device)      # CPU, or the GPU-eager run of oracle/gpu_eager.py


def pod(nbr_fea_l, ref_fea_l, nbr, ref, sd, spynet, taps=None, tag=''):
    """``POD.forward`` -- model/GPEMSR.py:99-150 (``sd`` restricted to ``align_module.``)."""
    up4 = lambda t: F.interpolate(t, scale_factor=4, mode='bilinear', align_corners=False)
    flow = spynet(up4(nbr), up4(ref))                                   # :99-100 (the two calls are identical)
    f1 = [_conv(flow, sd, 'flowdsconv0_1', 4, 1)]                       # :101-106
    f2 = [_conv(flow, sd, 'flowdsconv0_2', 4, 1)]
    for lv in (1, 2):
        f1.append(_conv(f1[-1], sd, f'flowdsconv{lv}_1', 2, 1))
        f2.append(_conv(f2[-1], sd, f'flowdsconv{lv}_2', 2, 1))
    half = lambda t: F.interpolate(t, scale_factor=1 / 2, mode='bilinear', align_corners=False)
    nb = [nbr, half(nbr)]; nb.append(half(nb[1]))                       # :107-110
    rf = [ref, half(ref)]; rf.append(half(rf[1]))
    cat = lambda k: torch.cat([nbr_fea_l[k], ref_fea_l[k], f1[k], f2[k], nb[k], rf[k]], dim=1)
    o3 = _lrelu(_conv(cat(2), sd, 'L3_offset_conv1'))                   # :112-115
    o3 = _lrelu(_conv(o3, sd, 'L3_offset_conv2'))
    fea3 = _lrelu(_dcn(nbr_fea_l[2], o3, sd, 'L3_dcnpack'))
    o2 = _lrelu(_conv(cat(1), sd, 'L2_offset_conv1'))                   # :117-124
    o2 = _lrelu(_conv(torch.cat([o2, _up2(o3) * 2], dim=1), sd, 'L2_offset_conv2'))
    o2 = _lrelu(_conv(o2, sd, 'L2_offset_conv3'))
    fea2 = _dcn(nbr_fea_l[1], o2, sd, 'L2_dcnpack')
    fea2 = _lrelu(_conv(torch.cat([fea2, _up2(fea3)], dim=1), sd, 'L2_fea_conv'))
    o1 = _lrelu(_conv(cat(0), sd, 'L1_offset_conv1'))                   # :126-133
    o1 = _lrelu(_conv(torch.cat([o1, _up2(o2) * 2], dim=1), sd, 'L1_offset_conv2'))
    o1 = _lrelu(_conv(o1, sd, 'L1_offset_conv3'))
    fea1 = _dcn(nbr_fea_l[0], o1, sd, 'L1_dcnpack')
    fea1 = _conv(torch.cat([fea1, _up2(fea2)], dim=1), sd, 'L1_fea_conv')
    off = _lrelu(_conv(torch.cat([fea1, ref_fea_l[0]], dim=1), sd, 'cas_offset_conv1'))    # :135-138
    off = _lrelu(_conv(off, sd, 'cas_offset_conv2'))
    out = _lrelu(_dcn(fea1, off, sd, 'cas_dcnpack'))
    if taps is not None:
        taps.setdefault('pod.flow', []).append(flow)
        for k, v in (('o3', o3), ('fea3', fea3), ('o2', o2), ('fea2', fea2), ('o1', o1), ('fea1', fea1), ('off', off)):
            taps.setdefault('pod.' + k, []).append(v)
    return out


def three_da(aligned, sd, center, taps=None):
    """``ThreeDA.forward`` -- model/GPEMSR.py:181-234 (``sd`` restricted to ``ThreeDA.``)."""
    b, t, c, h, w = aligned.size()
    emb_ref = _conv(aligned[:, center].clone(), sd, 'temporal_attn1')
    emb = _conv(aligned.view(-1, c, h, w), sd, 'temporal_attn2').view(b, t, -1, h, w)
    corr = [torch.sum(emb[:, i] * emb_ref, 1).unsqueeze(1) for i in range(t)]
    prob = torch.sigmoid(torch.cat(corr, dim=1)).unsqueeze(2).expand(b, t, c, h, w).contiguous().view(b, -1, h, w)
    al = aligned.view(b, -1, h, w) * prob
    feat = _lrelu(_conv(al, sd, 'feat_fusion', 1, 0))
    c3 = lambda n: F.conv3d(al.view(b, t, -1, h, w), sd[n + '.weight'], sd[n + '.bias'])
    f1 = _lrelu(_conv(_lrelu(c3('conv3D_1')).view(b, -1, h, w), sd, 'conv3D_fusion_1', 1, 0))
    f2 = _lrelu(_conv(_lrelu(c3('conv3D_2')).view(b, -1, h, w), sd, 'conv3D_fusion_2', 1, 0))
    feat = feat + f1
    f3 = _conv(feat, sd, 'conv2D_fusion_3', 1, 0)
    pool = lambda v: torch.cat([F.max_pool2d(v, 3, 2, 1), F.avg_pool2d(v, 3, 2, 1)], dim=1)
    attn = _lrelu(_conv(al, sd, 'spatial_attn1', 1, 0))
    attn = _lrelu(_conv(pool(attn), sd, 'spatial_attn2', 1, 0))
    lvl = _lrelu(_conv(attn, sd, 'spatial_attn_l1', 1, 0))
    lvl = _lrelu(_conv(pool(lvl), sd, 'spatial_attn_l2'))
    lvl = _up2(_lrelu(_conv(lvl, sd, 'spatial_attn_l3')))
    attn = _lrelu(_conv(attn, sd, 'spatial_attn3')) + lvl
    attn = _up2(_lrelu(_conv(attn, sd, 'spatial_attn4', 1, 0)))
    attn = _conv(attn, sd, 'spatial_attn5')
    add = _conv(_lrelu(_conv(attn, sd, 'spatial_attn_add1', 1, 0)), sd, 'spatial_attn_add2', 1, 0)
    out = feat * torch.sigmoid(attn) * 2 + add + f2 + f3
    if taps is not None:
        taps.update({'tda.al': al, 'tda.feat': feat, 'tda.f2': f2, 'tda.attn': attn, 'tda.add': add})
    return out


def forward(x, sd, scale, taps=None, idx_override=None, logits_out=None):
    """``GPEMSR.forward`` -- model/GPEMSR.py:323-456 with the option/*.yml configuration (w_ref, POD, ThreeDA)."""
    B, N, C, H, W = x.size()
    center = N // 2
    xf = x.view(-1, C, H, W)
    x_center = x[:, center].contiguous()
    L1 = _trunk(_lrelu(_conv(xf, sd, 'conv_first')), sd, 'feature_extraction')                 # :329-330
    lr = [L1]                                                                                   # :335-340 / 381-384
    for k in (2, 3, 4)[:(3 if scale == 16 else 2)]:
        lr.append(_lrelu(_convT(lr[-1], sd, f'reffea_L{k}_conv1')))
    lr = lr[::-1]                                                                               # finest first
    gen = 'refmodel.'
    dec_feats, _ = R.ref_extract(xf, _sub(sd, gen + 'indexer.'), sd[gen + 'codebook.embedding.weight'], _sub(sd, gen + 'decoder.'),
                                 idx_override=idx_override, logits_out=logits_out)     # (test hooks, see ref_ops.ref_extract)
    ref_img = dec_feats[-1]                                                                     # :342 / 385
    dec = dec_feats[:-1][::-1]                                                                  # ref_x2, ref_x4, ref_x8(, ref_x16)
    mask = R.similarity_mask(ref_img, xf, _sub(sd, 'vgg.'), scale)                              # :344-353 / 387-396
    for i in (1, 2, 3):
        mask = _lrelu(_conv(mask, sd, f'refmaskconv{i}'))
    mask = torch.sigmoid(mask)                                                                  # :357 / 400
    J = len(lr)
    carried = None
    for j in range(J):                                                                          # :360-376 / 403-414
        parts = [lr[j], dec[j]] + ([carried] if carried is not None else [])
        r = _trunk(_conv(torch.cat(parts, dim=1), sd, f'reffusionconv{j + 1}'), sd, f'fusion_fea_block{j + 1}')
        s = 8 >> j
        r = r * (F.interpolate(mask, scale_factor=s, mode='bilinear', align_corners=False) if s > 1 else mask)
        if taps is not None:
            taps[f'fusion.r{j}'] = r
        if j < J - 1:
            carried = _conv(r if carried is None else torch.cat((r, carried), dim=1), sd, f'down_fea_conv{j + 1}', 2, 1)
    L1 = _conv(torch.cat((r, carried, L1), dim=1), sd, 'reduce_dim_conv', 1, 0)                 # :377-378 / 416-417
    L2 = _lrelu(_conv(_lrelu(_conv(L1, sd, 'fea_L2_conv1', 2, 1)), sd, 'fea_L2_conv2'))         # :421-425
    L3 = _lrelu(_conv(_lrelu(_conv(L2, sd, 'fea_L3_conv1', 2, 1)), sd, 'fea_L3_conv2'))
    L1v, L2v, L3v = L1.view(B, N, -1, H, W), L2.view(B, N, -1, H // 2, W // 2), L3.view(B, N, -1, H // 4, W // 4)
    ref_l = [L1v[:, center].clone(), L2v[:, center].clone(), L3v[:, center].clone()]
    spynet = _spynet(sd, 'align_module.spynet.')
    sd_pod = _sub(sd, 'align_module.')
    aligned = [pod([L1v[:, i].clone(), L2v[:, i].clone(), L3v[:, i].clone()], ref_l, x[:, i], x_center, sd_pod, spynet, taps)
               for i in range(N)]                                                               # :432-438
    aligned = torch.stack(aligned, dim=1)
    fea = three_da(aligned, _sub(sd, 'ThreeDA.'), center, taps)                                 # :439-440
    out = R.sr_tail(fea, x_center, sd, scale, back_rbs=sum(1 for k in sd if k.startswith('recon_trunk.') and k.endswith('conv1.weight')))
    if taps is not None:
        taps.update({'L1_fea0': lr[-1], 'mask': mask, 'L1_fea': L1, 'L2_fea': L2, 'L3_fea': L3, 'aligned': aligned, 'fea': fea})
    return out, ref_img.view(B, N, C, H * scale, W * scale)
