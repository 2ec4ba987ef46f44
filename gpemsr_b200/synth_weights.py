"""Deterministic random-init parameters for the modules of the hot path (no checkpoints exist offline).

Used by ``bench.py`` (both arms build the SAME synthetic weights from here), by ``__graft_entry__.smoke()``, by the tests
and by the golden-fixture generator (``oracle/weights.py`` re-exports this module).  Every tensor is drawn from a CPU
``torch.Generator`` seeded by (seed, crc32(parameter name)), i.e. independent of construction order and of the modules'
constructors.  Parameter names / shapes restate the reference constructors (``model/decoder.py:8-33``,
``model/blocks.py:8-59``, ``model/GPEMSR.py:302-318``, ``model/codebook.py:12-13``, ``model/indexer.py:6-96``,
``model/VGG.py:21-22``, BasicSR ``spynet_arch.py``) and are asserted equal to the reference modules' own ``state_dict()``
in ``oracle/make_golden.py``.  Pure parameter generation: nothing here computes any part of the hot path.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch


def _gen(seed, name):
    g = torch.Generator(device='cpu')
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def fill(spec, seed, gain=1.0):
    """spec: OrderedDict name -> (kind, shape).  Returns OrderedDict name -> f32 tensor."""
    out = OrderedDict()
    for name, (kind, shape) in spec.items():
        g = _gen(seed, name)
        if kind == 'conv':            # [Cout, Cin, kh, kw]; PyTorch default bound 1/sqrt(fan_in) * gain
            fan = shape[1] * shape[2] * shape[3]
            t = _uniform(shape, gain / math.sqrt(fan), g)
        elif kind == 'convT':         # [Cin, Cout, kh, kw]
            fan = shape[0] * shape[2] * shape[3] / 4.0
            t = _uniform(shape, gain / math.sqrt(fan), g)
        elif kind == 'linear':        # [K, D]
            t = _uniform(shape, gain / math.sqrt(shape[1]), g)
        elif kind == 'codebook':      # model/codebook.py:13  U(-1/K, 1/K)
            t = _uniform(shape, 1.0 / shape[0], g)
        elif kind == 'bias':
            t = _uniform(shape, 0.05, g)
        elif kind == 'gn_w':
            t = 1.0 + _uniform(shape, 0.2, g)
        elif kind == 'gn_b':
            t = _uniform(shape, 0.1, g)
        else:
            raise ValueError(kind)
        out[name] = t.contiguous()
    return out


def _conv(spec, p, cout, cin, k):
    spec[p + '.weight'] = ('conv', (cout, cin, k, k))
    spec[p + '.bias'] = ('bias', (cout,))


def _gn(spec, p, c):
    spec[p + '.weight'] = ('gn_w', (c,))
    spec[p + '.bias'] = ('gn_b', (c,))


def _resblock(spec, p, cin, cout):     # model/blocks.py:8-23
    _conv(spec, p + '.block.0', cout, cin, 3)
    _gn(spec, p + '.block.1', cout)
    _conv(spec, p + '.block.3', cout, cout, 3)
    _gn(spec, p + '.block.4', cout)
    if cin != cout:
        _conv(spec, p + '.channel_up', cout, cin, 1)


def _nonlocal(spec, p, c):             # model/blocks.py:50-59
    _gn(spec, p + '.gn', c)
    for n in ('q', 'k', 'v', 'proj_out'):
        _conv(spec, p + '.' + n, c, c, 1)


def decoder_spec(channel_list=(512, 256, 128, 64, 64), latent_dim=512, num_input_resblck=3,
                 num_res_blocks=1, use_non_local=True, im_channel=1):
    """Parameter names/shapes of ``Decoder`` -- model/decoder.py:8-33."""
    spec = OrderedDict()
    c0 = channel_list[0]
    _conv(spec, 'input_layer.0', c0, latent_dim, 1)
    for i in range(num_input_resblck):
        _resblock(spec, f'input_layer.{i + 1}', c0, c0)
    li = 0
    if use_non_local:
        _nonlocal(spec, f'feat_extract.{li}', c0)
        li += 1
    for i in range(len(channel_list) - 1):
        cin, cout = channel_list[i], channel_list[i + 1]
        for _ in range(num_res_blocks):
            _resblock(spec, f'feat_extract.{li}', cin, cin)
            li += 1
        spec[f'feat_extract.{li}.upblock.weight'] = ('convT', (cin, cout, 3, 3))
        spec[f'feat_extract.{li}.upblock.bias'] = ('bias', (cout,))
        li += 1
    _conv(spec, 'output_layer', im_channel, channel_list[-1], 3)
    return spec


def tail_spec(nf=64, back_rbs=10, scale=8):
    """Parameters of the SR tail -- model/GPEMSR.py:302-318."""
    spec = OrderedDict()
    for i in range(back_rbs):
        _conv(spec, f'recon_trunk.{i}.conv1', nf, nf, 3)
        _conv(spec, f'recon_trunk.{i}.conv2', nf, nf, 3)
    _conv(spec, 'upconv1', nf * 4, nf, 3)
    _conv(spec, 'upconv2', 64 * 4, nf, 3)
    _conv(spec, 'upconv3', 64 * 4, 64, 3)
    if scale == 16:
        _conv(spec, 'upconv4', 64 * 4, 64, 3)
    _conv(spec, 'HRconv', 64, 64, 3)
    _conv(spec, 'conv_last', 1, 64, 3)
    return spec


def codebook_spec(num_codes=1024, latent_dim=512):
    return OrderedDict([('embedding.weight', ('codebook', (num_codes, latent_dim)))])


def indexer_head_spec(latent_dim=512, num_codes=1024):
    return OrderedDict([('embedding.weight', ('linear', (num_codes, latent_dim))),
                        ('embedding.bias', ('bias', (num_codes,)))])


def indexer_spec(variant=16, channel_list=(64, 64, 128, 256, 512), im_channel=1, num_res_blocks=2, num_output_resblck=3,
                 latent_dim=512, use_non_local=True, num_codes=1024):
    """Parameter names/shapes of ``Indexer16`` / ``Indexer8`` -- model/indexer.py:6-47 / 58-96."""
    spec = OrderedDict()
    _conv(spec, 'input_layer.0', channel_list[0], im_channel, 3)
    down_at = 4 if variant == 16 else 3
    li = 0
    for i in range(len(channel_list) - 1):
        cin, cout = channel_list[i], channel_list[i + 1]
        for _ in range(num_res_blocks - 1):
            _resblock(spec, f'feat_extract.{li}', cin, cin)
            li += 1
        if i == down_at:
            _conv(spec, f'feat_extract.{li}.downblock', cout, cin, 3)
        else:
            _resblock(spec, f'feat_extract.{li}', cin, cout)
        li += 1
    c = channel_list[-1]
    if variant == 16 and len(channel_list) == 4:
        for _ in range(num_res_blocks - 1):
            _resblock(spec, f'feat_extract.{li}', c, c)
            li += 1
        spec[f'feat_extract.{li}.upblock.weight'] = ('convT', (c, c, 3, 3))
        spec[f'feat_extract.{li}.upblock.bias'] = ('bias', (c,))
        li += 1
    if use_non_local:
        _nonlocal(spec, f'feat_extract.{li}', c)
        li += 1
    for i in range(num_output_resblck):
        _resblock(spec, f'output_layer.{i}', c, c)
    _conv(spec, f'output_layer.{num_output_resblck}', latent_dim, c, 1)
    spec['embedding.weight'] = ('linear', (num_codes, latent_dim))
    spec['embedding.bias'] = ('bias', (num_codes,))
    return spec


def vgg_slice1_spec():
    """``VGG19.slice1`` = torchvision vgg19.features[0:4] -- model/VGG.py:21-22."""
    spec = OrderedDict()
    _conv(spec, 'slice1.0', 64, 3, 3)
    _conv(spec, 'slice1.2', 64, 64, 3)
    return spec


def spynet_spec():
    """Parameters of BasicSR ``SpyNet`` (spynet_arch.py): six BasicModules of five 7x7 convolutions 8-32-64-32-16-2."""
    spec = OrderedDict()
    chans = [8, 32, 64, 32, 16, 2]
    for lv in range(6):
        for i in range(5):
            _conv(spec, f'basic_module.{lv}.basic_module.{2 * i}', chans[i + 1], chans[i], 7)
    return spec


def fill_state(shapes, seed, gain=3.0 ** 0.5, offset_gain=0.5):
    """Deterministic parameters for a whole module tree given only its ``state_dict`` names and shapes (``gpemsr_b200.GPEMSR``
    or the reference ``GPEMSR``: same names).  The kind of every tensor is read off its name / rank:

      5-D / 4-D ``weight``: convolution (``reffea_L*_conv1`` / ``upblock``: transposed), bound gain / sqrt(fan_in) (gain
        sqrt(3) = Kaiming-uniform keeps activations O(1) through the ~60-layer chain); ``conv_offset`` (zero-initialised in
        BasicSR) gets small random weights (``offset_gain``) so the deformable sampling is exercised with offsets of ~1 pixel;
      2-D: ``codebook.embedding.weight`` U(+-1/K) (model/codebook.py:13), else Linear;  1-D: GroupNorm affine or bias;
      ``spynet.mean`` / ``spynet.std``: the ImageNet constants BasicSR registers as buffers."""
    out = OrderedDict()
    for name, shape in shapes.items():
        shape = tuple(shape)
        g = _gen(seed, name)
        if name.endswith('spynet.mean'):
            t = torch.tensor([0.485, 0.456, 0.406]).view(shape)
        elif name.endswith('spynet.std'):
            t = torch.tensor([0.229, 0.224, 0.225]).view(shape)
        elif len(shape) >= 4:
            transposed = 'upblock' in name or 'reffea_L' in name
            fan = (shape[0] if transposed else shape[1]) * math.prod(shape[2:])
            if transposed:
                fan /= 4.0
            gn = offset_gain if 'conv_offset' in name else 1.0 if 'spynet' in name else gain
            t = _uniform(shape, gn / math.sqrt(fan), g)
        elif len(shape) == 2:
            t = _uniform(shape, 1.0 / shape[0], g) if name.endswith('codebook.embedding.weight') else _uniform(shape, 1.0 / math.sqrt(shape[1]), g)
        elif name.endswith('.weight'):                      # GroupNorm scale
            t = 1.0 + _uniform(shape, 0.2, g)
        elif name.endswith('.bias') and len(shapes.get(name[:-4] + 'weight', (0, 0))) == 1:
            t = _uniform(shape, 0.1, g)                     # GroupNorm shift
        else:
            t = _uniform(shape, 0.05, g)
        out[name] = t.contiguous()
    return out
