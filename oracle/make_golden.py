"""Generate tests/golden/*.npz by running the REFERENCE's own modules on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the authoring container
only (needs /root/reference):

    python -m oracle.make_golden [--ref /root/reference/GPEMSR-CREMI/GPEMSR]

For every fixture the reference module is constructed from the reference's own
source file, its parameters are overwritten with ``oracle.weights.fill`` (after
asserting names and shapes agree with the restated spec), it is run under
``torch.no_grad()`` in fp32 on CPU, and inputs/outputs are stored.  Weights are
NOT stored when they can be regenerated from (spec, seed).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

from . import weights as W
from . import basicsr_shim
from .flow_warp import flow_warp_torch

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _load_into(module, sd):
    own = module.state_dict()
    assert list(own.keys()) == list(sd.keys()), (list(own.keys())[:5], list(sd.keys())[:5])
    for k in own:
        assert tuple(own[k].shape) == tuple(sd[k].shape), (k, own[k].shape, sd[k].shape)
    module.load_state_dict(sd, strict=True)


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **arrs)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def gen_codebook(ref):
    from model.codebook import Codebook
    # (1) small, generic
    cb = Codebook({'num_codebook_vectors': 64, 'latent_dim': 32, 'beta': 1})
    emb = rand((64, 32), 11)
    cb.embedding.weight.data.copy_(emb)
    z = rand((2, 32, 5, 7), 12)
    logits = rand((2, 5, 7, 64), 13)
    zq, idx, loss = cb(z)
    zq_lr = cb.inference_lr(logits)
    # (2) ties: integer-valued data (all arithmetic exact in fp32 and bf16) with duplicated code rows
    g = torch.Generator().manual_seed(14)
    emb_t = torch.randint(-3, 4, (64, 32), generator=g).float()
    emb_t[40] = emb_t[7]; emb_t[41] = emb_t[7]; emb_t[3] = emb_t[20]
    z_t = torch.randint(-3, 4, (2, 32, 5, 7), generator=g).float()
    z_t[0, :, 0, 0] = emb_t[7]; z_t[0, :, 0, 1] = emb_t[20]; z_t[1, :, 4, 6] = emb_t[41]
    cb.embedding.weight.data.copy_(emb_t)
    zq_t, idx_t, loss_t = cb(z_t)
    logits_t = torch.randint(-2, 3, (2, 5, 7, 64), generator=g).float()
    zq_lr_t = cb.inference_lr(logits_t)
    _save('codebook_small', emb=_np(emb), z=_np(z), logits=_np(logits), zq=_np(zq), idx=_np(idx),
          loss=_np(loss), zq_lr=_np(zq_lr), emb_t=_np(emb_t), z_t=_np(z_t), zq_t=_np(zq_t),
          idx_t=_np(idx_t), loss_t=_np(loss_t), logits_t=_np(logits_t), zq_lr_t=_np(zq_lr_t))
    # (3) reference shape 1024 x 512 (option/output_GPEMSR_x8.yml:45-46), default U(+-1/K) init, weights regenerated
    cb = Codebook({'num_codebook_vectors': 1024, 'latent_dim': 512, 'beta': 1})
    _load_into(cb, W.fill(W.codebook_spec(1024, 512), seed=21))
    z = rand((1, 512, 9, 11), 22)
    zq, idx, loss = cb(z)
    head = W.fill(W.indexer_head_spec(512, 1024), seed=23)
    feat = rand((1, 512, 9, 11), 24)
    logits = torch.nn.functional.linear(feat.permute(0, 2, 3, 1), head['embedding.weight'], head['embedding.bias'])
    zq_lr = cb.inference_lr(logits)
    soft = torch.softmax(logits.view(-1, 1024), 1)
    idx_lr = torch.topk(soft, 1, dim=1)[1].squeeze(1)
    _save('codebook_1024x512', z=_np(z), idx=_np(idx), loss=_np(loss), zq=_np(zq), feat=_np(feat),
          idx_lr=_np(idx_lr), zq_lr=_np(zq_lr), seeds=np.array([21, 23]))


def gen_decoder(ref):
    from model.decoder import Decoder
    # small-width decoder, every layer type present (non-local, residual w/ GN, UpBlock, output conv)
    cfg = dict(channel_list=[64, 64, 32, 32, 32], im_channel=1, num_resblock_per_scale=1,
               num_input_resblck=2, latent_dim=64, use_non_local=True)
    dec = Decoder(cfg).eval()
    _load_into(dec, W.fill(W.decoder_spec(cfg['channel_list'], 64, 2, 1, True, 1), seed=31))
    x = rand((2, 64, 4, 6), 32)
    with torch.no_grad():
        feats = dec.multi_scale_feat_calculate(x)
        img = dec(x)
    assert torch.equal(img, feats[-1])
    _save('decoder_small', x=_np(x), **{f'feat{i}': _np(f) for i, f in enumerate(feats)}, seed=np.array([31]))
    # reference-width decoder (option/output_GPEMSR_x8.yml:48-54) on a 4x4 latent
    cfg = dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1,
               num_input_resblck=3, latent_dim=512, use_non_local=True)
    dec = Decoder(cfg).eval()
    _load_into(dec, W.fill(W.decoder_spec(), seed=33))
    x = rand((1, 512, 4, 4), 34)
    with torch.no_grad():
        feats = dec.multi_scale_feat_calculate(x)
    _save('decoder_full_4x4', x=_np(x), **{f'feat{i}': _np(f) for i, f in enumerate(feats)}, seed=np.array([33]))


def gen_blocks(ref):
    from model.blocks import ResidualBlock, UpBlock, NonLocalBlock
    out = {}
    rb = ResidualBlock(32, 64).eval()
    spec = W.OrderedDict(); W._resblock(spec, 'rb', 32, 64)
    _load_into(rb, {k[3:]: v for k, v in W.fill(spec, 41).items()})
    x = rand((2, 32, 9, 10), 42)
    up = UpBlock(32, 64).eval()
    spec_u = W.OrderedDict([('upblock.weight', ('convT', (32, 64, 3, 3))), ('upblock.bias', ('bias', (64,)))])
    _load_into(up, W.fill(spec_u, 43))
    nl = NonLocalBlock(64).eval()
    spec_n = W.OrderedDict(); W._nonlocal(spec_n, 'nl', 64)
    _load_into(nl, {k[3:]: v for k, v in W.fill(spec_n, 44).items()})
    xn = rand((2, 64, 5, 6), 45)
    with torch.no_grad():
        out = dict(x=_np(x), rb=_np(rb(x)), up=_np(up(x)), xn=_np(xn), nl=_np(nl(xn)))
    _save('blocks_small', **out, seeds=np.array([41, 43, 44]))


def gen_tail(ref):
    gp = basicsr_shim.install(ref)
    import yaml
    for scale, lr in ((8, 16), (16, 16)):
        with open(os.path.join(ref, 'option', f'output_GPEMSR_x{scale}.yml')) as f:
            opt = yaml.safe_load(f)
        net = opt['network']
        torch.manual_seed(50 + scale)
        import torchvision
        model = gp.GPEMSR(ref_path_G=None, ref_path_Indexer=None, argref=net['argref'], nf=net['nf'],
                          nframes=net['nframes'], groups=net['groups'], front_RBs=net['front_RBs'],
                          back_RBs=net['back_RBs'], w_ref=net['w_ref'],
                          ref_fusion_feat_RBs=net['ref_fusion_feat_RBs'], align_mode=net['align_mode'],
                          fusion_mode=net['fusion_mode'], mode=net['mode'], scale=scale).eval()
        tail = W.fill(W.tail_spec(64, net['back_RBs'], scale), seed=60 + scale, gain=3.0 ** 0.5)
        own = model.state_dict()
        for k, v in tail.items():
            assert tuple(own[k].shape) == tuple(v.shape), k
        missing = [k for k in own if k.split('.')[0] in ('recon_trunk', 'upconv1', 'upconv2', 'upconv3', 'upconv4', 'HRconv', 'conv_last') and k not in tail]
        assert not missing, missing
        model.load_state_dict(tail, strict=False)
        cap = {}
        model.recon_trunk.register_forward_pre_hook(lambda m, a: cap.__setitem__('fea', a[0].clone()))
        x = torch.rand(1, 5, 1, lr, lr, generator=torch.Generator().manual_seed(70 + scale))
        with torch.no_grad():
            out, ref_img = model(x)
        _save(f'tail_x{scale}', fea=_np(cap['fea']), x_center=_np(x[:, 2]), out=_np(out),
              seed=np.array([60 + scale]), scale=np.array([scale]))
    gp._oracle_restore()


def gen_flow_warp(ref):
    x = rand((2, 3, 20, 24), 81)
    flow = rand((2, 20, 24, 2), 82, 3.0)
    arrs = dict(x=_np(x), flow=_np(flow))
    for pm in ('border', 'zeros'):
        arrs['out_' + pm] = _np(flow_warp_torch(x, flow, 'bilinear', pm))
    # SpyNet-pyramid-like shape: 3 channels, level sizes 4..128
    x2 = rand((1, 3, 64, 64), 83)
    f2 = rand((1, 64, 64, 2), 84, 1.5)
    arrs.update(x2=_np(x2), flow2=_np(f2), out2_border=_np(flow_warp_torch(x2, f2, 'bilinear', 'border')))
    _save('flow_warp_small', **arrs)


def gen_indexer(ref):
    from model.indexer import Indexer16, Indexer8
    out = {}
    # small widths, every layer kind: channel-changing ResidualBlocks, NonLocalBlock, the DownBlock of Indexer8 (i == 3, on an
    # odd-sized input) and the ResidualBlock + UpBlock tail Indexer16 appends for 4-entry channel lists
    cases = [('i16', Indexer16, 16, [32, 32, 64, 64, 64], (2, 1, 6, 7), 61),
             ('i8', Indexer8, 8, [32, 32, 64, 64, 64], (2, 1, 9, 7), 63),
             ('i16up', Indexer16, 16, [32, 64, 64, 64], (1, 1, 5, 6), 65)]
    for tag, cls, variant, cl, shape, seed in cases:
        cfg = dict(channel_list=cl, im_channel=1, num_resblock_per_scale=2, num_output_resblck=1, latent_dim=64, use_non_local=True)
        m = cls(cfg).eval()
        _load_into(m, W.fill(W.indexer_spec(variant, cl, 1, 2, 1, 64, True), seed=seed))
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.rand(shape, generator=g)                      # EM intensities are in [0, 1]
        with torch.no_grad():
            feat = m.output_layer(m.feat_extract(m.input_layer(x)))
            logits = m(x)
        out.update({f'{tag}_x': _np(x), f'{tag}_feat': _np(feat), f'{tag}_logits': _np(logits)})
    _save('indexer_small', **out, seeds=np.array([61, 63, 65]))


def gen_vgg_mask(ref):
    """The mask lines of the reference forward (model/GPEMSR.py:344-353) executed with the reference's own ``VGG19``
    (model/VGG.py, random-init through the shim that neutralises its checkpoint load) and ``extract_image_patches``."""
    import torch.nn.functional as F
    M = basicsr_shim.install(ref)
    from model.VGG import VGG19
    vgg = VGG19().eval()
    M._oracle_restore()
    sd = W.fill(W.vgg_slice1_spec(), seed=81)
    own = vgg.slice1.state_dict()
    assert list(own.keys()) == ['0.weight', '0.bias', '2.weight', '2.bias']
    vgg.slice1.load_state_dict({k[len('slice1.'):]: v for k, v in sd.items()}, strict=True)
    g = torch.Generator().manual_seed(82)
    x = torch.rand(3, 1, 4, 6, generator=g)                                # LR frames
    ref_img = torch.rand(3, 1, 64, 96, generator=g)                        # the decoder's x16 image
    with torch.no_grad():
        up_lr = F.interpolate(x, scale_factor=16, mode='bilinear', align_corners=False)                    # :344
        r12 = getattr(vgg(ref_img.expand(-1, 3, -1, -1)), 'relu1_2')                                       # :345
        a = F.normalize(M.extract_image_patches(r12, ksizes=[16, 16], strides=[16, 16], rates=[1, 1], padding='same'), dim=1)
        lr12 = getattr(vgg(up_lr.expand(-1, 3, -1, -1)), 'relu1_2')                                        # :349
        b = F.normalize(M.extract_image_patches(lr12, ksizes=[16, 16], strides=[16, 16], rates=[1, 1], padding='same'), dim=1)
        mask = torch.sum(a.contiguous() * b.contiguous(), dim=1, keepdim=True).view(3, 1, 4, 6)           # :352-353
    _save('vgg_mask_small', x=_np(x), ref_img=_np(ref_img), relu1_2=_np(r12[:1, :, :16, :16]), mask=_np(mask), seed=np.array([81]))


def gen_full(ref):
    """The WHOLE reference model: ``GPEMSR(...)`` built from option/output_GPEMSR_x{8,16}.yml by the reference's own
    model/GPEMSR.py (through the BasicSR shim), every live parameter overwritten by ``fill_state`` (names / shapes asserted
    equal to ``gpemsr_b200.GPEMSR``'s own), run on a seeded 5-frame 16 x 16 LR window."""
    import yaml
    import gpemsr_b200
    from gpemsr_b200.gpemsr import DEAD_PREFIXES
    kw = lambda net: dict(argref=net['argref'], nf=net['nf'], nframes=net['nframes'], groups=net['groups'], front_RBs=net['front_RBs'],
                          back_RBs=net['back_RBs'], w_ref=net['w_ref'], ref_fusion_feat_RBs=net['ref_fusion_feat_RBs'],
                          align_mode=net['align_mode'], fusion_mode=net['fusion_mode'], mode=net['mode'])
    nets, shapes = {}, {}
    for scale in (8, 16):                                  # the mirror's own names / shapes (before torch.load is neutralised)
        with open(os.path.join(ref, 'option', f'output_GPEMSR_x{scale}.yml')) as f:
            nets[scale] = yaml.safe_load(f)['network']
        mine = gpemsr_b200.GPEMSR(None, None, scale=scale, **kw(nets[scale]))
        shapes[scale] = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    gp = basicsr_shim.install(ref)
    for scale in (8, 16):
        model = gp.GPEMSR(ref_path_G=None, ref_path_Indexer=None, scale=scale, **kw(nets[scale])).eval()
        live = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.startswith(DEAD_PREFIXES)}
        assert live == shapes[scale], (set(live) ^ set(shapes[scale]))
        sd = W.fill_state(shapes[scale], seed=900 + scale)
        res = model.load_state_dict(sd, strict=False)
        assert not res.unexpected_keys and all(k.startswith(DEAD_PREFIXES) for k in res.missing_keys)
        x = torch.rand(1, 5, 1, 16, 16, generator=torch.Generator().manual_seed(910 + scale))
        with torch.no_grad():
            out, ref_img = model(x)
        _save(f'full_x{scale}', x=_np(x), out=_np(out), ref_img_sub=_np(ref_img[0, :, 0, ::4, ::4]), seed=np.array([900 + scale]),
              scale=np.array([scale]))
    gp._oracle_restore()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference/GPEMSR-CREMI/GPEMSR')
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    sys.path.insert(0, a.ref)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gens = dict(codebook=gen_codebook, decoder=gen_decoder, blocks=gen_blocks, tail=gen_tail,
                flow_warp=gen_flow_warp, indexer=gen_indexer, vgg_mask=gen_vgg_mask, full=gen_full)
    for n, fn in gens.items():
        if a.only and n not in a.only.split(','):
            continue
        fn(a.ref)


if __name__ == '__main__':
    main()
