"""GPU parity: Decoder / blocks / SR tail (a-3, a-4) through the C ABI vs golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from oracle import weights as W

pytestmark = pytest.mark.gpu
T = torch.from_numpy
TOL = 1e-3          # north_star: HR images within 1e-3 max-abs of the reference


def _cuda_sd(sd):
    return {k: v.cuda() for k, v in sd.items()}


def _err(got, want):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else got
    want = want.detach().cpu().numpy() if torch.is_tensor(want) else want
    return float(np.abs(got - want).max()), float(np.abs(want).max())


def _mini_decoder():
    import gpemsr_b200
    cfg = dict(channel_list=[64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=0, latent_dim=64,
               use_non_local=False)
    return gpemsr_b200.Decoder(cfg).cuda()


def test_blocks_golden(golden, cuda_dev):
    """ResidualBlock (with channel_up), UpBlock and NonLocalBlock one by one (fixtures from model/blocks.py)."""
    from gpemsr_b200 import decoder as D, igemm as G
    g = golden('blocks_small')
    s41, s43, s44 = (int(s) for s in g['seeds'])
    host = _mini_decoder()
    x = T(g['x']).cuda()
    n, c, h, w = x.shape
    P = D._Plan(host, n, h, w, x.device)
    xin = P.act('x', G.Geom(n, h, w, True), c, f32=True)
    G.pack_nchw(x, xin)
    # ResidualBlock(32 -> 64)
    spec = W.OrderedDict(); W._resblock(spec, 'rb', 32, 64)
    rb = D.ResidualBlock(32, 64).cuda()
    rb.load_state_dict({k[3:]: v for k, v in W.fill(spec, s41).items()}, strict=True)
    y = host._res_block(P, 'rb', rb, xin)
    e, m = _err(G.unpack_nchw(y), g['rb'])
    assert e <= 2e-5 * max(1.0, m), (e, m)
    # UpBlock(32 -> 64)
    ub = D.UpBlock(32, 64).cuda()
    ub.load_state_dict(W.fill(W.OrderedDict([('upblock.weight', ('convT', (32, 64, 3, 3))), ('upblock.bias', ('bias', (64,)))]), s43))
    y = host._up_block(P, 'ub', ub, xin, need_f32=True)
    e, m = _err(G.unpack_nchw(y), g['up'])
    assert e <= 2e-5 * max(1.0, m), (e, m)
    # NonLocalBlock(64)
    xn = T(g['xn']).cuda()
    n, c, h, w = xn.shape
    P2 = D._Plan(host, n, h, w, xn.device)
    xa = P2.act('x', G.Geom(n, h, w, True), c, f32=True)
    G.pack_nchw(xn, xa)
    spec_n = W.OrderedDict(); W._nonlocal(spec_n, 'nl', 64)
    nl = D.NonLocalBlock(64).cuda()
    nl.load_state_dict({k[3:]: v for k, v in W.fill(spec_n, s44).items()}, strict=True)
    y = host._non_local(P2, 'nl', nl, xa)
    e, m = _err(G.unpack_nchw(y), g['nl'])
    G.check_pipeline(P.err); G.check_pipeline(P2.err)
    assert e <= 2e-5 * max(1.0, m), (e, m)


def test_decoder_small_golden(golden, cuda_dev):
    import gpemsr_b200
    g = golden('decoder_small')
    cfg = dict(channel_list=[64, 64, 32, 32, 32], im_channel=1, num_resblock_per_scale=1, num_input_resblck=2,
               latent_dim=64, use_non_local=True)
    dec = gpemsr_b200.Decoder(cfg).cuda()
    dec.load_state_dict(W.fill(W.decoder_spec(cfg['channel_list'], 64, 2, 1, True, 1), seed=int(g['seed'][0])), strict=True)
    feats = dec.multi_scale_feat_calculate(T(g['x']).cuda())
    dec.check()
    assert len(feats) == 5
    for i, f in enumerate(feats):
        assert tuple(f.shape) == g[f'feat{i}'].shape
        e, m = _err(f, g[f'feat{i}'])
        assert e <= 1e-4 * max(1.0, m), (i, e, m)
    img = dec(T(g['x']).cuda())
    assert _err(img, g['feat4'])[0] <= 1e-4
    # the un-composed final stage (separate UpBlock phases + output conv) must agree too
    dec2 = gpemsr_b200.Decoder(cfg, compose_final=False).cuda()
    dec2.load_state_dict(dec.state_dict(), strict=True)
    assert _err(dec2(T(g['x']).cuda()), g['feat4'])[0] <= 1e-4
    dec2.check()


def test_decoder_reference_width_golden(golden, cuda_dev):
    import gpemsr_b200
    g = golden('decoder_full_4x4')
    cfg = dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=3,
               latent_dim=512, use_non_local=True)
    dec = gpemsr_b200.Decoder(cfg).cuda()
    dec.load_state_dict(W.fill(W.decoder_spec(), seed=int(g['seed'][0])), strict=True)
    feats = dec.multi_scale_feat_calculate(T(g['x']).cuda())
    dec.check()
    for i, f in enumerate(feats):
        e, m = _err(f, g[f'feat{i}'])
        assert e <= 1e-4 * max(1.0, m), (i, e, m)
    assert _err(feats[-1], g['feat4'])[0] <= TOL


@pytest.mark.parametrize('scale', [8, 16])
def test_tail_golden(golden, scale, cuda_dev):
    import gpemsr_b200
    g = golden(f'tail_x{scale}')
    tail = gpemsr_b200.SRTail(64, 10, scale).cuda()
    tail.load_state_dict(W.fill(W.tail_spec(64, 10, scale), seed=int(g['seed'][0]), gain=3.0 ** 0.5), strict=True)
    out = tail(T(g['fea']).cuda(), T(g['x_center']).cuda())
    tail.check()
    assert out.shape == g['out'].shape
    e, m = _err(out, g['out'])
    assert e <= TOL, (e, m)
    assert e <= 1e-4 * max(1.0, m), (e, m)


def test_decoder_x8_config1_vs_oracle(cuda_dev):
    """BASELINE config 1 shape: 5 frames, 16x16 latents (32x32 LR, x8) -> 5 x 1 x 256 x 256."""
    import gpemsr_b200
    cfg = dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=3,
               latent_dim=512, use_non_local=True)
    sd = W.fill(W.decoder_spec(), seed=101)
    emb = W.fill(W.codebook_spec(), seed=102)['embedding.weight']
    idx = torch.randint(0, 1024, (5 * 16 * 16,), generator=torch.Generator().manual_seed(103))
    zq = emb[idx].view(5, 16, 16, 512).permute(0, 3, 1, 2).contiguous()
    want = R.decoder_multi_scale(zq, sd)
    dec = gpemsr_b200.Decoder(cfg).cuda()
    dec.load_state_dict(sd, strict=True)
    got = dec.multi_scale_feat_calculate(zq.cuda())
    dec.check()
    for i, (a, b) in enumerate(zip(got, want)):
        e, m = _err(a, b)
        assert e <= 1e-4 * max(1.0, m), (i, e, m)
    e, m = _err(got[-1], want[-1])
    assert e <= TOL
    # bf16 single-pass mode: same graph, looser numerics (reported, not the parity path)
    dec16 = gpemsr_b200.Decoder(cfg, precision='bf16').cuda()
    dec16.load_state_dict(sd, strict=True)
    got16 = dec16(zq.cuda())
    dec16.check()
    e16, m16 = _err(got16, want[-1])
    assert e16 <= 0.1 * max(1.0, m16), (e16, m16)


def test_tail_x8_config1_vs_oracle(cuda_dev):
    import gpemsr_b200
    sd = W.fill(W.tail_spec(64, 10, 8), seed=111, gain=3.0 ** 0.5)
    fea = torch.randn(1, 64, 32, 32, generator=torch.Generator().manual_seed(112))
    xc = torch.rand(1, 1, 32, 32, generator=torch.Generator().manual_seed(113))
    want = R.sr_tail(fea, xc, sd, 8)
    tail = gpemsr_b200.SRTail(64, 10, 8).cuda()
    tail.load_state_dict(sd, strict=True)
    got = tail(fea.cuda(), xc.cuda())
    tail.check()
    e, m = _err(got, want)
    assert e <= TOL and e <= 1e-4 * max(1.0, m), (e, m)


def test_cremi_shape_x16_one_frame_vs_oracle(cuda_dev):
    """BASELINE config 2 shape (x16, 80x80 LR): decoder on one 80x80 latent frame and the x16 tail, vs the CPU oracle."""
    import gpemsr_b200
    torch.set_num_threads(max(1, torch.get_num_threads()))
    cfg = dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=3,
               latent_dim=512, use_non_local=True)
    sd = W.fill(W.decoder_spec(), seed=121)
    emb = W.fill(W.codebook_spec(), seed=122)['embedding.weight']
    idx = torch.randint(0, 1024, (80 * 80,), generator=torch.Generator().manual_seed(123))
    zq = emb[idx].view(1, 80, 80, 512).permute(0, 3, 1, 2).contiguous()
    want = R.decoder_forward(zq, sd)
    dec = gpemsr_b200.Decoder(cfg).cuda()
    dec.load_state_dict(sd, strict=True)
    got = dec(zq.cuda())
    dec.check()
    e, m = _err(got, want)
    assert got.shape == (1, 1, 1280, 1280)
    assert e <= TOL and e <= 1e-4 * max(1.0, m), (e, m)
    sdt = W.fill(W.tail_spec(64, 10, 16), seed=124, gain=3.0 ** 0.5)
    fea = torch.randn(1, 64, 80, 80, generator=torch.Generator().manual_seed(125))
    xc = torch.rand(1, 1, 80, 80, generator=torch.Generator().manual_seed(126))
    want_t = R.sr_tail(fea, xc, sdt, 16)
    tail = gpemsr_b200.SRTail(64, 10, 16).cuda()
    tail.load_state_dict(sdt, strict=True)
    got_t = tail(fea.cuda(), xc.cuda())
    tail.check()
    e, m = _err(got_t, want_t)
    assert e <= TOL and e <= 1e-4 * max(1.0, m), (e, m)
