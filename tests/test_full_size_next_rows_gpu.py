"""GPU parity at BASELINE configs[1] sizes (x16, 80 x 80 LR -> 1280 x 1280) for the SURVEY.md 8(f) rows, one frame each so
that the CPU oracle finishes in seconds: reference-width Indexer16, VGG relu1_2 mask at 1280^2, SpyNet at 320^2, DCNv2Pack."""
import pytest
import torch

from oracle import ref_ops as R
from oracle import weights as W

pytestmark = pytest.mark.gpu


def test_indexer16_reference_width_one_frame(cuda_dev):
    from gpemsr_b200.indexer import Indexer16
    cfg = dict(channel_list=[64, 64, 128, 256, 512], im_channel=1, num_resblock_per_scale=2, num_output_resblck=3, latent_dim=512,
               use_non_local=True)                                      # option/output_GPEMSR_x16.yml:30-36
    sd = W.fill(W.indexer_spec(16), seed=201)
    m = Indexer16(cfg).cuda()
    m.load_state_dict(sd, strict=True)
    x = torch.rand(1, 1, 80, 80, generator=torch.Generator().manual_seed(202))
    torch.set_num_threads(max(1, torch.get_num_threads()))
    want = R.indexer_features(x, sd)
    got = m.features(x.cuda())
    m.check()
    assert (got.cpu() - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())


def test_vgg_mask_full_resolution_one_frame(cuda_dev):
    from gpemsr_b200.vgg import VGG19Slice1
    sd = W.fill(W.vgg_slice1_spec(), seed=203)
    m = VGG19Slice1().cuda()
    m.load_reference_state_dict(sd)
    g = torch.Generator().manual_seed(204)
    ref_img, lr = torch.rand(1, 1, 1280, 1280, generator=g), torch.rand(1, 1, 80, 80, generator=g)
    want = R.similarity_mask(ref_img, lr, sd, 16)
    got = m.similarity_mask(ref_img.cuda(), lr.cuda(), 16)
    m.check()
    assert tuple(got.shape) == (1, 1, 80, 80)
    assert (got.cpu() - want).abs().max().item() <= 1e-5


def test_spynet_320_two_pairs(cuda_dev):
    from gpemsr_b200.spynet import SpyNet
    from oracle.basicsr_shim import SpyNet as RefSpyNet
    ref = RefSpyNet().eval()
    sd = {**W.fill(W.spynet_spec(), seed=205, gain=2.0), 'mean': ref.mean, 'std': ref.std}
    ref.load_state_dict(sd, strict=True)
    m = SpyNet().cuda()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(206)
    a = torch.rand(2, 1, 320, 320, generator=g)
    b = (a + 0.05 * torch.rand(2, 1, 320, 320, generator=g)).clamp(0, 1)
    with torch.no_grad():
        want = ref(a.expand(-1, 3, -1, -1), b.expand(-1, 3, -1, -1))
    got = m(a.cuda(), b.cuda())
    m.check()
    assert (got.cpu() - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


def test_dcnv2pack_lr_resolution(cuda_dev):
    from gpemsr_b200.dcn import DCNv2Pack
    from oracle.basicsr_shim import DCNv2Pack as Ref
    ref = Ref(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8).eval()
    g = torch.Generator().manual_seed(207)
    ref.conv_offset.weight.data.copy_(torch.randn(ref.conv_offset.weight.shape, generator=g) * 0.03)
    ref.conv_offset.bias.data.copy_(torch.randn(216, generator=g) * 0.5)
    m = DCNv2Pack(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8).cuda()
    m.load_state_dict(ref.state_dict(), strict=True)
    x, feat = torch.randn(5, 64, 80, 80, generator=g), torch.randn(5, 64, 80, 80, generator=g)
    with torch.no_grad():
        want = ref(x, feat)
    got = m(x.cuda(), feat.cuda())
    m.check()
    assert (got.cpu() - want).abs().max().item() <= 5e-5 * max(1.0, want.abs().max().item())


def test_x8_config1_chain_vs_oracle(cuda_dev):
    """BASELINE configs[0] (x8, 32 x 32 LR -> 256 x 256, option/output_GPEMSR_x8.yml) through the x8 variants of every stage:
    lrGenerator8.ref_extract (Indexer8 with its DownBlock -> fused lookup -> decoder on the 16 x 16 latents) and the VGG
    mask at scale 8 (mask on the H/2 grid, model/GPEMSR.py:395-403), against the CPU oracle on the same weights."""
    from gpemsr_b200.indexer import lrGenerator8
    from gpemsr_b200.vgg import VGG19Slice1
    icfg = dict(channel_list=[64, 64, 128, 256, 512], im_channel=1, num_resblock_per_scale=2, num_output_resblck=3, latent_dim=512,
                use_non_local=True)
    dcfg = dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=3, latent_dim=512,
                use_non_local=True)
    args = dict(Indexer8=icfg, Decoder=dcfg, Codebook=dict(num_codebook_vectors=1024, latent_dim=512, beta=1))
    sd_i = W.fill(W.indexer_spec(8), seed=211)
    sd_d = W.fill(W.decoder_spec(), seed=212)
    emb = W.fill(W.codebook_spec(), seed=213, gain=300.0)['embedding.weight']
    gen = lrGenerator8(args).cuda()
    gen.indexer.load_state_dict(sd_i, strict=True)
    gen.decoder.load_state_dict(sd_d, strict=True)
    gen.codebook.embedding.weight.data.copy_(emb)
    x = torch.rand(3, 1, 32, 32, generator=torch.Generator().manual_seed(214))          # BASELINE: 3-slice 32 x 32 stack
    feat = gen.indexer.features(x.cuda())
    gen.indexer.check()
    want_feat = R.indexer_features(x, sd_i)
    assert tuple(feat.shape) == (3, 512, 16, 16)
    assert (feat.cpu() - want_feat).abs().max().item() <= 1e-4 * max(1.0, want_feat.abs().max().item())
    # downstream of the (discrete) lookup: decode the ORACLE's indices' neighbourhood by feeding the GPU's own features to the
    # oracle tail, so a near-tied logit cannot make the comparison flaky
    head = {'embedding.weight': sd_i['embedding.weight'], 'embedding.bias': sd_i['embedding.bias']}
    want, _ = R.ref_extract_from_feat(feat.cpu(), head, emb, sd_d)
    got = gen.ref_extract(x.cuda())
    gen.decoder.check()
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
    assert tuple(got[-1].shape) == (3, 1, 256, 256)
    for a, b in zip(got, want):
        assert (a.cpu() - b).abs().max().item() <= 1e-3 * max(1.0, b.abs().max().item())
    sd_v = W.fill(W.vgg_slice1_spec(), seed=215)
    vgg = VGG19Slice1().cuda()
    vgg.load_reference_state_dict(sd_v)
    mask = vgg.similarity_mask(got[-1], x.cuda(), 8)
    vgg.check()
    assert tuple(mask.shape) == (3, 1, 16, 16)                       # view(B*N, 1, H//2, W//2), model/GPEMSR.py:403
    assert (mask.cpu() - R.similarity_mask(got[-1].cpu(), x, sd_v, 8)).abs().max().item() <= 1e-5
