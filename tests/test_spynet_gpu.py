"""GPU parity: SpyNet (BasicSR spynet_arch.py restated in oracle/basicsr_shim.py -- parity unpinned at the
BasicSR boundary) and its helper kernels through the C ABI vs torch / the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def test_resize_bilinear_matches_aten(cuda_dev):
    from gpemsr_b200.spynet import resize_bilinear, avg_pool2
    x = _rand((2, 3, 9, 7), 1)
    for ho, wo in ((18, 14), (32, 32), (5, 4), (9, 7)):
        got = resize_bilinear(x.cuda(), ho, wo, False).cpu()
        want = F.interpolate(x, size=(ho, wo), mode='bilinear', align_corners=False)
        assert (got - want).abs().max().item() <= 2e-6, (ho, wo)
    got = resize_bilinear(x.cuda(), 36, 28, False, scale=4).cpu()
    assert (got - F.interpolate(x, scale_factor=4, mode='bilinear', align_corners=False)).abs().max().item() <= 2e-6
    # SpyNet's flow upsampling: x2, align_corners=True, values doubled, replicate tail to an odd target size, NHWC copy
    f = _rand((2, 2, 5, 6), 2)
    up = F.interpolate(f, scale_factor=2, mode='bilinear', align_corners=True) * 2.0
    up = F.pad(F.pad(up, [0, 0, 0, 1], mode='replicate'), [0, 1, 0, 0], mode='replicate')
    two = torch.full((2,), 2.0).cuda()
    out = torch.empty(2, 2, 11, 13).cuda()
    nhwc = torch.empty(2, 11, 13, 2).cuda()
    resize_bilinear(f.cuda(), 11, 13, True, rep=(10, 12), mul=two, out=out, out_nhwc=nhwc)
    assert (out.cpu() - up).abs().max().item() <= 2e-6
    assert torch.equal(nhwc.permute(0, 3, 1, 2).contiguous(), out)
    acc = torch.ones(2, 2, 11, 13).cuda()
    resize_bilinear(f.cuda(), 11, 13, True, rep=(10, 12), mul=two, out=acc, accumulate=True)
    assert (acc.cpu() - (up + 1.0)).abs().max().item() <= 2e-6
    # normalisation with channel broadcast
    g1 = _rand((2, 1, 8, 8), 3)
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    got = resize_bilinear(g1.cuda(), 8, 8, False, c_out=3, sub=mean.cuda(), div=std.cuda()).cpu()
    assert (got - (g1 - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)).abs().max().item() <= 2e-6
    p = avg_pool2(x[:, :, :8, :6].contiguous().cuda()).cpu()
    assert (p - F.avg_pool2d(x[:, :, :8, :6], 2, 2, count_include_pad=False)).abs().max().item() <= 1e-6


def test_conv7x7_ring3_vs_torch(cuda_dev):
    """A 49-tap implicit GEMM on the 3-pixel-ring geometry against F.conv2d (8 -> 32 and 64 -> 32, like BasicModule)."""
    from gpemsr_b200 import igemm as G
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    for ci, co, h, w in ((8, 32, 12, 10), (64, 32, 16, 21)):
        x = _rand((2, ci, h, w), 10 + ci)
        wgt, b = _rand((co, ci, 7, 7), 11 + ci, 0.05), _rand((co,), 12 + ci)
        g = G.Geom(2, h, w, padded=3)
        xa = G.Act(g, ci, 'cuda', f32=False)
        G.pack_nchw(x.cuda(), xa)
        out = torch.empty(2, co, h, w, device='cuda')
        G.igemm(xa, G.Weights(wgt.cuda(), 'conv'), err, bias=b.cuda(), act=G.ACT_RELU, out_nchw=out, nchw_c=co)
        G.check_pipeline(err)
        want = F.relu(F.conv2d(x, wgt, b, 1, 3))
        assert (out.cpu() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item()), ci


def _pair(seed_net, shape, seed_in, cin):
    from gpemsr_b200.spynet import SpyNet
    from oracle.basicsr_shim import SpyNet as RefSpyNet
    ref_net = RefSpyNet().eval()
    g = torch.Generator().manual_seed(seed_net)
    for p in ref_net.parameters():                       # larger-than-default weights so that the flow is not ~0
        p.data.copy_(torch.randn(p.shape, generator=g) * (0.04 if p.dim() == 4 else 0.1))
    net = SpyNet().cuda()
    net.load_state_dict(ref_net.state_dict(), strict=True)
    a = torch.rand((shape[0], cin) + shape[1:], generator=torch.Generator().manual_seed(seed_in))
    b = (a + 0.1 * torch.rand(a.shape, generator=torch.Generator().manual_seed(seed_in + 1))).clamp(0, 1)
    return ref_net, net, a, b


@pytest.mark.parametrize('shape,cin', [((2, 64, 96), 1), ((1, 160, 160), 3)])
def test_spynet_process_vs_oracle(cuda_dev, shape, cin):
    ref_net, net, a, b = _pair(21, shape, 22, cin)
    with torch.no_grad():
        want = ref_net.process(a.expand(-1, 3, -1, -1) if cin == 1 else a, b.expand(-1, 3, -1, -1) if cin == 1 else b)
    got = net.process(a.cuda(), b.cuda())
    net.check()
    assert tuple(got.shape) == tuple(want.shape)
    assert want.abs().max().item() > 0.05                                  # a real flow, not zeros
    assert (got.cpu() - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


def test_spynet_forward_odd_size_vs_oracle(cuda_dev):
    """forward(): sizes that are not multiples of 32 (resize in, resize + rescale out); 5 * 32 = 160 makes the coarsest level
    odd (5 x 5 -> flow 2 x 2 -> replicate-padded upsampling)."""
    ref_net, net, a, b = _pair(31, (2, 150, 139), 32, 1)
    with torch.no_grad():
        want = ref_net(a.expand(-1, 3, -1, -1), b.expand(-1, 3, -1, -1))
    got = net(a.cuda(), b.cuda())
    net.check()
    assert tuple(got.shape) == (2, 2, 150, 139)
    assert (got.cpu() - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
