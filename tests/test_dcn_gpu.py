"""GPU parity: DCNv2Pack (BasicSR arch_util.DCNv2Pack -> torchvision.ops.deform_conv2d) through the C ABI vs
the CPU oracle (oracle/basicsr_shim.py, which calls torchvision's own deform_conv2d)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(seed, off_scale):
    from gpemsr_b200.dcn import DCNv2Pack
    from oracle.basicsr_shim import DCNv2Pack as Ref
    ref = Ref(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8).eval()
    g = torch.Generator().manual_seed(seed)
    ref.bias.data.copy_(torch.randn(64, generator=g) * 0.1)
    ref.conv_offset.weight.data.copy_(torch.randn(ref.conv_offset.weight.shape, generator=g) * off_scale)
    ref.conv_offset.bias.data.copy_(torch.randn(216, generator=g) * 0.5)
    m = DCNv2Pack(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8).cuda()
    m.load_state_dict(ref.state_dict(), strict=True)
    return ref, m


@pytest.mark.parametrize('shape,off_scale', [((2, 64, 12, 10), 0.02), ((1, 64, 33, 41), 0.08), ((1, 64, 8, 8), 0.0)])
def test_dcnv2pack_vs_torchvision(cuda_dev, shape, off_scale):
    ref, m = _pair(51, off_scale)
    g = torch.Generator().manual_seed(52)
    x, feat = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    with torch.no_grad():
        want = ref(x, feat)
        om = ref.conv_offset(feat)
    got = m(x.cuda(), feat.cuda())
    m.check()
    assert tuple(got.shape) == tuple(want.shape)
    # offsets of several pixels (incl. samples outside the image) when off_scale > 0; plain 3x3 conv with mask 0.5+ when 0
    if off_scale:
        assert om[:, :144].abs().max().item() > 1.5
    assert (got.cpu() - want).abs().max().item() <= 5e-5 * max(1.0, want.abs().max().item())


def test_dcnv2pack_refuses_other_configs(cuda_dev):
    from gpemsr_b200.dcn import DCNv2Pack
    with pytest.raises(Exception):
        DCNv2Pack(64, 64, 3, stride=2, padding=1, deformable_groups=8)
    with pytest.raises(Exception):
        DCNv2Pack(32, 64, 3, stride=1, padding=1, deformable_groups=8)
