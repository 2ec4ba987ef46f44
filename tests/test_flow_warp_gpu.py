"""GPU parity: gpemsr_flow_warp (through the C ABI) vs the CPU oracle, fixtures and KATs."""
import numpy as np
import pytest
import torch

from oracle.flow_warp import flow_warp_numpy, flow_warp_torch

pytestmark = pytest.mark.gpu
TOL = 1e-5      # north_star: warped features within 1e-5 max-abs of the reference


def _run(x, flow, pm, ac=True, form='cpu'):
    """form='cpu': the rounding of ATen's CPU kernels (what the CPU oracle and the fixtures hold); 'cuda': ATen's CUDA form."""
    import gpemsr_b200
    out = gpemsr_b200.flow_warp(torch.from_numpy(x).cuda(), torch.from_numpy(flow).cuda(), 'bilinear', pm, ac, coord_form=form)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_golden_fixture(golden, cuda_dev):
    g = golden('flow_warp_small')
    for pm in ('border', 'zeros'):
        got = _run(g['x'], g['flow'], pm)
        assert np.abs(got - g['out_' + pm]).max() <= TOL
    got = _run(g['x2'], g['flow2'], 'border')
    assert np.abs(got - g['out2_border']).max() <= TOL


@pytest.mark.parametrize('shape', [(1, 3, 4, 4), (1, 3, 20, 20), (2, 5, 33, 47), (1, 64, 156, 156), (1, 3, 128, 128),
                                   (1, 1, 1, 1), (1, 2, 1, 9), (1, 2, 7, 1)])
@pytest.mark.parametrize('pm', ['border', 'zeros'])
def test_vs_oracle_seeded(shape, pm, cuda_dev):
    n, c, h, w = shape
    rng = np.random.default_rng(1000 + h * w + c)
    x = rng.standard_normal(shape).astype(np.float32)
    flow = (4.0 * rng.standard_normal((n, h, w, 2))).astype(np.float32)
    want = flow_warp_numpy(x, flow, 'bilinear', pm)
    got = _run(x, flow, pm)
    err = np.abs(got - want).max()
    assert err <= TOL, err
    # and against ATen on CPU (the reference's code on the CPU device)
    ref = flow_warp_torch(torch.from_numpy(x), torch.from_numpy(flow), 'bilinear', pm).numpy()
    assert np.abs(got - ref).max() <= TOL
    # the library default reproduces ATen's CUDA rounding of the normalisation (reciprocal multiply)
    want_c = flow_warp_numpy(x, flow, 'bilinear', pm, recip=True)
    assert np.abs(_run(x, flow, pm, form='cuda') - want_c).max() <= TOL


def test_align_corners_false(cuda_dev):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((1, 4, 19, 23)).astype(np.float32)
    flow = (2.0 * rng.standard_normal((1, 19, 23, 2))).astype(np.float32)
    for pm in ('border', 'zeros'):
        want = flow_warp_numpy(x, flow, 'bilinear', pm, align_corners=False)
        assert np.abs(_run(x, flow, pm, ac=False) - want).max() <= TOL


def test_kats(cuda_dev):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 2, 6, 9)).astype(np.float32)
    zero = np.zeros((1, 6, 9, 2), np.float32)
    for pm in ('border', 'zeros'):
        assert np.allclose(_run(x, zero, pm), x, atol=2e-6)
    sh = zero.copy(); sh[..., 0] = 2.0; sh[..., 1] = -1.0
    exp_b = x[:, :, np.clip(np.arange(6) - 1, 0, 5)][:, :, :, np.clip(np.arange(9) + 2, 0, 8)]
    assert np.allclose(_run(x, sh, 'border'), exp_b, atol=2e-6)
    exp_z = exp_b.copy(); exp_z[:, :, 0, :] = 0; exp_z[:, :, :, 7:] = 0
    assert np.allclose(_run(x, sh, 'zeros'), exp_z, atol=2e-6)
    # far out-of-range flow: zeros -> 0 everywhere, border -> corner values; huge magnitudes must not fault
    far = np.full((1, 6, 9, 2), 1e30, np.float32)
    assert np.array_equal(_run(x, far, 'zeros'), np.zeros_like(x))
    assert np.allclose(_run(x, far, 'border'), np.broadcast_to(x[:, :, -1:, -1:], x.shape), atol=0)


def test_empty_and_errors(cuda_dev):
    import gpemsr_b200
    e = gpemsr_b200.flow_warp(torch.empty(0, 3, 4, 4).cuda(), torch.empty(0, 4, 4, 2).cuda())
    assert e.shape == (0, 3, 4, 4)
    with pytest.raises(AssertionError):
        gpemsr_b200.flow_warp(torch.zeros(1, 1, 4, 4).cuda(), torch.zeros(1, 5, 4, 2).cuda())
    with pytest.raises(NotImplementedError):
        gpemsr_b200.flow_warp(torch.zeros(1, 1, 4, 4).cuda(), torch.zeros(1, 4, 4, 2).cuda(), 'nearest')
    with pytest.raises(gpemsr_b200.GpemsrError):
        gpemsr_b200.flow_warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4, 2))      # CPU tensors: no fallback


def test_full_size_properties(cuda_dev):
    """BASELINE config 4 at full size (64 x 1250^2): size-independent properties."""
    import gpemsr_b200
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(1, 64, 1250, 1250, device='cuda', generator=g)
    flow = 2.0 * torch.randn(1, 1250, 1250, 2, device='cuda', generator=g)
    out = gpemsr_b200.flow_warp(x, flow, 'bilinear', 'border')
    # linearity in x
    y = torch.randn(1, 64, 1250, 1250, device='cuda', generator=g)
    out2 = gpemsr_b200.flow_warp(x + 2.0 * y, flow, 'bilinear', 'border')
    outy = gpemsr_b200.flow_warp(y, flow, 'bilinear', 'border')
    assert (out2 - (out + 2.0 * outy)).abs().max().item() < 1e-4
    # convex combination of inputs: output bounded by input range
    assert out.max() <= x.max() and out.min() >= x.min()
    # integer flow = shift, up to the reference's own normalise/unnormalise coordinate error
    # (~1e-4 px at w=1250) times the local gradient of x
    fi = torch.zeros_like(flow); fi[..., 0] = 3.0; fi[..., 1] = -2.0
    sh = gpemsr_b200.flow_warp(x, fi, 'bilinear', 'zeros')
    assert torch.allclose(sh[:, :, 2:, :-3], x[:, :, :-2, 3:], atol=5e-3)
    # a CPU-oracle spot check on a crop-free subsample of channels
    sub = x[:, :2].contiguous()
    want = flow_warp_numpy(sub.cpu().numpy(), flow.cpu().numpy(), 'bilinear', 'border', recip=True)
    got = gpemsr_b200.flow_warp(sub, flow, 'bilinear', 'border').cpu().numpy()
    assert np.abs(got - want).max() <= TOL
    want = flow_warp_numpy(sub.cpu().numpy(), flow.cpu().numpy(), 'bilinear', 'border')
    got = gpemsr_b200.flow_warp(sub, flow, 'bilinear', 'border', coord_form='cpu').cpu().numpy()
    assert np.abs(got - want).max() <= TOL
