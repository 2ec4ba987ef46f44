"""GPU parity: VGG19 relu1_2 patch-similarity mask (model/GPEMSR.py:344-353, model/VGG.py) through the
C ABI vs the golden vector made with the reference's VGG19 + extract_image_patches, and vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from oracle import weights as W

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _vgg(seed):
    from gpemsr_b200.vgg import VGG19Slice1
    m = VGG19Slice1().cuda()
    sd = W.fill(W.vgg_slice1_spec(), seed=seed)
    m.load_reference_state_dict(sd)
    return m, sd


def test_relu1_2_golden(golden, cuda_dev):
    g = golden('vgg_mask_small')
    m, sd = _vgg(int(g['seed'][0]))
    got = m.relu1_2(T(g['ref_img']).cuda())
    want = R.vgg_relu1_2(T(g['ref_img']).expand(-1, 3, -1, -1), sd)
    assert np.abs(got[:1, :, :16, :16].cpu().numpy() - g['relu1_2']).max() <= 2e-5 * max(1.0, float(np.abs(g['relu1_2']).max()))
    assert (got.cpu() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_similarity_mask_golden(golden, cuda_dev):
    g = golden('vgg_mask_small')
    m, _ = _vgg(int(g['seed'][0]))
    mask = m.similarity_mask(T(g['ref_img']).cuda(), T(g['x']).cuda(), 16)
    m.check()
    assert tuple(mask.shape) == g['mask'].shape
    assert np.abs(mask.cpu().numpy() - g['mask']).max() <= 1e-5           # a cosine: values in [0, 1]


def test_patch_similarity_vs_oracle_torchvision_keys(cuda_dev):
    """Larger, signed inputs (negative pre-activations exercise the ReLUs), weights given under torchvision's key names,
    a patch grid that straddles warp boundaries, and identical inputs -> similarity exactly ~1."""
    from gpemsr_b200.vgg import VGG19Slice1
    sd = W.fill(W.vgg_slice1_spec(), seed=91)
    m = VGG19Slice1().cuda()
    m.load_reference_state_dict({'features.0.weight': sd['slice1.0.weight'], 'features.0.bias': sd['slice1.0.bias'],
                                 'features.2.weight': sd['slice1.2.weight'], 'features.2.bias': sd['slice1.2.bias'],
                                 'features.5.weight': torch.zeros(1)})
    g = torch.Generator().manual_seed(92)
    a = torch.randn(2, 1, 48, 80, generator=g)
    b = a + 0.5 * torch.randn(2, 1, 48, 80, generator=g)
    got = m.patch_similarity(a.cuda(), b.cuda())
    m.check()
    want = R.patch_similarity(a, b, sd)
    assert (got.cpu() - want).abs().max().item() <= 1e-5
    same = m.patch_similarity(a.cuda(), a.cuda())
    assert (same.cpu() - 1.0).abs().max().item() <= 1e-5
    with pytest.raises(Exception):
        m.patch_similarity(a[:, :, :40].cuda(), b[:, :, :40].cuda())       # 40 is not a multiple of 16: refused, not approximated
    with pytest.raises(Exception):
        m.patch_similarity(a, b)                                           # CPU tensors: no fallback
