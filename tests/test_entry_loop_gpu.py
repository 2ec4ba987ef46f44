"""GPU: the reference's whole inference loop (output_GPEMSR.py:18-128, restated in oracle/output_loop.py: PNG stack -> 5-frame
windows with padded ends -> model -> tensor2img -> PNG) on the native model, against the same loop on the reference's own device
path (PyTorch eager, oracle/gpu_eager.py) -- uint8 images equal except |diff| <= 1 on < 0.5 % of the pixels (an HR error of
e moves a pixel across a rounding boundary with probability 2 * 255 * e: ~0.1 % at the measured e ~ 2e-6 mean) -- and against
``super_resolve_volume`` (the per-frame-cache driver).  The real, unmodified script is run on the mirror by tests/test_entry_point.py
in the authoring container."""
import os

import numpy as np
import pytest
import torch

from oracle import gpu_eager as GE
from oracle import output_loop as OL
from full_model_util import build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('scale', [8, 16])
def test_png_stack_through_the_slice_loop(cuda_dev, tmp_path, scale):
    import cv2
    from gpemsr_b200.volume import super_resolve_volume
    S, lr = 9, 16
    rng = np.random.default_rng(60 + scale)
    # smooth-ish synthetic EM slices (uint8), correlated along z like a real stack
    base = rng.random((lr + 8, lr + 8))
    vol = np.stack([np.clip(255 * (0.6 * base[i % 4:i % 4 + lr, 2:2 + lr] + 0.4 * rng.random((lr, lr))), 0, 255).astype(np.uint8)
                    for i in range(S)])
    d_lr, d_hr = tmp_path / 'LR', tmp_path / 'HR'
    d_lr.mkdir(); d_hr.mkdir()
    for i in range(S):
        cv2.imwrite(str(d_lr / f'{i}.png'), vol[i])
        cv2.imwrite(str(d_hr / f'{i}.png'), np.zeros((scale * lr, scale * lr), np.uint8))
    model, sd = build(scale, seed=400 + scale, device=cuda_dev)
    got = OL.run(model, str(d_hr), str(d_lr), str(tmp_path / 'SR_native'), device='cuda')
    model.check()
    sd_dev = GE.to_device(sd)
    want = OL.run(lambda x: GE.forward(x, sd_dev, scale), str(d_hr), str(d_lr), str(tmp_path / 'SR_eager'), device='cuda')
    assert len(got) == len(want) == S
    assert sorted(os.listdir(tmp_path / 'SR_native'), key=lambda n: int(n[:-4])) == [f'{i}.png' for i in range(S)]
    for k in range(S):
        a = cv2.imread(str(tmp_path / 'SR_native' / f'{k}.png'), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(str(tmp_path / 'SR_eager' / f'{k}.png'), cv2.IMREAD_UNCHANGED)
        assert a.shape == (scale * lr, scale * lr) and a.dtype == np.uint8
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 5e-3, (k, int(d.max()), float((d > 0).mean()))
    # the volume driver (every slice encoded once) writes the same images
    volf = torch.from_numpy(vol.astype(np.float32) / 255.0).view(S, 1, lr, lr).cuda()
    hr = super_resolve_volume(model, volf)
    for k in range(S):
        d = np.abs(OL.tensor2img(hr[k]).astype(np.int32) - got[k].astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 5e-3, k
