"""GPU parity of the WHOLE model: ``gpemsr_b200.GPEMSR.forward`` (reference fusion + POD + ThreeDA + SR tail through the C ABI)
vs the golden outputs of the unmodified reference model/GPEMSR.py and, stage by stage, vs the CPU oracle restatement.

Tolerance (BASELINE north_star): HR image <= 1e-3 max-abs; the stage checks use 1e-3 relative to the stage's max-abs."""
import numpy as np
import pytest
import torch

from oracle import gpemsr_model as GM
from full_model_util import build

pytestmark = pytest.mark.gpu
T = torch.from_numpy
STAGES = ['L1_fea0', 'fusion.r0', 'fusion.r1', 'fusion.r2', 'fusion.r3', 'L1_fea', 'L2_fea', 'L3_fea', 'pod.flow', 'pod.o3', 'pod.fea3',
          'pod.o2', 'pod.fea2', 'pod.o1', 'pod.fea1', 'pod.off', 'aligned', 'tda.feat', 'tda.f2', 'tda.attn', 'tda.add', 'fea']


def _record_index_parity(key, rec):
    import json, os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, 'r02_index_parity.json')
        old = json.load(open(path)) if os.path.exists(path) else {}
        old[key] = rec
        json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
    except OSError:
        pass


def _stage_report(model, sd, x, scale):
    """Stage-by-stage comparison: for the all-split-3 model (precision='fp32'); the default precision plan is judged on the
    outputs only (HR image / reference images <= 1e-3), which is BASELINE.json's criterion."""
    model.debug = {}
    out, ref_img = model(x.cuda())
    model.check()
    dbg, model.debug = model.debug, None
    taps = {}
    with torch.no_grad():
        want_out, want_ref = GM.forward(x, sd, scale, taps)
    rows, bad = [], []
    for k in STAGES:
        if k not in taps:
            continue
        w = taps[k]
        w = torch.cat(w, 0) if isinstance(w, list) else w
        got = dbg[k].cpu()
        w = w.reshape(got.shape)
        e, mx = float((got - w).abs().max()), float(w.abs().max())
        rows.append(f'{k:10s} err {e:.3e}  max {mx:.3e}')
        if not e <= 1e-3 * max(1.0, mx):
            bad.append(k)
    e_ref = float((ref_img.cpu() - want_ref).abs().max())
    e_out = float((out.cpu() - want_out).abs().max())
    rows.append(f'ref_img    err {e_ref:.3e}')
    rows.append(f'out        err {e_out:.3e}')
    print('\n'.join(rows))
    return out, ref_img, want_out, want_ref, bad, rows


@pytest.mark.parametrize('scale', [8, 16])
def test_full_model_golden_and_stages(golden, cuda_dev, scale):
    g = golden(f'full_x{scale}')
    model, sd = build(scale, device=cuda_dev, precision='fp32')
    x = T(g['x'])
    out, ref_img, want_out, want_ref, bad, rows = _stage_report(model, sd, x, scale)
    assert not bad, (bad, rows)
    assert tuple(out.shape) == g['out'].shape and tuple(ref_img.shape) == (1, 5, 1, 16 * scale, 16 * scale)
    assert float(np.abs(out.cpu().numpy() - g['out']).max()) <= 1e-3, rows
    assert float(np.abs(ref_img[0, :, 0, ::4, ::4].cpu().numpy() - g['ref_img_sub']).max()) <= 1e-3, rows
    # the second call reuses the cached plan / packed weights / activation buffers (stale contents must not leak); it is
    # not bit-identical because the GroupNorm and patch statistics are accumulated with atomics
    # (float atomics in the patch-similarity sums: ~1e-5 run-to-run).  A different window in between must not leak either.
    model(torch.rand(1, 5, 1, 16, 16, device='cuda'))
    out2, _ = model(x.cuda())
    assert float((out - out2).abs().max()) <= 1e-4
    # the default precision plan (VGG branch and SpyNet in one bf16 pass) against the same golden outputs
    plan, _ = build(scale, device=cuda_dev)
    out_p, ref_p = plan(x.cuda())
    plan.check()
    assert float(np.abs(out_p.cpu().numpy() - g['out']).max()) <= 1e-3
    assert float(np.abs(ref_p[0, :, 0, ::4, ::4].cpu().numpy() - g['ref_img_sub']).max()) <= 1e-3


def test_config1_x8_window_32(cuda_dev):
    """BASELINE configs[0]: x8 on a 5-frame 32 x 32 LR window -> 256 x 256, vs the oracle on the same parameters."""
    model, sd = build(8, seed=77, device=cuda_dev, precision='fp32')
    x = torch.rand(1, 5, 1, 32, 32, generator=torch.Generator().manual_seed(78))
    out, ref_img, want_out, want_ref, bad, rows = _stage_report(model, sd, x, 8)
    assert not bad, (bad, rows)
    assert float((out.cpu() - want_out).abs().max()) <= 1e-3, rows
    mse = float(((out.cpu() - want_out) ** 2).mean())
    assert mse < 1e-8                                  # PSNR delta < 0.01 dB territory: the images differ by ~1e-5


def test_x16_non_square_window(cuda_dev):
    """x16 on a 20 x 24 window (sizes that are not powers of two: SpyNet resizes 80 x 96 -> 96 x 96 internally)."""
    model, sd = build(16, seed=79, device=cuda_dev, precision='fp32')
    x = torch.rand(1, 5, 1, 20, 24, generator=torch.Generator().manual_seed(80))
    out, ref_img, want_out, want_ref, bad, rows = _stage_report(model, sd, x, 16)
    assert not bad, (bad, rows)
    assert float((out.cpu() - want_out).abs().max()) <= 1e-3, rows


def test_rejects_cpu_and_bad_sizes(cuda_dev):
    import gpemsr_b200
    model, _ = build(8, device=cuda_dev)
    with pytest.raises(gpemsr_b200.GpemsrError):
        model(torch.rand(1, 5, 1, 16, 16))
    with pytest.raises(gpemsr_b200.GpemsrError):
        model(torch.rand(1, 5, 1, 18, 16, device='cuda'))


def test_forward_volume_matches_windows_and_oracle(cuda_dev):
    """Per-frame feature cache: forward_volume (every slice encoded once) == forward on the explicit
    windows of output_GPEMSR.py:54-128 (replicate padding at both ends), and == the oracle on the first / a middle slice."""
    from gpemsr_b200.volume import window_indices
    model, sd = build(8, seed=81, device=cuda_dev)
    vol = torch.rand(7, 1, 16, 20, generator=torch.Generator().manual_seed(82))
    got = model.forward_volume(vol.cuda())
    model.check()
    assert tuple(got.shape) == (7, 1, 128, 160)
    for i in range(7):
        win = vol[window_indices(i, 7)].unsqueeze(0)
        ref, _ = model(win.cuda())
        assert float((got[i:i + 1] - ref).abs().max()) <= 1e-4, i
        if i in (0, 3):
            with torch.no_grad():
                want, _ = GM.forward(win, sd, 8)
            assert float((got[i:i + 1].cpu() - want).abs().max()) <= 1e-3, i
    block = model.forward_volume(vol.cuda(), 2, 5)             # a rank's block: slices 2..4 with their halo
    assert float((block - got[2:5]).abs().max()) <= 1e-4


def test_batch_of_windows_equals_single_windows(cuda_dev):
    """forward(x[B > 1]) = the per-window results stacked (windows are independent; output_GPEMSR.py uses B = 1)."""
    model, _ = build(8, seed=83, device=cuda_dev)
    x = torch.rand(2, 5, 1, 16, 16, generator=torch.Generator().manual_seed(84)).cuda()
    out, ref = model(x)
    assert tuple(out.shape) == (2, 1, 128, 128) and tuple(ref.shape) == (2, 5, 1, 128, 128)
    for b in range(2):
        o1, r1 = model(x[b:b + 1])
        assert float((out[b:b + 1] - o1).abs().max()) <= 1e-4 and float((ref[b:b + 1] - r1).abs().max()) <= 1e-4


@pytest.mark.parametrize('scale,lr', [(16, 80), (8, 156)])
def test_full_size_windows(cuda_dev, scale, lr):
    """BASELINE configs[1] (x16, 5 x 80 x 80 -> 1280^2) and the CREMI x8 shape of configs[4] (5 x 156 x 156 -> 1248^2) at FULL size
    against the CPU oracle (about 10 s of host time each).  With ~30 000 latents per window some top-2 logits are closer than fp32
    summation-order noise, and the lookup is DISCRETE: the codebook indices are therefore judged like the VQ rows (every GPU
    index must be an arg-max of the oracle's logits up to 1e-3 of the logit range; disagreements must be rare), and everything
    downstream is compared with the oracle following the GPU's indices: HR image and reference images <= 1e-3 max-abs."""
    model, sd = build(scale, seed=85 + scale, device=cuda_dev)
    x = torch.rand(1, 5, 1, lr, lr, generator=torch.Generator().manual_seed(86 + scale))
    out, ref_img = model(x.cuda())
    model.check()
    idx = model.refmodel.codebook.last_idx.cpu()
    torch.set_num_threads(max(1, __import__('os').cpu_count() or 1))
    logits = []
    with torch.no_grad():
        want, want_ref = GM.forward(x, sd, scale, idx_override=idx, logits_out=logits)
    lg = logits[0].reshape(-1, logits[0].shape[-1])
    assert idx.numel() == lg.shape[0]
    top = lg.max(dim=1)
    regret = top.values - lg.gather(1, idx.view(-1, 1)).squeeze(1)
    flips = int((top.indices != idx).sum())
    print(f'x{scale} {lr}x{lr}: {flips} of {idx.numel()} indices differ from the oracle arg-max, max logit regret {float(regret.max()):.2e}, '
          f'logit range {float(lg.max() - lg.min()):.2f}')
    _record_index_parity(f'x{scale}_{lr}x{lr}_vs_cpu_oracle', dict(latents=int(idx.numel()), flips=flips, max_regret=float(regret.max()),
                                                                   logit_range=float(lg.max() - lg.min())))
    # 10x what was measured against the GPU-eager reference (tests/test_gpu_eager_parity_gpu.py, profiles/r02_index_parity.json:
    # 1 flip of 32 000, regret 2.8e-5); the un-overridden comparison on the frames without flips lives in that test
    assert float(regret.max()) <= 3e-4
    assert flips <= 10
    assert tuple(out.shape) == (1, 1, scale * lr, scale * lr)
    e, er = float((out.cpu() - want).abs().max()), float((ref_img.cpu() - want_ref).abs().max())
    print(f'out err {e:.3e}, ref_img err {er:.3e}')
    assert e <= 1e-3 and er <= 1e-3, (e, er)
    assert float(((out.cpu() - want) ** 2).mean()) < 1e-8


def test_three_slice_window_x8(cuda_dev):
    """BASELINE configs[0] says "3-slice 32 x 32 LR -> 256 x 256": the same model built with nframes = 3 (centre = 1; ThreeDA's
    Conv3d and 1x1 fusions are sized by the frame count), against the oracle."""
    model, sd = build(8, seed=87, device=cuda_dev, nframes=3)
    x = torch.rand(1, 3, 1, 32, 32, generator=torch.Generator().manual_seed(88))
    out, ref_img = model(x.cuda())
    model.check()
    with torch.no_grad():
        want, want_ref = GM.forward(x, sd, 8)
    assert tuple(out.shape) == (1, 1, 256, 256) and tuple(ref_img.shape) == (1, 3, 1, 256, 256)
    assert float((out.cpu() - want).abs().max()) <= 1e-3 and float((ref_img.cpu() - want_ref).abs().max()) <= 1e-3


def test_forward_volume_tiny_volumes(cuda_dev):
    """Edge cases of the slice loop: one slice (every window is that slice five times) and two slices."""
    from gpemsr_b200.volume import window_indices
    model, _ = build(8, seed=89, device=cuda_dev)
    for S in (1, 2):
        vol = torch.rand(S, 1, 16, 16, generator=torch.Generator().manual_seed(90 + S))
        got = model.forward_volume(vol.cuda())
        model.check()
        assert tuple(got.shape) == (S, 1, 128, 128)
        for i in range(S):
            ref, _ = model(vol[window_indices(i, S)].unsqueeze(0).cuda())
            assert float((got[i:i + 1] - ref).abs().max()) <= 1e-4, (S, i)
    assert tuple(model.forward_volume(torch.rand(3, 1, 16, 16).cuda(), 1, 1).shape) == (0, 1, 128, 128)       # empty block
