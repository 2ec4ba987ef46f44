"""CPU: the oracle restatement reproduces the golden vectors made by the reference's own modules."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from oracle import weights as W
from oracle.flow_warp import flow_warp_numpy, flow_warp_torch
from oracle.vq import classify

T = torch.from_numpy


def test_codebook_small_bitexact(golden):
    g = golden('codebook_small')
    zq, idx, loss = R.codebook_forward(T(g['z']), T(g['emb']))
    assert np.array_equal(idx.numpy(), g['idx'])
    assert np.array_equal(zq.numpy(), g['zq'])
    assert loss.item() == g['loss'].item()
    zq_lr, _ = R.codebook_inference_lr(T(g['logits']), T(g['emb']))
    assert np.array_equal(zq_lr.numpy(), g['zq_lr'])


def test_codebook_ties_lowest_index(golden):
    g = golden('codebook_small')
    zq, idx, loss = R.codebook_forward(T(g['z_t']), T(g['emb_t']))
    assert np.array_equal(idx.numpy(), g['idx_t'])
    assert np.array_equal(zq.numpy(), g['zq_t'])
    # the planted duplicates (rows 7/40/41 and 3/20) resolve to the lowest index
    idx = g['idx_t'].reshape(2, 5, 7)
    assert idx[0, 0, 0] == 7 and idx[0, 0, 1] == 3 and idx[1, 4, 6] == 7
    zq_lr, _ = R.codebook_inference_lr(T(g['logits_t']), T(g['emb_t']))
    assert np.array_equal(zq_lr.numpy(), g['zq_lr_t'])
    # fp64 classification agrees: integer data => every row is exact
    rows = g['z_t'].transpose(0, 2, 3, 1).reshape(-1, 32)
    c = classify(rows, g['emb_t'], g['idx_t'], tau=0.0)
    assert c['max_regret'] == 0.0


def test_codebook_reference_shape(golden):
    g = golden('codebook_1024x512')
    emb = W.fill(W.codebook_spec(1024, 512), seed=int(g['seeds'][0]))['embedding.weight']
    zq, idx, loss = R.codebook_forward(T(g['z']), emb)
    assert np.array_equal(idx.numpy(), g['idx'])
    assert np.array_equal(zq.numpy(), g['zq'])
    head = W.fill(W.indexer_head_spec(512, 1024), seed=int(g['seeds'][1]))
    logits = R.indexer_logits(T(g['feat']), head['embedding.weight'], head['embedding.bias'])
    zq_lr, idx_lr = R.codebook_inference_lr(logits, emb)
    assert np.array_equal(idx_lr.numpy(), g['idx_lr'])
    assert np.array_equal(zq_lr.numpy(), g['zq_lr'])
    # H2: top-1 of softmax == argmax of the logits on this data
    assert np.array_equal(logits.reshape(-1, 1024).argmax(1).numpy(), g['idx_lr'])


def test_blocks_bitexact(golden):
    g = golden('blocks_small')
    s41, s43, s44 = (int(s) for s in g['seeds'])
    spec = W.OrderedDict(); W._resblock(spec, 'rb', 32, 64)
    sd = {k[3:]: v for k, v in W.fill(spec, s41).items()}
    assert np.array_equal(R.residual_block(T(g['x']), sd).numpy(), g['rb'])
    spec_u = W.OrderedDict([('upblock.weight', ('convT', (32, 64, 3, 3))), ('upblock.bias', ('bias', (64,)))])
    assert np.array_equal(R.up_block(T(g['x']), W.fill(spec_u, s43)).numpy(), g['up'])
    spec_n = W.OrderedDict(); W._nonlocal(spec_n, 'nl', 64)
    sdn = {k[3:]: v for k, v in W.fill(spec_n, s44).items()}
    assert np.array_equal(R.non_local_block(T(g['xn']), sdn).numpy(), g['nl'])


def test_decoder_small_bitexact(golden):
    g = golden('decoder_small')
    sd = W.fill(W.decoder_spec([64, 64, 32, 32, 32], 64, 2, 1, True, 1), seed=int(g['seed'][0]))
    feats = R.decoder_multi_scale(T(g['x']), sd, num_input_resblck=2)
    assert len(feats) == 5
    for i, f in enumerate(feats):
        assert np.array_equal(f.numpy(), g[f'feat{i}']), i


def test_decoder_full_width(golden):
    g = golden('decoder_full_4x4')
    sd = W.fill(W.decoder_spec(), seed=int(g['seed'][0]))
    feats = R.decoder_multi_scale(T(g['x']), sd)
    assert [tuple(f.shape) for f in feats] == [(1, 512, 4, 4), (1, 256, 8, 8), (1, 128, 16, 16),
                                                (1, 64, 32, 32), (1, 1, 64, 64)]
    for i, f in enumerate(feats):
        assert np.array_equal(f.numpy(), g[f'feat{i}']), i


def test_tail_bitexact(golden):
    for scale in (8, 16):
        g = golden(f'tail_x{scale}')
        sd = W.fill(W.tail_spec(64, 10, scale), seed=int(g['seed'][0]), gain=3.0 ** 0.5)
        out = R.sr_tail(T(g['fea']), T(g['x_center']), sd, scale)
        assert np.array_equal(out.numpy(), g['out']), scale
        assert np.abs(g['out']).max() > 0.1      # the fixture is not degenerate


def test_flow_warp_golden(golden):
    g = golden('flow_warp_small')
    for pm in ('border', 'zeros'):
        a = flow_warp_torch(T(g['x']), T(g['flow']), 'bilinear', pm).numpy()
        assert np.array_equal(a, g['out_' + pm])
        b = flow_warp_numpy(g['x'], g['flow'], 'bilinear', pm)
        assert np.abs(a - b).max() <= 1e-6       # spelled-out arithmetic vs ATen: a few ulp
    b = flow_warp_numpy(g['x2'], g['flow2'], 'bilinear', 'border')
    assert np.abs(b - g['out2_border']).max() <= 1e-6


def test_flow_warp_kats():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 2, 6, 9)).astype(np.float32)
    zero = np.zeros((1, 6, 9, 2), np.float32)
    for pm in ('border', 'zeros'):
        # not bit-exact even in the reference: the normalise -> unnormalise round trip perturbs coordinates
        assert np.allclose(flow_warp_numpy(x, zero, padding_mode=pm), x, atol=2e-6)
    sh = zero.copy(); sh[..., 0] = 2.0; sh[..., 1] = -1.0          # sample from (x+2, y-1)
    out_b = flow_warp_numpy(x, sh, padding_mode='border')
    out_z = flow_warp_numpy(x, sh, padding_mode='zeros')
    exp_b = x[:, :, np.clip(np.arange(6) - 1, 0, 5)][:, :, :, np.clip(np.arange(9) + 2, 0, 8)]
    assert np.allclose(out_b, exp_b, atol=2e-6)
    exp_z = exp_b.copy(); exp_z[:, :, 0, :] = 0; exp_z[:, :, :, 7:] = 0
    assert np.allclose(out_z, exp_z, atol=2e-6)


INDEXER_CASES = [('i16', 16, [32, 32, 64, 64, 64], 0), ('i8', 8, [32, 32, 64, 64, 64], 1), ('i16up', 16, [32, 64, 64, 64], 2)]


def test_indexer_small_bitexact(golden):
    """Indexer16 / Indexer8 conv stacks (model/indexer.py), incl. DownBlock on an odd-sized input and the UpBlock tail."""
    g = golden('indexer_small')
    for tag, variant, cl, si in INDEXER_CASES:
        sd = W.fill(W.indexer_spec(variant, cl, 1, 2, 1, 64, True), seed=int(g['seeds'][si]))
        feat = R.indexer_features(T(g[f'{tag}_x']), sd)
        assert np.array_equal(feat.numpy(), g[f'{tag}_feat']), tag
        assert np.array_equal(R.indexer_forward(T(g[f'{tag}_x']), sd).numpy(), g[f'{tag}_logits']), tag


def test_vgg_mask_bitexact(golden):
    """relu1_2 patch-similarity mask (model/GPEMSR.py:344-353 with the reference's VGG19 + extract_image_patches)."""
    g = golden('vgg_mask_small')
    sd = W.fill(W.vgg_slice1_spec(), seed=int(g['seed'][0]))
    r12 = R.vgg_relu1_2(T(g['ref_img']).expand(-1, 3, -1, -1), sd)
    assert np.array_equal(r12[:1, :, :16, :16].numpy(), g['relu1_2'])
    mask = R.similarity_mask(T(g['ref_img']), T(g['x']), sd, 16)
    assert np.array_equal(mask.numpy(), g['mask'])


@pytest.mark.parametrize('scale', [8, 16])
def test_full_model_restatement_bitexact(golden, scale):
    """oracle/gpemsr_model.py (GPEMSR.forward + POD + ThreeDA restated) == the unmodified reference model/GPEMSR.py on the
    committed window; the parameter names / shapes of the mirror module were asserted equal to the reference's at generation."""
    from oracle import gpemsr_model as GM
    from full_model_util import build
    g = golden(f'full_x{scale}')
    _, sd = build(scale)
    with torch.no_grad():
        out, ref_img = GM.forward(T(g['x']), sd, scale)
    assert np.array_equal(out.numpy(), g['out'])
    assert np.array_equal(ref_img[0, :, 0, ::4, ::4].numpy(), g['ref_img_sub'])
