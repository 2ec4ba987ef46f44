"""GPU parity: codebook lookup (a-1 Codebook.forward, a-2 Indexer head + inference_lr) through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from oracle import weights as W
from oracle.vq import classify

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _rows(z):
    b, d, h, w = z.shape
    return np.ascontiguousarray(z.transpose(0, 2, 3, 1).reshape(-1, d))


def _check_idx(z_np, emb_np, got_idx, ref_idx=None):
    """Exact on every row whose fp64 margin is clear; bounded fp64 regret elsewhere."""
    c = classify(_rows(z_np), emb_np, got_idx)
    assert c['clear_mismatch'] == 0, c
    assert c['max_regret'] <= c['tau'], c
    if ref_idx is not None:
        cr = classify(_rows(z_np), emb_np, ref_idx)       # the reference's own fp32 result obeys the same contract
        assert cr['clear_mismatch'] == 0
    return c


def test_golden_small(golden, cuda_dev):
    import gpemsr_b200
    g = golden('codebook_small')
    zq, idx, sq = gpemsr_b200.vq_lookup(T(g['z']).cuda(), T(g['emb']).cuda(), want_sq_err=True)
    idx = idx.cpu().numpy()
    _check_idx(g['z'], g['emb'], idx, g['idx'])
    assert np.array_equal(idx, g['idx'])                  # well-separated random data: identical to the reference
    assert np.array_equal(zq.cpu().numpy(), g['emb'][idx].reshape(2, 5, 7, 32).transpose(0, 3, 1, 2))
    assert np.abs(zq.cpu().numpy() - g['zq']).max() <= 2.4e-7        # reference z + (z_q - z): <= 1 rounding
    loss = (sq / g['z'].size * 2.0).item()
    assert abs(loss - g['loss'].item()) <= 1e-5 * abs(g['loss'].item())


def test_golden_ties_lowest_index(golden, cuda_dev):
    import gpemsr_b200
    g = golden('codebook_small')
    zq, idx, _ = gpemsr_b200.vq_lookup(T(g['z_t']).cuda(), T(g['emb_t']).cuda())
    assert np.array_equal(idx.cpu().numpy(), g['idx_t'])  # integer data: every evaluation order is exact
    assert np.array_equal(zq.cpu().numpy(), g['zq_t'])
    zq_lr, idx_lr = gpemsr_b200.argmax_gather(T(g['logits_t']).cuda(), T(g['emb_t']).cuda())
    assert np.array_equal(zq_lr.cpu().numpy(), g['zq_lr_t'])
    zq_lr, _ = gpemsr_b200.argmax_gather(T(g['logits']).cuda(), T(g['emb']).cuda())
    assert np.array_equal(zq_lr.cpu().numpy(), g['zq_lr'])


def test_golden_reference_shape(golden, cuda_dev):
    import gpemsr_b200
    g = golden('codebook_1024x512')
    emb = W.fill(W.codebook_spec(1024, 512), seed=int(g['seeds'][0]))['embedding.weight']
    zq, idx, sq = gpemsr_b200.vq_lookup(T(g['z']).cuda(), emb.cuda(), want_sq_err=True)
    idx = idx.cpu().numpy()
    c = _check_idx(g['z'], emb.numpy(), idx, g['idx'])
    # rows where we differ from the reference must be fp32-ambiguous ones
    diff = idx != g['idx']
    assert diff.sum() <= c['n'] - c['clear_rows']
    assert np.array_equal(zq.cpu().numpy().transpose(0, 2, 3, 1).reshape(-1, 512), emb.numpy()[idx])
    # a-2: fused Linear + argmax + gather
    head = W.fill(W.indexer_head_spec(512, 1024), seed=int(g['seeds'][1]))
    zq_lr, idx_lr = gpemsr_b200.logits_argmax_gather(T(g['feat']).cuda(), head['embedding.weight'].cuda(),
                                                     head['embedding.bias'].cuda(), emb.cuda())
    idx_lr = idx_lr.cpu().numpy()
    logits64 = _rows(g['feat']).astype(np.float64) @ head['embedding.weight'].numpy().astype(np.float64).T \
        + head['embedding.bias'].numpy().astype(np.float64)
    best = logits64.argmax(1)
    top2 = np.partition(logits64, -2, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-5
    assert np.array_equal(idx_lr[clear], best[clear])
    assert (logits64[np.arange(len(best)), best] - logits64[np.arange(len(best)), idx_lr]).max() <= 1e-5
    assert np.array_equal(idx_lr, g['idx_lr'])
    assert np.array_equal(zq_lr.cpu().numpy(), g['zq_lr'])


@pytest.mark.parametrize('n_d', [(1024, 256), (4096, 512), (65536, 256), (65536, 512), (1000, 512), (130, 40)])
def test_vs_oracle_microbench_shapes(n_d, cuda_dev):
    """BASELINE config 3 inputs: z = randn(N, D) seed 1, E = randn(1024, D) seed 2 (as NCHW [1, D, N, 1])."""
    import gpemsr_b200
    n, d = n_d
    z = torch.randn(n, d, generator=torch.Generator().manual_seed(1))
    emb = torch.randn(1024, d, generator=torch.Generator().manual_seed(2))
    z4 = z.t().contiguous().view(1, d, n, 1)
    zq, idx, _ = gpemsr_b200.vq_lookup(z4.cuda(), emb.cuda())
    idx = idx.cpu().numpy()
    zq_ref, idx_ref, _ = R.codebook_forward(z4, emb)
    c = _check_idx(z4.numpy(), emb.numpy(), idx, idx_ref.numpy())
    assert (idx != idx_ref.numpy()).sum() <= c['n'] - c['clear_rows']
    assert np.array_equal(zq.cpu().numpy()[0, :, :, 0].T, emb.numpy()[idx])


def test_default_init_codebook_and_duplicates(cuda_dev):
    """U(+-1/K) codebook (model/codebook.py:13) with duplicated rows: near-ties everywhere, overflow path included."""
    import gpemsr_b200
    k, d = 1024, 512
    emb = W.fill(W.codebook_spec(k, d), seed=5)['embedding.weight'].clone()
    emb[100:110] = emb[7]                 # 11 identical codes -> more than CMAX candidates in one tile
    emb[900] = emb[300]
    z = torch.randn(2, d, 16, 16, generator=torch.Generator().manual_seed(6))
    z[0, :, 0, 0] = emb[7] * 3.0          # closest to the duplicated group
    z[0, :, 0, 1] = emb[300]
    zq, idx, _ = gpemsr_b200.vq_lookup(z.cuda(), emb.cuda())
    idx = idx.cpu().numpy()
    assert idx[1] == 300
    assert idx[0] == 7
    _check_idx(z.numpy(), emb.numpy(), idx)
    # the reference smoke input (model/codebook.py:55): all-zero z
    z0 = torch.zeros(4, d, 10, 10)
    zq0, idx0, _ = gpemsr_b200.vq_lookup(z0.cuda(), emb.cuda())
    _, idx0_ref, _ = R.codebook_forward(z0, emb)
    assert np.array_equal(idx0.cpu().numpy(), idx0_ref.numpy())


def test_codebook_module_api(cuda_dev):
    import gpemsr_b200
    args = {'num_codebook_vectors': 1024, 'latent_dim': 512, 'beta': 1}
    cb = gpemsr_b200.Codebook(args).cuda().eval()
    assert list(cb.state_dict().keys()) == ['embedding.weight']
    # reference smoke (model/codebook.py:45-56)
    assert cb.inference_lr(torch.zeros(4, 10, 10, 1024, device='cuda')).shape == (4, 512, 10, 10)
    zq, idx, loss = cb(torch.zeros(4, 512, 10, 10, device='cuda'))
    assert zq.shape == (4, 512, 10, 10) and idx.shape == (400,) and idx.dtype == torch.int64 and loss.dim() == 0
    z = torch.randn(2, 512, 6, 5, generator=torch.Generator().manual_seed(8))
    zq, idx, loss = cb(z.cuda())
    zq_r, idx_r, loss_r = R.codebook_forward(z, cb.embedding.weight.detach().cpu(), 1.0)
    assert abs(loss.item() - loss_r.item()) <= 1e-5 * loss_r.item()
    _check_idx(z.numpy(), cb.embedding.weight.detach().cpu().numpy(), idx.cpu().numpy(), idx_r.numpy())
    # empty batch
    zq, idx, loss = cb(torch.zeros(0, 512, 4, 4, device='cuda'))
    assert zq.shape == (0, 512, 4, 4) and idx.numel() == 0
    with pytest.raises(gpemsr_b200.GpemsrError):
        gpemsr_b200.vq_lookup(torch.zeros(1, 512, 2, 2), torch.zeros(1024, 512))


def test_full_size_properties(cuda_dev):
    """N = 2^20 rows x 512 (BASELINE config 3 upper end): idempotence + fp64 spot check."""
    import gpemsr_b200
    n, d = 1 << 20, 512
    g = torch.Generator(device='cuda').manual_seed(1)
    z = torch.randn(1, d, n, 1, device='cuda', generator=g)
    emb = torch.randn(1024, d, device='cuda', generator=g)
    zq, idx, _ = gpemsr_b200.vq_lookup(z, emb)
    assert int(idx.min()) >= 0 and int(idx.max()) < 1024
    zq2, idx2, _ = gpemsr_b200.vq_lookup(zq, emb)             # quantising a code vector returns the same code
    assert torch.equal(idx2, idx) and torch.equal(zq2, zq)
    sel = torch.randint(0, n, (4096,), generator=torch.Generator().manual_seed(3))
    rows = z[0, :, :, 0].t()[sel.cuda()].cpu().numpy()
    c = classify(rows, emb.cpu().numpy(), idx[sel.cuda()].cpu().numpy())
    assert c['clear_mismatch'] == 0 and c['max_regret'] <= c['tau'], c
