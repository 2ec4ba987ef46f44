"""CPU: slice sharding and the output gather (world_size 2, gloo)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpemsr_b200.volume import gather_slices, shard_range, super_resolve_volume, window_indices


def test_shard_range_partitions():
    for n, w in ((125, 8), (125, 1), (7, 8), (0, 3), (16, 4)):
        blocks = [shard_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert [b - a for a, b in (shard_range(125, 8, r) for r in range(8))] == [16] * 5 + [15] * 3
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_window_indices_replicate_edges():
    assert window_indices(0, 125) == [0, 0, 0, 1, 2]          # output_GPEMSR.py:55-60
    assert window_indices(1, 125) == [0, 0, 1, 2, 3]          # :71-76
    assert window_indices(60, 125) == [58, 59, 60, 61, 62]
    assert window_indices(123, 125) == [121, 122, 123, 124, 124]   # :100-105
    assert window_indices(124, 125) == [122, 123, 124, 124, 124]   # :116-121


def _worker(rank, world, port, n_units):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(n_units, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1) * torch.ones(1, 2, 3)
    full = gather_slices(local, n_units, world, rank, dist)
    assert full.shape == (n_units, 2, 3)
    assert torch.equal(full[:, 0, 0], torch.arange(n_units, dtype=torch.float32))
    dist.barrier()
    dist.destroy_process_group()


class _FakeModel:
    """Stands in for gpemsr_b200.GPEMSR on CPU: slice i of the output is a known function of its window's slice indices."""

    def forward_volume(self, vol, lo=0, hi=None):
        hi = vol.shape[0] if hi is None else hi
        rows = [sum(float(vol[j, 0, 0, 0]) * (k + 1) for k, j in enumerate(window_indices(i, vol.shape[0]))) for i in range(lo, hi)]
        return torch.tensor(rows).view(-1, 1, 1, 1) * torch.ones(1, 1, 2, 2)


def _volume_worker(rank, world, port, n_slices):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    vol = torch.arange(n_slices, dtype=torch.float32).view(-1, 1, 1, 1) * torch.ones(1, 1, 4, 4)
    full = super_resolve_volume(_FakeModel(), vol, rank, world, dist)
    want = _FakeModel().forward_volume(vol)
    assert full.shape == want.shape and torch.equal(full, want)
    dist.barrier()
    dist.destroy_process_group()


def test_super_resolve_volume_two_ranks_gloo():
    """N > 1 path of the volume driver: contiguous slice blocks per rank, one all_gather, no other collective."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_volume_worker, args=(2, port, 9), nprocs=2, join=True)


def test_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, 7), nprocs=2, join=True)


def test_compose_upblock_conv_matches_torch():
    """Host-side weight preprocessing: ConvTranspose2d(k3,s2,p1,op1) followed by a zero-padded 3x3 conv, composed into a
    4-phase 3x3 map with per-border-class weights, equals the two-step evaluation (CPU, float64 reference)."""
    import torch.nn.functional as F
    from gpemsr_b200.igemm import compose_upblock_conv
    torch.manual_seed(0)
    ci, c, co = 5, 6, 2
    wt, bu = torch.randn(ci, c, 3, 3, dtype=torch.float64), torch.randn(c, dtype=torch.float64)
    wo, bo = torch.randn(co, c, 3, 3, dtype=torch.float64), torch.randn(co, dtype=torch.float64)
    x = torch.randn(1, ci, 4, 5, dtype=torch.float64)
    ref = F.conv2d(F.conv_transpose2d(x, wt, bu, 2, 1, 1), wo, bo, 1, 1)
    wc, bias = compose_upblock_conv(wt, bu, wo, bo)
    wc, bias = wc.double(), bias.double()
    H, W = 4, 5
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for oy in range(2 * H):
        for ox in range(2 * W):
            ry = 0 if oy == 0 else (2 if oy == 2 * H - 1 else 1)
            rx = 0 if ox == 0 else (2 if ox == 2 * W - 1 else 1)
            a, b, ph = oy // 2, ox // 2, (oy & 1) * 2 + (ox & 1)
            out[0, :, oy, ox] = bias[ry * 3 + rx] + torch.einsum('oikl,ikl->o', wc[ry * 3 + rx, ph], xp[0, :, a:a + 3, b:b + 3])
    assert (out - ref).abs().max().item() < 1e-5
