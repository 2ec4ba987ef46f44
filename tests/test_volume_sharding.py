"""CPU: slice sharding and the output gather (world_size 2, gloo)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpemsr_b200.volume import exchange_halo, gather_slices, shard_range, super_resolve_volume, window_indices


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
    """Stands in for gpemsr_b200.GPEMSR on CPU with the same two-phase structure as ``forward_volume``: every needed slice is
    "encoded" once into a feature bank (here: 10 * slice value + 1, in a strided bank like the real [cells][rows][8] planes), each
    output slice is a known function of its window's BANK entries.  With ``halo_exchange`` the rank encodes its own block only
    and the halo entries must arrive from the neighbours (gpemsr_b200.volume.exchange_halo) -- a missing or misplaced halo
    shows up as a wrong output."""

    def __init__(self):
        self.encoded = []

    def forward_volume(self, vol, lo=0, hi=None, halo_exchange=None):
        S = vol.shape[0]
        hi = S if hi is None else hi
        f_lo, f_hi = max(lo - 2, 0), min(hi + 2, S)
        e_lo, e_hi = (lo, hi) if halo_exchange else (f_lo, f_hi)
        bank = torch.full((3, f_hi - f_lo, 2), float('nan'))           # [planes][slots][..]: slot views are strided
        for s_ in range(e_lo, e_hi):
            bank[:, s_ - f_lo] = 10.0 * vol[s_, 0, 0, 0] + 1.0
            self.encoded.append(s_)
        if halo_exchange:
            dist_, rank, world = halo_exchange
            sl = lambda a, n: [bank[:, a - f_lo:a - f_lo + n]]
            down, up = lo > 0, hi < S
            exchange_halo(sl(lo, 2) if down else [], sl(lo - 2, 2) if down else [], sl(hi - 2, 2) if up else [],
                          sl(hi, 2) if up else [], rank, world, dist_)
        rows = [sum(float(bank[0, j - f_lo, 0]) * (k + 1) for k, j in enumerate(window_indices(i, S))) for i in range(lo, hi)]
        return torch.tensor(rows).view(-1, 1, 1, 1) * torch.ones(1, 1, 2, 2)


def _volume_worker(rank, world, port, n_slices):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    vol = torch.arange(n_slices, dtype=torch.float32).view(-1, 1, 1, 1) * torch.ones(1, 1, 4, 4)
    want = _FakeModel().forward_volume(vol)
    lo, hi = shard_range(n_slices, world, rank)
    for halo in ('exchange', 'recompute'):
        m = _FakeModel()
        full = super_resolve_volume(m, vol, rank, world, dist, halo=halo)
        assert full.shape == want.shape and torch.equal(full, want), halo
        # exchange: a rank encodes exactly its own block; recompute: its block plus the 2-slice halo
        assert m.encoded == (list(range(lo, hi)) if halo == 'exchange' else list(range(max(lo - 2, 0), min(hi + 2, n_slices))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_super_resolve_volume_gloo(world):
    """N > 1 path of the volume driver: contiguous slice blocks per rank, the halo exchanged between neighbours (a middle rank has
    two) or recomputed, one all_gather."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_volume_worker, args=(world, port, 11), nprocs=world, join=True)


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
