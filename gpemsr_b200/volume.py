"""Slice-sharded volume driver: independent units, no collective on the hot path.

The unit of work is one OUTPUT slice, which needs the LR slices i-2 .. i+2 with replicate padding at the volume ends
(the window construction of output_GPEMSR.py:54-128).  Output slices are split into contiguous blocks, one per rank;
every rank holds the whole (tiny) LR volume, so there is no halo exchange.  NCCL (or gloo on CPU) is used once, to
gather the HR slices.
"""
from __future__ import annotations


def shard_range(n_units, world_size, rank):
    """Contiguous block [lo, hi) of `n_units` for `rank`; sizes differ by at most one (e.g. 125 over 8 -> 16x5 + 15x3)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank {rank} / world {world_size}')
    base, extra = divmod(n_units, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def window_indices(i, n_slices, n_frames=5):
    """LR slice indices feeding output slice i: centre i, replicate padding at the ends (output_GPEMSR.py:54-128)."""
    half = n_frames // 2
    return [min(max(i + d, 0), n_slices - 1) for d in range(-half, half + 1)]


def gather_slices(local, n_units, world_size, rank, dist=None):
    """All ranks contribute their [hi-lo, ...] block; returns the [n_units, ...] stack on every rank.
    Blocks are padded to the largest block so one all_gather suffices."""
    import torch
    if world_size == 1 or dist is None:
        return local
    per = max(shard_range(n_units, world_size, r)[1] - shard_range(n_units, world_size, r)[0] for r in range(world_size))
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world_size):
        lo, hi = shard_range(n_units, world_size, r)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts, 0)


def super_resolve_volume(model, vol, rank=0, world_size=1, dist=None, gather=True):
    """The reference's output loop (output_GPEMSR.py:54-128) over a whole LR volume vol [S, 1, H, W], slice-sharded:
    rank r super-resolves the contiguous block ``shard_range(S, world_size, r)`` with ``model.forward_volume`` (which encodes
    the block's slices plus a 2-slice halo once each -- no halo exchange: every rank holds the LR volume) and, if `gather`,
    the HR slices are all-gathered once (the only collective).  Returns [S, 1, sH, sW] (or the local block if not gather)."""
    lo, hi = shard_range(vol.shape[0], world_size, rank)
    local = model.forward_volume(vol, lo, hi)
    out = gather_slices(local, vol.shape[0], world_size, rank, dist) if gather else local
    if local.is_cuda:
        # a volume is a unit of work whose result leaves the GPU next: wait for the pipeline-error read-back here (one host
        # sync per volume) so a timed-out GEMM pipeline raises instead of returning invalid slices
        from . import igemm as G
        G.poll_error(local.device, wait=True)
    return out
