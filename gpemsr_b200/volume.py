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


def exchange_halo(send_down, recv_down, send_up, recv_up, rank, world_size, dist):
    """Neighbour exchange of per-slice feature maps between slice blocks (one batched point-to-point group, NCCL over NVLink
    on the GPU box, gloo in the CPU tests).  send_down / recv_down: lists of tensors (any strides) going to / coming from rank - 1
    (this rank's FIRST halo-many encoded slices; rank - 1's LAST ones), send_up / recv_up likewise for rank + 1; the lists of
    two neighbours must pair up element by element.  Ranks at the ends pass empty lists for the missing side."""
    if world_size == 1 or dist is None:
        return
    ops, landing = [], []
    for peer, sends, recvs in ((rank - 1, send_down, recv_down), (rank + 1, send_up, recv_up)):
        if not (0 <= peer < world_size):
            continue
        for t in sends:
            ops.append(dist.P2POp(dist.isend, t.contiguous(), peer))
        for t in recvs:
            buf = t if t.is_contiguous() else torch_empty_like(t)
            landing.append((t, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for t, buf in landing:
        if buf is not t:
            t.copy_(buf)


def torch_empty_like(t):
    import torch
    return torch.empty(t.shape, dtype=t.dtype, device=t.device)


def super_resolve_volume(model, vol, rank=0, world_size=1, dist=None, gather=True, halo='exchange'):
    """The reference's output loop (output_GPEMSR.py:54-128) over a whole LR volume vol [S, 1, H, W], slice-sharded:
    rank r super-resolves the contiguous block ``shard_range(S, world_size, r)`` with ``model.forward_volume``.  The windows at
    a block's ends need the per-slice features of 2 slices of the neighbouring blocks: halo='exchange' (default) fetches them from
    the neighbour ranks (one point-to-point exchange of ~17 MB per slice over NVLink, ~0.2 ms), halo='recompute' encodes them
    again locally (every rank holds the whole LR volume; no transfer at all, but 4 extra slice encodings of ~6 ms per rank:
    the 0.86 strong-scaling efficiency of round 1 at N = 8).  If `gather`, the HR slices are all-gathered once.
    Returns [S, 1, sH, sW] (or the local block if not gather)."""
    lo, hi = shard_range(vol.shape[0], world_size, rank)
    ex = (dist, rank, world_size) if (halo == 'exchange' and world_size > 1 and dist is not None) else None
    local = model.forward_volume(vol, lo, hi, halo_exchange=ex) if ex else model.forward_volume(vol, lo, hi)
    out = gather_slices(local, vol.shape[0], world_size, rank, dist) if gather else local
    if local.is_cuda:
        # a volume is a unit of work whose result leaves the GPU next: wait for the pipeline-error read-back here (one host
        # sync per volume) so a timed-out GEMM pipeline raises instead of returning invalid slices
        from . import igemm as G
        G.poll_error(local.device, wait=True)
    return out
