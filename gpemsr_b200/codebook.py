"""Host-side mirror of the reference ``Codebook`` (model/codebook.py) on the sm_100a lookup kernel.

Same constructor arguments, parameter name (``embedding.weight``, so reference checkpoints load with
``strict=True``) and method signatures as the reference:

  * ``forward(z)``        -> ``(z_q, min_encoding_indices, loss)``          (model/codebook.py:15-32)
  * ``inference_lr(p)``   -> ``z_q``                                        (model/codebook.py:34-43)

plus the fused form the reference spreads over two modules:

  * ``inference_from_feat(feat, weight, bias)`` = ``inference_lr(Linear(feat.permute(0,2,3,1)))``
    (model/indexer.py:47,53 / 96,100 + model/codebook.py:34-43) without materialising the logits.

Inference only (the reference runs this path under ``torch.no_grad()``, output_GPEMSR.py:49).  ``forward``
returns ``E[idx]`` exactly; the reference's straight-through expression ``z + (z_q - z)`` (:28) differs from
that by at most one rounding of ``z_q - z`` (documented in DESIGN.md).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

_ws_cache = {}
_ws_retired = []        # outgrown buffers stay allocated: a captured CUDA graph may have their pointers baked in


def workspace(nbytes, device):
    """Scratch buffer per (device, stream), grown on demand (the C ABI never allocates).  A buffer that was outgrown is kept
    alive rather than freed, so replaying a CUDA graph captured with it can never write into recycled memory; streams do not
    share a buffer, so concurrent lookups on two streams do not race on it."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _check(t, name):
    if not t.is_cuda:
        raise _lib.GpemsrError(-3, f'{name} must be a CUDA tensor: there is no CPU fallback')
    if t.dtype != torch.float32:
        raise TypeError(f'{name}: fp32 only (the reference path is fp32), got {t.dtype}')


def vq_lookup(z, emb, want_sq_err=False):
    """z f32[B, D, H, W], emb f32[K, D] -> (z_q f32[B, D, H, W], idx int64[B*H*W], sq_err f32[] or None)."""
    _check(z, 'z'); _check(emb, 'emb')
    B, D, H, W = z.shape
    K, D2 = emb.shape
    if D != D2:
        raise ValueError(f'latent_dim mismatch: z has {D} channels, codebook has {D2}')
    z = z.contiguous(); emb = emb.contiguous()
    L = _lib.lib()
    zq = torch.empty_like(z)
    idx = torch.empty(B * H * W, dtype=torch.int64, device=z.device)
    sq = torch.zeros((), dtype=torch.float32, device=z.device) if want_sq_err else None
    nb = L.gpemsr_vq_workspace_bytes(B * H * W, D, K)
    ws = workspace(nb, z.device)
    _lib.check(L.gpemsr_vq_lookup_nchw(_lib.ptr(z), _lib.ptr(emb), B, D, H * W, K, _lib.ptr(zq), _lib.ptr(idx),
                                       _lib.ptr(sq), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return zq, idx, sq


def logits_argmax_gather(feat, weight, bias, emb):
    """feat f32[B, D, H, W]; weight f32[K, D]; bias f32[K]; emb f32[K, Dq] -> (z_q f32[B, Dq, H, W], idx int64[B*H*W])."""
    for t, n in ((feat, 'feat'), (weight, 'weight'), (bias, 'bias'), (emb, 'emb')):
        _check(t, n)
    B, D, H, W = feat.shape
    K, D2 = weight.shape
    if D != D2 or bias.shape != (K,) or emb.shape[0] != K:
        raise ValueError('logits_argmax_gather: inconsistent shapes')
    Dq = emb.shape[1]
    feat = feat.contiguous(); weight = weight.contiguous(); bias = bias.contiguous(); emb = emb.contiguous()
    L = _lib.lib()
    zq = torch.empty(B, Dq, H, W, dtype=torch.float32, device=feat.device)
    idx = torch.empty(B * H * W, dtype=torch.int64, device=feat.device)
    nb = L.gpemsr_vq_workspace_bytes(B * H * W, D, K)
    ws = workspace(nb, feat.device)
    _lib.check(L.gpemsr_logits_argmax_gather(_lib.ptr(feat), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(emb), B, D, H * W, K,
                                             Dq, _lib.ptr(zq), _lib.ptr(idx), None, _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr()))
    return zq, idx


def argmax_gather(p, emb):
    """p f32[B, H, W, K] logits; emb f32[K, Dq] -> (z_q f32[B, Dq, H, W], idx int64[B*H*W])."""
    _check(p, 'p'); _check(emb, 'emb')
    B, H, W, K = p.shape
    if emb.shape[0] != K:
        raise ValueError('argmax_gather: inconsistent shapes')
    Dq = emb.shape[1]
    p = p.contiguous(); emb = emb.contiguous()
    zq = torch.empty(B, Dq, H, W, dtype=torch.float32, device=p.device)
    idx = torch.empty(B * H * W, dtype=torch.int64, device=p.device)
    _lib.check(_lib.lib().gpemsr_argmax_gather(_lib.ptr(p), _lib.ptr(emb), B, H * W, K, Dq, _lib.ptr(zq), _lib.ptr(idx),
                                               _lib.stream_ptr()))
    return zq, idx


class Codebook(nn.Module):
    """Drop-in for ``model.codebook.Codebook`` (same ``args`` dict, same parameter name)."""

    def __init__(self, args):
        super().__init__()
        self.num_codebook_vectors = args['num_codebook_vectors']
        self.latent_dim = args['latent_dim']
        self.beta = args['beta']
        self.embedding = nn.Embedding(self.num_codebook_vectors, self.latent_dim)
        self.embedding.weight.data.uniform_(-1.0 / self.num_codebook_vectors, 1.0 / self.num_codebook_vectors)

    @torch.no_grad()
    def forward(self, z):
        zq, idx, sq = vq_lookup(z, self.embedding.weight.detach(), want_sq_err=True)
        m = sq / z.numel()
        return zq, idx, m + self.beta * m

    @torch.no_grad()
    def inference_lr(self, p):
        return argmax_gather(p, self.embedding.weight.detach())[0]

    @torch.no_grad()
    def inference_from_feat(self, feat, weight, bias):
        zq, self.last_idx = logits_argmax_gather(feat, weight, bias, self.embedding.weight.detach())[:2]     # indices kept for the tests
        return zq
