"""CPU restatement of BasicSR ``flow_warp`` (+ ATen ``grid_sampler_2d``).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned at the
BasicSR boundary**: ``basicsr`` (PyPI, upstream XPixelGroup/BasicSR, no
version pinned by the reference; v1.4.2 restated here) is imported by the
reference at ``model/GPEMSR.py:4,7,8`` and is absent from /root/reference and
from this image.  Its published algorithm, ``basicsr/archs/arch_util.py::flow_warp``:

    grid  = stack(meshgrid(arange(h), arange(w))) as (x, y), float
    vgrid = grid + flow                                   # flow[..., 0] = dx, flow[..., 1] = dy (pixels)
    gx    = 2.0 * vgrid[..., 0] / max(w - 1, 1) - 1.0
    gy    = 2.0 * vgrid[..., 1] / max(h - 1, 1) - 1.0
    out   = F.grid_sample(x, stack(gx, gy), mode=interp_mode,
                          padding_mode=padding_mode, align_corners=align_corners)

Call site on the hot path: ``SpyNet.process`` -> ``flow_warp(supp, flow.permute(0,2,3,1),
interp_mode='bilinear', padding_mode='border')`` (reached from model/GPEMSR.py:99-100).

Two implementations live here:
  * ``flow_warp_torch``  -- the algorithm above on torch CPU (``F.grid_sample`` is
    ATen, present in this image): this is what the reference executes.
  * ``flow_warp_numpy``  -- every step spelled out in numpy fp32 (or fp64),
    following ATen ``GridSampler.h`` (``grid_sampler_unnormalize``,
    ``clip_coordinates``, bilinear corner weights nw/ne/sw/se); pinned against
    ``flow_warp_torch`` by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np


def flow_warp_torch(x, flow, interp_mode='bilinear', padding_mode='zeros', align_corners=True):
    import torch
    import torch.nn.functional as F
    assert x.shape[-2:] == flow.shape[1:3]
    _, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(0, h).type_as(x), torch.arange(0, w).type_as(x), indexing='ij')
    grid = torch.stack((gx, gy), 2).float()
    v = grid + flow
    vx = 2.0 * v[:, :, :, 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * v[:, :, :, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((vx, vy), dim=3), mode=interp_mode,
                         padding_mode=padding_mode, align_corners=align_corners)


def source_coords(flow, h, w, padding_mode='zeros', align_corners=True, dtype=np.float32, recip=False):
    """Pixel-space sampling coordinates exactly as BasicSR + ATen derive them.

    Each line is one separately-rounded elementwise op of the reference
    (no fused multiply-add across them).  ``recip``: ``tensor / python_scalar`` the way ATen's CUDA kernel evaluates it
    (aten/src/ATen/native/cuda/BinaryDivTrueKernel.cu: a CPU-scalar divisor b becomes a multiplication by ``1.0f / b``);
    False = ATen's CPU kernel (a true division).  The reference runs on CUDA (output_GPEMSR.py:44).
    """
    f = dtype
    flow = np.asarray(flow, dtype=f)
    gx = np.arange(w, dtype=f)[None, None, :]
    gy = np.arange(h, dtype=f)[None, :, None]
    vx = gx + flow[..., 0]
    vy = gy + flow[..., 1]
    if recip:
        nx = (f(2.0) * vx) * (f(1.0) / f(max(w - 1, 1))) - f(1.0)
        ny = (f(2.0) * vy) * (f(1.0) / f(max(h - 1, 1))) - f(1.0)
    else:
        nx = (f(2.0) * vx) / f(max(w - 1, 1)) - f(1.0)
        ny = (f(2.0) * vy) / f(max(h - 1, 1)) - f(1.0)

    def unnormalize(c, size):
        if align_corners:                       # ((c + 1) / 2) * (size - 1)
            return ((c + f(1.0)) / f(2.0)) * f(size - 1)
        return ((c + f(1.0)) * f(size) - f(1.0)) / f(2.0)

    ix = unnormalize(nx, w)
    iy = unnormalize(ny, h)
    if padding_mode == 'border':                # clip_coordinates: min(size-1, max(c, 0))
        ix = np.minimum(f(w - 1), np.maximum(ix, f(0.0)))
        iy = np.minimum(f(h - 1), np.maximum(iy, f(0.0)))
    elif padding_mode != 'zeros':
        raise NotImplementedError(padding_mode)
    return ix.astype(f), iy.astype(f)


def flow_warp_numpy(x, flow, interp_mode='bilinear', padding_mode='zeros', align_corners=True,
                    dtype=np.float32, recip=False):
    """x: [n, c, h, w]; flow: [n, h, w, 2] -> [n, c, h, w]."""
    if interp_mode != 'bilinear':
        raise NotImplementedError(interp_mode)
    f = dtype
    x = np.asarray(x, dtype=f)
    n, c, h, w = x.shape
    assert flow.shape == (n, h, w, 2)
    ix, iy = source_coords(flow, h, w, padding_mode, align_corners, dtype, recip)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + f(1.0)
    y1 = y0 + f(1.0)
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    out = np.zeros_like(x)
    bi = np.arange(n)[:, None, None]
    for (xx, yy, ww) in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        xi = xx.astype(np.int64)
        yi = yy.astype(np.int64)
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        xi = np.clip(xi, 0, w - 1)
        yi = np.clip(yi, 0, h - 1)
        vals = x[bi, :, yi, xi]                 # [n, h, w, c]
        vals = np.where(ok[..., None], vals, f(0.0))
        out += np.transpose(vals * ww[..., None].astype(f), (0, 3, 1, 2))
    return out
