"""Host-side plumbing for the implicit-GEMM kernels: geometry, activation buffers, packed weights, launches.

Everything here is bookkeeping around the C ABI (``gpemsr_igemm`` and friends in include/gpemsr_b200.h):
PyTorch only provides device memory and the stream.  The activation format is described in the header
("padded K8-blocked"); ``Geom`` mirrors ``gpemsr_geom_t`` and ``IgemmDesc`` mirrors ``gpemsr_igemm_desc_t``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_EXP = 0, 1, 2, 3


class GeomC(C.Structure):
    _fields_ = [('n', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('padded', C.c_int32),
                ('r_img', C.c_int64), ('m0', C.c_int64), ('rows_alloc', C.c_int64)]


class IgemmDesc(C.Structure):
    _fields_ = [('a_hi', C.c_void_p), ('a_lo', C.c_void_p), ('b_hi', C.c_void_p), ('b_lo', C.c_void_p),
                ('a_geom', GeomC),
                ('k_pad', C.c_int32), ('taps', C.c_int32), ('tap_dy', C.c_int32 * 49), ('tap_dx', C.c_int32 * 49),
                ('b_rows', C.c_int32), ('b_packed', C.c_int32), ('n_cols', C.c_int32), ('split', C.c_int32),
                ('scale', C.c_float), ('bias', C.c_void_p), ('bias_per_row', C.c_int32), ('act', C.c_int32),
                ('slope', C.c_float), ('residual', C.c_void_p),
                ('o_geom', GeomC),
                ('up', C.c_int32), ('py', C.c_int32), ('px', C.c_int32), ('pixel_shuffle', C.c_int32), ('phase_cols', C.c_int32),
                ('c_off', C.c_int32),
                ('out_f32', C.c_void_p), ('out_hi', C.c_void_p), ('out_lo', C.c_void_p),
                ('out_nchw', C.c_void_p), ('nchw_c', C.c_int32),
                ('out_rowmajor', C.c_void_p), ('ld', C.c_int64),
                ('gn_sums', C.c_void_p), ('gn_cpg', C.c_int32),
                ('patch_other', C.c_void_p), ('patch_sums', C.c_void_p), ('patch_size', C.c_int32),
                ('err_flag', C.c_void_p),
                ('row_max_out', C.c_void_p), ('row_max', C.c_void_p), ('row_sum', C.c_void_p), ('row_div', C.c_void_p),
                ('patch_other_bf16', C.c_int32)]


def _round_up(v, m):
    return (v + m - 1) // m * m


def cached(cache, name, params, build):
    """Packed / derived weights are cached per plan, keyed on the live parameters they were built from: a parameter that was
    reloaded (``load_state_dict`` copies in place: ``_version`` changes), moved (``.to()``: new storage) or updated in place is
    re-packed on the next call, like the reference nn.Module, which always reads the live parameters."""
    key = tuple((p.data_ptr(), p._version) for p in params)
    ent = cache.get(name)
    if ent is None or ent[0] != key:
        ent = cache[name] = (key, build())
    return ent[1]


class Precision:
    """Which arithmetic every GEMM launch runs: 3 = the fp32-faithful split (hi*hi + lo*hi + hi*lo on the bf16 tensor pipe),
    1 = one bf16 pass.  spec: 'fp32' (3 everywhere) | 'bf16' (1 everywhere) | {layer-name prefix: 1 | 3, 'default': 3}
    (longest matching prefix wins; layer names are the plan keys of the host mirrors, e.g. 'tail.up', 'vgg', 'tda.')."""

    def __init__(self, spec='fp32'):
        if isinstance(spec, Precision):
            self.table, self.default = dict(spec.table), spec.default
        elif isinstance(spec, dict):
            self.table = {k: int(v) for k, v in spec.items() if k != 'default'}
            self.default = int(spec.get('default', 3))
        elif spec in ('fp32', 'bf16'):
            self.table, self.default = {}, 3 if spec == 'fp32' else 1
        else:
            raise ValueError(f'precision must be "fp32", "bf16" or a {{prefix: split}} dict, got {spec!r}')
        if any(v not in (1, 3) for v in list(self.table.values()) + [self.default]):
            raise ValueError('precision: splits are 1 (one bf16 pass) or 3 (fp32-faithful)')

    def split(self, name):
        best, val = -1, self.default
        for k, v in self.table.items():
            if name.startswith(k) and len(k) > best:
                best, val = len(k), v
        return val

    def planes(self):
        """3 when any launch may need the lo planes of the activations, else 1."""
        return 3 if self.default == 3 or 3 in self.table.values() else 1

    def sub(self, prefix):
        """The precision of a sub-module whose layer names live under `prefix` + '.' in this table."""
        pre = prefix + '.'
        t = {k[len(pre):]: v for k, v in self.table.items() if k.startswith(pre)}
        t['default'] = self.split(prefix)
        return Precision(t)


# ---- the device-side pipeline error flag (every mbarrier wait is bounded; a time-out sets it and the kernel drains)
_ERR = {}


class _ErrState:
    def __init__(self, device):
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.host = None                   # pinned mirror, allocated by the first read-back
        self.event = None


def _err_state(device):
    key = (device.type, device.index if device.index is not None or device.type != 'cuda' else torch.cuda.current_device())
    st = _ERR.get(key)
    if st is None:
        st = _ERR[key] = _ErrState(device)
    return st


def err_flag(device):
    """ONE flag per device, shared by every plan of every module, so a forward needs one 4-byte read-back to know."""
    return _err_state(device).flag


def _raise_pipeline_error(st, code):
    st.flag.zero_()                       # otherwise every later launch would abort after its first 1024 probes
    if st.host is not None:
        st.host.zero_()
    st.event = None
    raise _lib.GpemsrError(-4, f'a GEMM pipeline timed out at wait site {code} (preemption / time-slicing longer than the bounded '
                               'mbarrier waits, or a pipeline bug): the outputs of that call are INVALID; the flag has been reset')


_nested = [0]


class nested:
    """Inside a composite forward (GPEMSR calling its sub-modules' public methods) only the outermost call posts the read-back."""

    def __enter__(self):
        _nested[0] += 1

    def __exit__(self, *a):
        _nested[0] -= 1


def post_error_check(device):
    """Queue an asynchronous read-back of the flag behind the kernels launched so far (no host sync).  Called at the end of every
    public forward; skipped while a CUDA graph is being captured."""
    if _nested[0] or torch.cuda.is_current_stream_capturing():
        return
    st = _err_state(device)
    if st.host is None:
        st.host = torch.zeros(1, dtype=torch.int32).pin_memory()
    st.host.copy_(st.flag, non_blocking=True)
    st.event = torch.cuda.Event()
    st.event.record()


def poll_error(device, wait=False):
    """Raise if a finished read-back shows a time-out.  wait=True synchronises on the read-back first (used where the caller
    synchronises anyway: the volume driver's D2H, ``check()``)."""
    st = _err_state(device)
    if st.event is None or torch.cuda.is_current_stream_capturing():
        return
    if wait:
        st.event.synchronize()
    elif not st.event.query():
        return
    st.event = None
    code = int(st.host[0])
    if code:
        _raise_pipeline_error(st, code)


class Geom:
    """Row geometry of a flattened image batch (mirrors gpemsr_geom_t)."""

    def __init__(self, n, h, w, padded=True, m0=None, rows_alloc=None, r_img=None):
        """padded: zero-ring width (True = 1: 3x3 taps; 3: 7x7 taps; False / 0: compact rows)."""
        self.n, self.h, self.w, self.padded = int(n), int(h), int(w), int(padded)
        P = self.padded
        per = (h + 2 * P) * (w + 2 * P)
        self.r_img = _round_up(per, 128) if r_img is None else int(r_img)
        margin = _round_up(P * (w + 2 * P) + P, 128) if P else 0
        self.m0 = margin if m0 is None else int(m0)
        # tail: room for the +1-row tap shifts (padded) / for a last column tile read as B operand (compact)
        tail = margin if padded else 256
        self.rows_alloc = (self.m0 + self.n * self.r_img + tail) if rows_alloc is None else int(rows_alloc)

    @property
    def c(self):
        return GeomC(self.n, self.h, self.w, int(self.padded), self.r_img, self.m0, self.rows_alloc)

    def sample(self, i):
        """The same storage seen as the single image i (used to run per-sample GEMMs on a batch buffer)."""
        return Geom(1, self.h, self.w, self.padded, m0=self.m0 + i * self.r_img, rows_alloc=self.rows_alloc, r_img=self.r_img)

    def key(self):
        return (self.n, self.h, self.w, self.padded)


class Act:
    """An activation tensor in the internal format: optional fp32 master + optional (hi, lo) bf16 operand planes."""

    def __init__(self, geom, c, device, f32=False, planes=True, split=3):
        self.geom, self.c = geom, int(c)
        self.c_pad = _round_up(self.c, 64)
        shape = (self.c_pad // 8, geom.rows_alloc, 8)
        self.f32 = torch.zeros(shape, dtype=torch.float32, device=device) if f32 else None
        self.hi = torch.zeros(shape, dtype=torch.bfloat16, device=device) if planes else None
        self.lo = torch.zeros(shape, dtype=torch.bfloat16, device=device) if (planes and split == 3) else None


class Weights:
    """B-operand planes [taps][k_pad/8][b_rows][8] packed once from a reference-layout parameter."""

    def __init__(self, w, kind, taps=None, split=3, block_rows=256, min_rows=0, pixel_shuffle=False, plain=False, flop_scale=1.0):
        """kind: 'conv' [co, ci, kh, kw] | 'convT' [ci, co, kh, kw] | 'linear' [n, k].
        taps: list of (src_index, dy, dx); default = all kh*kw taps of a 'same' convolution.
        flop_scale: algorithmic FLOPs of the reference layer / FLOPs of the launched tap grid (9/16 for the merged
        ConvTranspose phases and the space-to-depth stride-2 convs); bookkeeping for bench.py only."""
        self.flop_scale = flop_scale
        self.split = split
        if not w.is_cuda:
            raise _lib.GpemsrError(-3, f'weights live on {w.device}: move the module to the GPU first (there is no CPU fallback)')
        w = w.detach().contiguous().float()
        dev = w.device
        if kind == 'conv':
            co, ci, kh, kw = w.shape
            n, k, n_stride, k_stride = co, ci, ci * kh * kw, kh * kw
            if taps is None:
                taps = [(ky * kw + kx, ky - kh // 2, kx - kw // 2) for ky in range(kh) for kx in range(kw)]
            elif taps == 'offsets01':          # kernel index == input offset (merged ConvTranspose phases)
                taps = [(ky * kw + kx, ky, kx) for ky in range(kh) for kx in range(kw)]
        elif kind == 'convT':
            ci, co, kh, kw = w.shape
            n, k, n_stride, k_stride = co, ci, kh * kw, co * kh * kw
            assert taps is not None
        elif kind == 'linear':
            n, k = w.shape
            n_stride, k_stride = k, 1
            taps = [(0, 0, 0)]
        else:
            raise ValueError(kind)
        self.n, self.k = n, k
        self.taps = [(dy, dx) for _, dy, dx in taps]
        src = torch.tensor([t[0] for t in taps], dtype=torch.int32, device=dev)
        L = _lib.lib()
        # 3x3 convs with more than 64 input channels and <= 64 output channels (reference fusion, POD offsets): all nine taps'
        # weights do not fit shared memory next to the A stages, so the tap-fused kernel cannot run them and the streaming kernel
        # re-reads A nine times; the dy-fused kernel reads it three times: measured ~2x (reffusionconv1: 1.96 -> 0.99 ms).
        dy3 = len(taps) == 9 and k > 64 and os.environ.get('GPEMSR_DYFUSE3X3', '1') == '1'
        if kind == 'conv' and (len(taps) > 9 or dy3) and n <= 64 and not plain and self._full_grid():
            # large tap grids on narrow layers (SpyNet's 7x7): weights streamed per (16-wide k slab, tap row) -- the layout
            # [slab][dy][dx][cell][plane][block_n][8] makes every pipeline stage's weights one contiguous copy
            bn = 16 if n <= 16 else 32 if n <= 32 else 64
            k16 = _round_up(k, 16)
            n_dx = sum(1 for dy, _ in self.taps if dy == self.taps[0][0])
            n_dy = len(taps) // n_dx
            planes = []
            for _ in range(2 if split == 3 else 1):
                planes.append(torch.empty(len(taps), k16 // 8, bn, 8, dtype=torch.bfloat16, device=dev))
            _lib.check(L.gpemsr_pack_weights(_lib.ptr(w), n, k, n_stride, k_stride, len(taps), _lib.ptr(src), bn, k16,
                                             _lib.ptr(planes[0]), _lib.ptr(planes[1]) if split == 3 else None, _lib.stream_ptr()))
            st = torch.stack([p.view(n_dy, n_dx, k16 // 16, 2, bn, 8) for p in planes])      # [plane, dy, dx, slab, cell, bn, 8]
            # [slab, dy, dx, cell, plane, bn, 8]: the hi and lo rows of a (tap, k-cell) side by side = ONE N = 2 * bn B operand
            self.hi = st.permute(3, 1, 2, 4, 0, 5, 6).contiguous()
            self.lo = None
            self.packed, self.b_rows, self.k_pad = 2, bn, k16
            self.kernel = f'gemm_dyfuse_kernel<{bn}>'
            self._keep = (w, src)
            return
        self.k_pad = _round_up(k, 32 if split == 3 else 64)      # one k-chunk of the streaming kernel (narrow inputs: less padding)
        self.b_rows = _round_up(n, 16) if n <= 16 else _round_up(n, 64) if n <= 64 else _round_up(n, 128) if n <= 128 \
            else _round_up(n, block_rows)
        self.b_rows = max(self.b_rows, min_rows)
        # ask the library which kernel variant this shape runs: the streaming kernel wants the tiled single-copy layout
        d = IgemmDesc()
        d.n_cols, d.k_pad, d.taps, d.split, d.pixel_shuffle = n, self.k_pad, len(taps), split, int(pixel_shuffle)
        for i, (dy, dx) in enumerate(self.taps):
            d.tap_dy[i], d.tap_dx[i] = dy, dx
        bn, fused = C.c_int32(), C.c_int32()
        _lib.check(L.gpemsr_igemm_plan(C.byref(d), C.byref(bn), C.byref(fused)))
        self.packed = int((not fused.value) and min_rows == 0 and not plain)
        self.kernel = f'gemm_tapfuse_kernel<{bn.value}>' if fused.value else f'gemm_kernel<{bn.value}>'       # what gpemsr_igemm() will launch
        if self.packed:
            nbytes = L.gpemsr_pack_weights_tiled_bytes(n, self.k_pad, len(taps), bn.value, split)
            self.hi = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=dev)
            self.lo = None
            _lib.check(L.gpemsr_pack_weights_tiled(_lib.ptr(w), n, k, n_stride, k_stride, len(taps), _lib.ptr(src), bn.value,
                                                   self.k_pad, split, _lib.ptr(self.hi), _lib.stream_ptr()))
        else:
            shape = (len(taps), self.k_pad // 8, self.b_rows, 8)
            self.hi = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            self.lo = torch.empty(shape, dtype=torch.bfloat16, device=dev) if split == 3 else None
            _lib.check(L.gpemsr_pack_weights(_lib.ptr(w), n, k, n_stride, k_stride, len(taps), _lib.ptr(src),
                                             self.b_rows, self.k_pad, _lib.ptr(self.hi), _lib.ptr(self.lo), _lib.stream_ptr()))
        self._keep = (w, src)


def _is_full_grid(taps):
    """taps: list of (dy, dx) -- a full rectangular grid in dy-major order with unit steps?"""
    n_dx = sum(1 for dy, _ in taps if dy == taps[0][0])
    if len(taps) % n_dx:
        return False
    return all(t == (taps[0][0] + i // n_dx, taps[0][1] + i % n_dx) for i, t in enumerate(taps))


Weights._full_grid = lambda self: _is_full_grid(self.taps)


def convT_phase_taps(py, px):
    """ConvTranspose2d(k3, s2, p1, op1): output (2a+py, 2b+px) = sum over the listed (ky, kx) of
    in(a + dy, b + dx) * W[:, :, ky, kx]  with  oy = 2*iy - 1 + ky  (model/blocks.py:35)."""
    ks = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}          # parity -> [(k index, input offset)]
    return [(ky * 3 + kx, dy, dx) for ky, dy in ks[py] for kx, dx in ks[px]]


def convT_merged_weight(w):
    """ConvTranspose2d(k3, s2, p1, op1) weight [ci, co, 3, 3] -> dense [4*co, ci, 2, 2]: the four output-parity phases side
    by side along the output channels, over the union of their taps (input offsets (dy, dx) in {0,1}^2); taps a phase does
    not use are zero.  Used where the layer is bound by memory traffic, not MMAs: one launch reads the input once and writes
    whole output cells."""
    ci, co = w.shape[0], w.shape[1]
    m = torch.zeros(4 * co, ci, 2, 2, dtype=torch.float32, device=w.device)
    for py in (0, 1):
        for px in (0, 1):
            p = py * 2 + px
            for src, dy, dx in convT_phase_taps(py, px):
                m[p * co:(p + 1) * co, :, dy, dx] = w[:, :, src // 3, src % 3].t()
    return m


def compose_upblock_conv(wt, bu, wo, bo):
    """Compose ConvTranspose2d(k3, s2, p1, op1) [weight wt: ci x c x 3 x 3, bias bu] with a following zero-padded 3x3
    Conv2d [weight wo: co x c x 3 x 3, bias bo] (no non-linearity in between: model/decoder.py:31,33) into a 4-phase,
    3x3-tap linear map on the INPUT grid:

        out[co, 2a+py, 2b+px] = bias[cls][co] + sum_ci sum_{dy,dx in -1..1} x[ci, a+dy, b+dx] * wc[cls][py*2+px][co][ci][dy+1][dx+1]

    `cls = ry*3 + rx` is the position class of the output pixel (0 first row/col, 1 interior, 2 last row/col): the conv's
    zero padding removes the taps that fall outside the up-sampled image, so ring pixels have their own weights.
    Derivation: the conv reads U at o + t (t in -1..1); U[P] = sum_k x[(P + 1 - k) / 2] * wt[k] over the k of matching
    parity; with P = 2a + p + t this is x[a + d] for d = (p + t + 1 - k) / 2.  Done once per module in float64."""
    wt64, wo64 = wt.detach().double().cpu(), wo.detach().double().cpu()
    bu64, bo64 = bu.detach().double().cpu(), bo.detach().double().cpu()
    ci, co = wt64.shape[0], wo64.shape[0]
    wc = torch.zeros(9, 4, co, ci, 3, 3, dtype=torch.float64)
    bias = torch.zeros(9, co, dtype=torch.float64)
    valid = {0: (0, 1), 1: (-1, 0, 1), 2: (-1, 0)}           # taps t that stay inside, per position class
    for ry in range(3):
        for rx in range(3):
            cls = ry * 3 + rx
            bias[cls] = bo64
            for ty in valid[ry]:
                for tx in valid[rx]:
                    wo_t = wo64[:, :, ty + 1, tx + 1]                      # co x c
                    bias[cls] += wo_t @ bu64
                    for py in range(2):
                        for px in range(2):
                            for ky in range(3):
                                ny = py + ty + 1 - ky
                                if ny % 2:
                                    continue
                                for kx in range(3):
                                    nx = px + tx + 1 - kx
                                    if nx % 2:
                                        continue
                                    wc[cls, py * 2 + px, :, :, ny // 2 + 1, nx // 2 + 1] += wo_t @ wt64[:, :, ky, kx].t()
    return wc.float(), bias.float()


def border_phase_conv(x, wc, bias, cout, out):
    g = x.geom.c
    _lib.check(_lib.lib().gpemsr_border_phase_conv(_lib.ptr(x.f32), x.c, C.byref(g), _lib.ptr(wc), _lib.ptr(bias), cout,
                                                   _lib.ptr(out), _lib.stream_ptr()))


def _p(t):
    if t is None or isinstance(t, int):
        return t
    if not t.is_cuda:
        raise _lib.GpemsrError(-3, f'expected a CUDA tensor, got one on {t.device}: there is no CPU fallback')
    return t.data_ptr()


def igemm(a, w, err, *, n_cols=None, split=3, scale=1.0, bias=None, bias_per_row=False, act=ACT_NONE, slope=0.0,
          residual=None, out=None, a_geom=None, o_geom=None, up=1, py=0, px=0, pixel_shuffle=False, phase_cols=0, c_off=0,
          out_f32=True, out_planes=True, out_nchw=None, nchw_c=0, out_rowmajor=None, ld=0,
          b_hi=None, b_lo=None, b_rows=None, k_pad=None, taps=None, gn_sums=None, gn_cpg=0,
          patch_other=None, patch_sums=None, patch_size=0, row_max_out=None, row_max=None, row_sum=None, row_div=None, patch_other_bf16=False):
    """One fused implicit-GEMM launch.  `a`: Act (A operand); `w`: Weights or None when b_* are given explicitly;
    `out`: Act receiving fp32 master / planes (whichever it owns and the flags allow)."""
    d = IgemmDesc()
    d.a_hi, d.a_lo = _p(a.hi), _p(a.lo)
    d.a_geom = (a_geom or a.geom).c
    if w is not None:
        d.b_hi, d.b_lo, d.b_rows, d.k_pad = _p(w.hi), _p(w.lo), w.b_rows, w.k_pad
        d.b_packed = int(w.packed)
        tp = w.taps
        d.n_cols = w.n if n_cols is None else n_cols
    else:
        d.b_hi, d.b_lo, d.b_rows, d.k_pad = b_hi, b_lo, b_rows, k_pad
        tp = taps or [(0, 0)]
        d.n_cols = n_cols
    d.taps = len(tp)
    for i, (dy, dx) in enumerate(tp):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.split = split
    d.scale = scale
    d.bias, d.bias_per_row, d.act, d.slope = _p(bias), int(bias_per_row), act, slope
    d.residual = _p(residual)
    og = o_geom or (out.geom if out is not None else (a_geom or a.geom))
    d.o_geom = og.c
    d.up, d.py, d.px, d.pixel_shuffle, d.phase_cols, d.c_off = up, py, px, int(pixel_shuffle), phase_cols, c_off
    if out is not None:
        d.out_f32 = _p(out.f32) if out_f32 else None
        d.out_hi = _p(out.hi) if out_planes else None
        d.out_lo = _p(out.lo) if out_planes else None
    d.out_nchw, d.nchw_c = _p(out_nchw), nchw_c
    d.out_rowmajor, d.ld = _p(out_rowmajor), ld
    d.gn_sums, d.gn_cpg = _p(gn_sums), gn_cpg
    d.patch_other, d.patch_sums, d.patch_size, d.patch_other_bf16 = _p(patch_other), _p(patch_sums), patch_size, int(patch_other_bf16)
    d.err_flag = _p(err)
    d.row_max_out, d.row_max, d.row_sum, d.row_div = _p(row_max_out), _p(row_max), _p(row_sum), _p(row_div)
    _lib.check(_lib.lib().gpemsr_igemm(C.byref(d), _lib.stream_ptr()))


def pack_nchw(x, act, c_off=0):
    x = x.contiguous()
    g = act.geom.c
    _lib.check(_lib.lib().gpemsr_act_pack_nchw(_lib.ptr(x), x.shape[1], C.byref(g), c_off, _lib.ptr(act.f32),
                                               _lib.ptr(act.hi), _lib.ptr(act.lo), _lib.stream_ptr()))


def unpack_nchw(act, c=None, c_off=0):
    g = act.geom
    c = act.c if c is None else c
    out = torch.empty(g.n, c, g.h, g.w, dtype=torch.float32, device=act.f32.device)
    gc = g.c
    _lib.check(_lib.lib().gpemsr_act_unpack_nchw(_lib.ptr(act.f32), c, C.byref(gc), c_off, _lib.ptr(out), _lib.stream_ptr()))
    return out


def space_to_depth(x, out):
    """x: Act with c channels (c % 8 == 0) at h x w -> out: Act with 4*c channels at ceil(h/2) x ceil(w/2) (planes only)."""
    gi, go = x.geom.c, out.geom.c
    _lib.check(_lib.lib().gpemsr_space_to_depth(_lib.ptr(x.hi), _lib.ptr(x.lo), C.byref(gi), x.c, _lib.ptr(out.hi),
                                                _lib.ptr(out.lo), C.byref(go), _lib.stream_ptr()))


def down_conv_weight(w):
    """Conv2d(k3, s2, p1) weight [co, ci, 3, 3] (model/blocks.py:44) -> [co, 4*ci, 2, 2] acting on the space-to-depth input:
    input row 2y + dy (dy = ky - 1) is phase p = dy & 1 of s2d row y + (dy - p) / 2, i.e. s2d offset sy in {-1, 0}; the
    (p = 0, sy = -1) combinations do not occur and stay zero.  Returns (weight, taps) for ``Weights(kind='conv')``."""
    co, ci = w.shape[0], w.shape[1]
    m = torch.zeros(co, 4 * ci, 2, 2, dtype=torch.float32, device=w.device)
    for ky in range(3):
        py = (ky - 1) & 1
        sy = (ky - 1 - py) // 2
        for kx in range(3):
            px = (kx - 1) & 1
            sx = (kx - 1 - px) // 2
            ph = py * 2 + px
            m[:, ph * ci:(ph + 1) * ci, sy + 1, sx + 1] = w[:, :, ky, kx]
    taps = [(a * 2 + b, a - 1, b - 1) for a in range(2) for b in range(2)]
    return m, taps


class GroupNormScratch:
    def __init__(self, n, c, device):
        self.sums = torch.zeros(n, c, 2, dtype=torch.float64, device=device)
        self.ss = torch.empty(n, c, 2, dtype=torch.float32, device=device)


def group_norm_act(x, gamma, beta, scratch, out, act=ACT_NONE, slope=0.0, residual=None, groups=32, eps=1e-6,
                   out_f32=True, out_planes=True, out_nchw=None, fused_stats=False):
    """GroupNorm(32, eps=1e-6) of the fp32 master of `x` (model/blocks.py:5-6), then act (+ residual) into `out`.
    fused_stats: the producing GEMM already accumulated per-group sums into scratch.sums (igemm gn_sums=...)."""
    L = _lib.lib()
    g = x.geom.c
    og = out.geom.c
    st = _lib.stream_ptr()
    if not fused_stats:
        _lib.check(L.gpemsr_gn_stats(_lib.ptr(x.f32), x.c, C.byref(g), _lib.ptr(scratch.sums), st))
    _lib.check(L.gpemsr_gn_scale_shift(_lib.ptr(scratch.sums), int(fused_stats), _lib.ptr(gamma), _lib.ptr(beta), x.geom.n, x.c,
                                       groups, float(x.geom.h * x.geom.w), eps, _lib.ptr(scratch.ss), st))
    _lib.check(L.gpemsr_affine_act(_lib.ptr(x.f32), x.c, C.byref(g), _lib.ptr(scratch.ss), act, slope, _lib.ptr(residual),
                                   C.byref(og), _lib.ptr(out.f32) if out_f32 else None,
                                   _lib.ptr(out.hi) if out_planes else None, _lib.ptr(out.lo) if out_planes else None,
                                   _lib.ptr(out_nchw), st))


def softmax_cells_blocked(s_cells, t, rows_alloc, t_pad, scratch, p_hi, p_lo):
    _lib.check(_lib.lib().gpemsr_softmax_cells_blocked(_lib.ptr(s_cells), t, rows_alloc, t_pad, _lib.ptr(scratch), _lib.ptr(p_hi),
                                                       _lib.ptr(p_lo), _lib.stream_ptr()))


def taps_as_columns(w):
    """Conv2d weight [n_out, ci, 3, 3] -> the 1x1 weight [9 * n_out, ci, 1, 1] whose column tap * n_out + o is tap (ky, kx) of
    output o (tap = ky * 3 + kx): the GEMM half of ``conv3x3_few_outputs``."""
    n_out, ci = w.shape[0], w.shape[1]
    return w.detach().float().permute(2, 3, 0, 1).reshape(9 * n_out, ci, 1, 1).contiguous()


class TapCells:
    """fp32 cells [ceil(9 * n_out / 8)][rows_alloc][8] receiving the per-tap partial products (zero ring: never written)."""

    def __init__(self, geom, n_out, device):
        self.geom, self.c = geom, 9 * n_out
        self.f32 = torch.zeros((self.c + 7) // 8, geom.rows_alloc, 8, dtype=torch.float32, device=device)
        self.hi = self.lo = None


def conv3x3_few_outputs(x, wt, taps, err, split, bias, out_nchw, n_out, up=1, co=0, act=ACT_NONE, slope=0.0, base=None,
                        base_scale=1):
    """3x3 convolution with <= 4 output columns: ONE 1x1 GEMM producing the nine per-tap partial products as fp32 cells of
    `taps` (a TapCells), then the nine-point shifted sum (+ bias, activation, + bilinear base image) on CUDA cores.
    wt = Weights(taps_as_columns(w), 'conv')."""
    igemm(x, wt, err, split=split, out=taps, out_planes=False)
    g = x.geom.c
    bh, bw = (base.shape[2], base.shape[3]) if base is not None else (0, 0)
    _lib.check(_lib.lib().gpemsr_tap_gather_sum(_lib.ptr(taps.f32), C.byref(g), n_out, up, co, _lib.ptr(bias), act, slope,
                                                _lib.ptr(base), bh, bw, base_scale, _lib.ptr(out_nchw), _lib.stream_ptr()))


def add_bilinear_base(x_center, scale, out):
    n, _, h, w = x_center.shape
    _lib.check(_lib.lib().gpemsr_add_bilinear_base(_lib.ptr(x_center.contiguous()), n, h, w, scale, _lib.ptr(out),
                                                   _lib.stream_ptr()))


def check_pipeline(err):
    """Synchronises and raises if any GEMM pipeline timed out."""
    code = int(err.item())
    if code:
        _raise_pipeline_error(_err_state(err.device), code)
