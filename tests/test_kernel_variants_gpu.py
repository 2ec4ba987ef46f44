"""GPU: the round-2 kernel variants against the forms they replace, on the same inputs.

The CTA-pair streaming GEMM (tcgen05.mma.cta_group::2, tensor-map operands) accumulates the same products in the same order as the
cta_group::1 + multicast kernel, and the tensor-map / multi-slab / plain-epilogue forms of the tap-fused and dy-fused kernels only
change how operands travel and which epilogue branches exist -- so every result must be BIT-identical to a run of the same
library with GPEMSR_PAIR=0 GPEMSR_TMA=0 GPEMSR_LEAN=0 (the switches are read once per process: the reference run is a child
process; GPEMSR_LEAN=0 sends every launch through the generic epilogue instead of its specialised instantiation).
Also checks that the tensor-map path really ran here (maps built, none refused)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name, images, Cin, Cout, size, kernel size, options
CASES = [
    ('pair256_split3', 2, 128, 256, 40, 3, {}),                       # packed weights, N = 256, K = 128, 9 taps
    ('pair256_pixel_shuffle', 1, 64, 256, 48, 3, dict(ps=True)),      # the up-convs: PixelShuffle epilogue
    ('pair128_split3', 2, 128, 128, 40, 3, dict(residual=True)),      # N = 128, plain epilogue + residual
    ('pair64_stream', 2, 1024, 64, 24, 1, {}),                        # N = 64 streaming (1x1, K = 1024: too large to stay resident)
    ('pair256_single_pass', 1, 128, 256, 40, 1, dict(split=1)),       # BLOCK_K = 64, one bf16 pass
    ('tapfuse64_split3', 3, 64, 64, 50, 3, dict(act=2, slope=0.1)),   # tensor-map A tiles, plain epilogue
    ('tapfuse64_f32_residual', 1, 64, 64, 70, 3, dict(f32=True, residual=True)),
    ('tapfuse64_single_pass', 2, 64, 64, 60, 3, dict(split=1, act=1)),     # four k-slabs per stage
    ('dyfuse_7x7', 2, 32, 64, 40, 7, dict(split=1, act=1)),           # SpyNet layer
    ('dyfuse_k128', 1, 128, 64, 40, 3, {}),                           # K > 64, N = 64: (k-slab, tap row) stages
    ('odd_tiles', 1, 128, 256, 19, 3, {}),                            # odd number of row tiles: the pair's surplus tile is discarded
    ('convT_merged_phases', 2, 128, 64, 24, 3, dict(convT=True)),     # ConvTranspose2d as ONE 4-phase GEMM: phase-scatter epilogue
    ('tap_gemm_36_columns', 1, 64, 4, 40, 3, dict(taps_gemm=True)),   # few-output conv: 1x1 GEMM with 9 * 4 columns (partial tile)
]


def run_cases():
    from gpemsr_b200 import igemm as G
    out = {}
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    for i, (name, n, ci, co, s, ks, opt) in enumerate(CASES):
        gen = torch.Generator(device='cuda').manual_seed(1000 + i)
        g = G.Geom(n, s, s, ks // 2 if ks > 1 else True)
        x = G.Act(g, ci, 'cuda', f32=False)
        v = torch.randn(n, ci, s, s, device='cuda', generator=gen)
        G.pack_nchw(v, x)
        w = torch.randn(co, ci, ks, ks, device='cuda', generator=gen) * 0.05
        b = torch.randn(co, device='cuda', generator=gen)
        wt = G.Weights(w, 'conv', split=opt.get('split', 3), pixel_shuffle=bool(opt.get('ps')))
        kw = dict(split=opt.get('split', 3), bias=b, act=opt.get('act', G.ACT_NONE), slope=opt.get('slope', 0.0))
        if opt.get('convT'):
            wT = torch.randn(ci, co, 3, 3, device='cuda', generator=gen) * 0.05
            wt = G.Weights(G.convT_merged_weight(wT), 'conv', taps='offsets01')
            y = G.Act(G.Geom(n, 2 * s, 2 * s, True), co, 'cuda', f32=True)
            G.igemm(x, wt, err, bias=b.repeat(4).contiguous(), out=y, up=2, phase_cols=co, out_f32=True)
            out[name] = (y.hi.clone().cpu(), y.lo.clone().cpu(), y.f32.clone().cpu())
            continue
        if opt.get('taps_gemm'):
            wt = G.Weights(G.taps_as_columns(w), 'conv')
            cells = G.TapCells(g, co, 'cuda')
            G.igemm(x, wt, err, out=cells, out_planes=False)
            out[name] = (None, None, cells.f32.clone().cpu())
            continue
        if opt.get('ps'):
            y = G.Act(G.Geom(n, 2 * s, 2 * s, True), co // 4, 'cuda', f32=False)
            G.igemm(x, wt, err, out=y, up=2, pixel_shuffle=True, out_f32=False, **kw)
        else:
            y = G.Act(G.Geom(n, s, s, True), co, 'cuda', f32=bool(opt.get('f32')))
            res = None
            if opt.get('residual'):
                r = G.Act(y.geom, co, 'cuda', f32=True)
                r.f32.normal_(generator=gen)
                res = r.f32
            G.igemm(x, wt, err, out=y, out_f32=bool(opt.get('f32')), residual=res, **kw)
        out[name] = (y.hi.clone().cpu(), y.lo.clone().cpu() if y.lo is not None else None,
                     y.f32.clone().cpu() if getattr(y, 'f32', None) is not None and opt.get('f32') else None)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return out


def test_pair_and_tensor_map_kernels_match_the_plain_kernels_bit_for_bit(cuda_dev, tmp_path):
    from gpemsr_b200 import _lib
    got = run_cases()
    built, rej = C.c_int64(0), C.c_int64(0)
    _lib.lib().gpemsr_tensor_map_stats(C.byref(built), C.byref(rej))
    assert built.value > 0 and rej.value == 0, (built.value, rej.value)          # the tensor-map path ran, the driver took every shape
    ref_file = str(tmp_path / 'ref.pt')
    env = dict(os.environ, GPEMSR_PAIR='0', GPEMSR_TMA='0', GPEMSR_LEAN='0', PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, 'tests'))
    code = ("import sys, torch; sys.path[:0] = [%r, %r]; import test_kernel_variants_gpu as T; "
            "torch.save(T.run_cases(), %r)" % (ROOT, os.path.join(ROOT, 'tests'), ref_file))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = torch.load(ref_file)
    assert set(ref) == set(got)
    for name in got:
        for a, b, what in zip(got[name], ref[name], ('hi', 'lo', 'f32')):
            assert (a is None) == (b is None), (name, what)
            if a is not None:
                assert torch.equal(a.view(torch.int16) if a.dtype == torch.bfloat16 else a, b.view(torch.int16) if b.dtype == torch.bfloat16 else b), \
                    (name, what, float((a.float() - b.float()).abs().max()))
