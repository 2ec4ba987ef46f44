"""Quick single-GPU microbenchmarks (CUDA events, L2-flushed) used while developing; bench.py is the contract."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import gpemsr_b200


def timeit(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def flow_warp_bench():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    out = []
    for c, s in ((64, 156), (64, 312), (64, 624), (64, 1250), (3, 128), (3, 640)):
        x = torch.randn(1, c, s, s, device='cuda')
        fl = 2.0 * torch.randn(1, s, s, 2, device='cuda')
        fl = torch.nn.functional.avg_pool2d(fl.permute(0, 3, 1, 2), 5, 1, 2).permute(0, 2, 3, 1).contiguous()
        med, best = timeit(lambda: gpemsr_b200.flow_warp(x, fl, 'bilinear', 'border'), flush=flush)
        byt = 8 * c * s * s + 8 * s * s
        from oracle.flow_warp import flow_warp_torch
        med_t, _ = timeit(lambda: flow_warp_torch(x, fl, 'bilinear', 'border'), flush=flush)
        white = 4.0 * torch.randn(1, s, s, 2, device='cuda')
        med_w, _ = timeit(lambda: gpemsr_b200.flow_warp(x, white, 'bilinear', 'border'), flush=flush)
        out.append(dict(op='flow_warp', c=c, s=s, ms=med, ms_best=best, gbs=byt / med / 1e6, frac=byt / med / 1e6 / 6550.1,
                        ms_white=med_w, gbs_white=byt / med_w / 1e6, torch_grid_sample_ms=med_t))
    return out


def vq_bench(sizes=tuple((1 << e, d) for d in (512, 256) for e in (20, 18, 16, 14, 12, 10))):       # BASELINE configs[2]: N = 1 Ki ... 1 Mi
    from oracle import ref_ops as R
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    out = []
    for n, d in sizes:
        z = torch.randn(1, d, n, 1, device='cuda')
        emb = torch.randn(1024, d, device='cuda')
        med, best = timeit(lambda: gpemsr_b200.vq_lookup(z, emb), iters=5, flush=flush)
        flops = 2.0 * n * 1024 * d
        byt = 8.0 * n * d + 8 * n + 4096 * d
        torch.backends.cuda.matmul.allow_tf32 = False
        med_t, _ = timeit(lambda: R.codebook_forward(z, emb), iters=3, warm=1, flush=flush)
        out.append(dict(op='vq_lookup', n=n, d=d, ms=med, ms_best=best, tflops=flops / med / 1e9, tc_frac=flops / med / 1e9 / 1649.2,
                        gbs=byt / med / 1e6, hbm_frac=byt / med / 1e6 / 6550.1, torch_ref_ms=med_t))
    return out


def conv_bench(which=None):
    """Single conv layers of the bench step, fp32-faithful split-3 path: ms and algorithmic TFLOP/s."""
    from gpemsr_b200 import igemm as G
    cases = [('rb512_80', 5, 512, 512, 80, 3, {}), ('rb256_160', 5, 256, 256, 160, 3, {}), ('rb128_320', 5, 128, 128, 320, 3, {}),
             ('rb64_640', 5, 64, 64, 640, 3, {}), ('rb64_80', 5, 64, 64, 80, 3, {}), ('rb64_160', 5, 64, 64, 160, 3, {}), ('hr64_1280', 1, 64, 64, 1280, 3, dict(act=G.ACT_LRELU, slope=0.1)),
             ('up256_640', 1, 64, 256, 640, 3, dict(ps=True)), ('out1_1280', 5, 64, 1, 1280, 3, dict(nchw=True)),
             ('q512_80', 5, 512, 512, 80, 1, {}),
             ('vgg64_1280', 2, 64, 64, 1280, 3, dict(split=1, act=G.ACT_RELU, planes_only=True)),      # VGG conv1_2, one bf16 pass
             ('hr64_1280p', 1, 64, 64, 1280, 3, dict(act=G.ACT_LRELU, slope=0.1, planes_only=True)),     # (hi, lo) planes out, no fp32 copy
             ('spy7x7_320', 5, 32, 64, 320, 7, dict(split=1, act=G.ACT_RELU, planes_only=True))]         # a SpyNet 7x7 layer (dy-fused kernel)
    out = []
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    for name, n, ci, co, s, ks, opt in cases:
        if which and name not in which:
            continue
        g = G.Geom(n, s, s, ks // 2 if ks > 1 else True)
        x = G.Act(g, ci, 'cuda', f32=False)
        x.hi.normal_(); x.lo.normal_(std=0.004)
        w = torch.randn(co, ci, ks, ks, device='cuda') * 0.05
        b = torch.randn(co, device='cuda')
        wt = G.Weights(w, 'conv')
        kw = dict(split=opt.get('split', 3), bias=b, act=opt.get('act', G.ACT_NONE), slope=opt.get('slope', 0.0))
        if opt.get('ps'):
            y = G.Act(G.Geom(n, 2 * s, 2 * s, True), co // 4, 'cuda', f32=False)
            fn = lambda: G.igemm(x, wt, err, out=y, up=2, pixel_shuffle=True, out_f32=False, **kw)
        elif opt.get('nchw'):
            img = torch.empty(n, co, s, s, device='cuda')
            fn = lambda: G.igemm(x, wt, err, out_nchw=img, nchw_c=co, **kw)
        elif opt.get('planes_only'):
            y = G.Act(g, co, 'cuda', f32=False)
            fn = lambda: G.igemm(x, wt, err, out=y, out_f32=False, **kw)
        else:
            y = G.Act(g, co, 'cuda', f32=True)
            fn = lambda: G.igemm(x, wt, err, out=y, **kw)
        med, best = timeit(fn, iters=5, warm=2)
        fl = 2.0 * n * s * s * co * ci * ks * ks
        out.append(dict(op='conv', name=name, n=n, cin=ci, cout=co, hw=s, k=ks, ms=med, tflops=fl / med / 1e9,
                        frac_of_split3_ceiling=fl / med / 1e9 / (1388.4 / 3)))
        del x, wt
    assert int(err.item()) == 0
    import ctypes
    from gpemsr_b200 import _lib
    built, rej = ctypes.c_int64(0), ctypes.c_int64(0)
    if hasattr(_lib.lib(), 'gpemsr_tensor_map_stats'):
        _lib.lib().gpemsr_tensor_map_stats(ctypes.byref(built), ctypes.byref(rej))
        print(json.dumps(dict(tensor_maps_built=built.value, tensor_maps_rejected=rej.value)), file=sys.stderr)
    return out


def small_conv_bench(reps=20):
    """Fixed cost of a tap-fused launch: the 64 -> 64 3x3 conv on SMALL images (2 .. 9 tiles per SM), `reps` launches captured in
    one CUDA graph so that the host's launch path is out of the measurement; us per launch."""
    from gpemsr_b200 import igemm as G
    out = []
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    for name, n, s, split in (('rb64_80', 5, 80, 3), ('rb64_160', 5, 160, 3), ('rb64_320', 5, 320, 3), ('vgg64_160', 5, 160, 1)):
        g = G.Geom(n, s, s, True)
        x = G.Act(g, 64, 'cuda', f32=False)
        x.hi.normal_(); x.lo.normal_(std=0.004)
        wt = G.Weights(torch.randn(64, 64, 3, 3, device='cuda') * 0.05, 'conv')
        b = torch.randn(64, device='cuda')
        y = G.Act(g, 64, 'cuda', f32=False)
        fn = lambda: G.igemm(x, wt, err, out=y, out_f32=False, split=split, bias=b)
        fn(); torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(reps):
                fn()
        med, best = timeit(graph.replay, iters=7, warm=2)
        out.append(dict(op='small_conv', name=name, n=n, hw=s, split=split, tiles_per_sm=round(n * g.r_img / 128 / 148, 2), us_per_launch=med * 1e3 / reps))
    assert int(err.item()) == 0
    return out


def conv_last_bench(iters=5):
    """conv_last (64 -> 1 @ 1280^2) + bilinear base: nine taps as the columns of one 1x1 GEMM + the nine-point sum kernel."""
    from gpemsr_b200 import igemm as G
    s = 1280
    g = G.Geom(1, s, s, True)
    x = G.Act(g, 64, 'cuda', f32=False)
    x.hi.normal_(); x.lo.normal_(std=0.004)
    w = torch.randn(1, 64, 3, 3, device='cuda') * 0.05
    b = torch.randn(1, device='cuda')
    wt = G.Weights(G.taps_as_columns(w), 'conv')
    taps = G.TapCells(g, 1, 'cuda')
    base = torch.rand(1, 1, s // 16, s // 16, device='cuda')
    out = torch.empty(1, 1, s, s, device='cuda')
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    fn = lambda: G.conv3x3_few_outputs(x, wt, taps, err, 3, b, out, 1, base=base, base_scale=16)
    med, best = timeit(fn, iters=iters, warm=2)
    return [dict(op='conv_last_1280', ms=med, ms_best=best, bytes_in=4.0 * 64 * s * s, gbs=(4.0 * 64 * s * s + 4.0 * s * s) / med / 1e6,
                 note='64 -> 1 3x3 @ 1 x 1280^2 (+ bilinear base): round 1 ran this as nine N = 16 tensor-pipe tiles per k-step (~0.2 ms)')]


def flow1_bench():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    c, s = 64, 1250
    x = torch.randn(1, c, s, s, device='cuda')
    fl = torch.nn.functional.avg_pool2d(2.0 * torch.randn(1, 2, s, s, device='cuda'), 5, 1, 2).permute(0, 2, 3, 1).contiguous()
    white = 4.0 * torch.randn(1, s, s, 2, device='cuda')
    byt = 8 * c * s * s + 8 * s * s
    med, best = timeit(lambda: gpemsr_b200.flow_warp(x, fl, 'bilinear', 'border'), flush=flush)
    med_w, _ = timeit(lambda: gpemsr_b200.flow_warp(x, white, 'bilinear', 'border'), flush=flush)
    from oracle.flow_warp import flow_warp_torch
    med_t, _ = timeit(lambda: flow_warp_torch(x, fl, 'bilinear', 'border'), flush=flush)
    return [dict(op='flow_warp', c=c, s=s, ms=med, ms_best=best, gbs=byt / med / 1e6, frac=byt / med / 1e6 / 6550.1, ms_white_noise_flow=med_w,
                 gbs_white=byt / med_w / 1e6, frac_white=byt / med_w / 1e6 / 6550.1, torch_grid_sample_ms=med_t)]


def attn_bench(n=1, c=512, hw=80, iters=5):
    """One NonLocalBlock (model/blocks.py:61-83) at the decoder's shape: GroupNorm, q / k / v^T, row-max pre-pass, scores + exp,
    P v^T / row sum, proj_out -- 7 GEMM launches per image + 3 GroupNorm helpers."""
    from gpemsr_b200 import igemm as G
    from gpemsr_b200.decoder import Decoder, NonLocalBlock, _Plan
    dec = Decoder(dict(channel_list=[c, c], im_channel=1, num_resblock_per_scale=1, num_input_resblck=0, latent_dim=c, use_non_local=True)).cuda()
    nl = [m for m in dec.feat_extract if isinstance(m, NonLocalBlock)][0]
    P = _Plan(dec, n, hw, hw, torch.device('cuda'))
    g = G.Geom(n, hw, hw, True)
    x = G.Act(g, c, 'cuda', f32=True)
    G.pack_nchw(torch.randn(n, c, hw, hw, device='cuda'), x)
    fn = lambda: dec._non_local(P, 'nl', nl, x)
    med, best = timeit(fn, iters=iters, warm=1)
    t = hw * hw
    fl = n * (4.0 * t * t * c + 8.0 * t * c * c)
    return [dict(op='non_local_block', n=n, c=c, tokens=t, ms=med, ms_best=best, tflops=fl / med / 1e9,
                 note='GroupNorm + q/k/v + fused-softmax attention + proj_out; algorithmic FLOPs 4 T^2 C + 8 T C^2 per image')]


if __name__ == '__main__':
    torch.cuda.init()
    res = []
    if 'attn' in sys.argv[1:]:
        res += attn_bench()
    if 'last' in sys.argv[1:]:
        res += conv_last_bench()
    if 'flow1' in sys.argv[1:]:
        res += flow1_bench()
    if 'flow' in sys.argv[1:] or len(sys.argv) == 1:
        res += flow_warp_bench()
    if 'small' in sys.argv[1:]:
        res += small_conv_bench()
    if 'conv' in sys.argv[1:]:
        res += conv_bench([a for a in sys.argv[2:]] or None)
    if 'vq1' in sys.argv[1:]:
        res += vq_bench(sizes=((1 << 20, 512),))
    if 'vq' in sys.argv[1:] or len(sys.argv) == 1:
        res += vq_bench()
    for r in res:
        print(json.dumps(r))
