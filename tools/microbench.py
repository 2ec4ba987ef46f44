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
             ('rb64_640', 5, 64, 64, 640, 3, {}), ('hr64_1280', 1, 64, 64, 1280, 3, dict(act=G.ACT_LRELU, slope=0.1)),
             ('up256_640', 1, 64, 256, 640, 3, dict(ps=True)), ('out1_1280', 5, 64, 1, 1280, 3, dict(nchw=True)),
             ('q512_80', 5, 512, 512, 80, 1, {})]
    out = []
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    for name, n, ci, co, s, ks, opt in cases:
        if which and name not in which:
            continue
        g = G.Geom(n, s, s, True)
        x = G.Act(g, ci, 'cuda', f32=False)
        x.hi.normal_(); x.lo.normal_(std=0.004)
        w = torch.randn(co, ci, ks, ks, device='cuda') * 0.05
        b = torch.randn(co, device='cuda')
        wt = G.Weights(w, 'conv')
        kw = dict(split=3, bias=b, act=opt.get('act', G.ACT_NONE), slope=opt.get('slope', 0.0))
        if opt.get('ps'):
            y = G.Act(G.Geom(n, 2 * s, 2 * s, True), co // 4, 'cuda', f32=False)
            fn = lambda: G.igemm(x, wt, err, out=y, up=2, pixel_shuffle=True, out_f32=False, **kw)
        elif opt.get('nchw'):
            img = torch.empty(n, co, s, s, device='cuda')
            fn = lambda: G.igemm(x, wt, err, out_nchw=img, nchw_c=co, **kw)
        else:
            y = G.Act(g, co, 'cuda', f32=True)
            fn = lambda: G.igemm(x, wt, err, out=y, **kw)
        med, best = timeit(fn, iters=5, warm=2)
        fl = 2.0 * n * s * s * co * ci * ks * ks
        out.append(dict(op='conv', name=name, n=n, cin=ci, cout=co, hw=s, k=ks, ms=med, tflops=fl / med / 1e9,
                        frac_of_split3_ceiling=fl / med / 1e9 / (1388.4 / 3)))
        del x, wt
    assert int(err.item()) == 0
    return out


if __name__ == '__main__':
    torch.cuda.init()
    res = []
    if 'flow' in sys.argv[1:] or len(sys.argv) == 1:
        res += flow_warp_bench()
    if 'conv' in sys.argv[1:]:
        res += conv_bench([a for a in sys.argv[2:]] or None)
    if 'vq1' in sys.argv[1:]:
        res += vq_bench(sizes=((1 << 20, 512),))
    if 'vq' in sys.argv[1:] or len(sys.argv) == 1:
        res += vq_bench()
    for r in res:
        print(json.dumps(r))
