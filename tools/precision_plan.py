"""Precision plan: where may a GEMM run ONE bf16 pass instead of the fp32-faithful 3-term split?

The criterion is BASELINE.json's, not a per-stage one: HR image within 1e-3 max-abs of the reference's own device path
(PyTorch eager on the same GPU, TF32 off), PSNR delta < 0.01 dB, codebook indices unchanged.  For every layer group the tool
builds the model with that group at split 1 (everything else at 3), runs BASELINE configs[1] (x16, 5 x 80 x 80) and measures

    * the HR-image / reference-image error against the GPU-eager forward (following the same codebook indices),
    * index flips against the all-split-3 run,
    * the step time (CUDA-graph replay, CUDA events),

then adds groups greedily (largest time saving first) while the accumulated HR error stays below the budget, and checks the
chosen plan on the CREMI x8 window (5 x 156 x 156).  Output: gpurun_out/r02_precision_plan.json (copied to profiles/).

    python tools/precision_plan.py [--budget 3e-4] [--quick]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import gpemsr_b200  # noqa: E402
from gpemsr_b200.graph import GraphedStep  # noqa: E402
from oracle import gpu_eager as GE  # noqa: E402
from oracle import weights as W  # noqa: E402
from full_model_util import network_kwargs  # noqa: E402

GROUPS = {
    'vgg': ['vgg'],
    'spynet': ['spynet'],
    'pod': ['pod'],
    'tda': ['tda'],
    'tail.trunk': ['tail.rt'],
    'tail.up0': ['tail.up0'], 'tail.up1': ['tail.up1'], 'tail.up2': ['tail.up2'], 'tail.up3': ['tail.up3'],
    'tail.hr': ['tail.hr'],
    'tail.last': ['tail.last'],
    'enc.mask': ['enc.mask'],
    'enc.lr_features': ['enc.conv_first', 'enc.fe', 'enc.reffea'],
    'enc.reffusion': ['enc.reffusionconv'],
    'enc.fusion_blocks': ['enc.ffb'],
    'enc.down': ['enc.down_fea_conv', 'enc.reduce_dim_conv', 'enc.fea_L'],
    'decoder.attention': ['decoder.feat_extract.0'],
    'decoder.512': ['decoder.input_layer', 'decoder.feat_extract.1', 'decoder.feat_extract.2'],
    'decoder.256-64': ['decoder.feat_extract.3', 'decoder.feat_extract.4', 'decoder.feat_extract.5', 'decoder.feat_extract.6',
                       'decoder.feat_extract.7', 'decoder.feat_extract.8', 'decoder.final', 'decoder.output_layer'],
    'indexer': ['indexer'],
}


def build(scale, sd, table):
    m = gpemsr_b200.GPEMSR(None, None, precision=dict(table, default=3), **network_kwargs(scale))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def measure(scale, sd, x, table, ref, idx0, steps):
    m = build(scale, sd, table)
    out, ref_img = m(x)
    m.check()
    idx = m.refmodel.codebook.last_idx.clone()
    flips = int((idx != idx0).sum()) if idx0 is not None else 0
    res = dict(flips=flips)
    if ref is not None:
        res['out_err'] = float((out - ref[0]).abs().max())
        res['ref_img_err'] = float((ref_img - ref[1]).abs().max())
        res['out_mse'] = float(((out - ref[0]) ** 2).mean())
    g = GraphedStep(lambda d: m(d['x']), {'x': x})
    for _ in range(2):
        g()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        g()
    e.record()
    torch.cuda.synchronize()
    res['ms'] = s.elapsed_time(e) / steps
    del g, m
    torch.cuda.empty_cache()
    return res, idx, (out, ref_img)


def errors_only(scale, sd, x, table, ref, idx0):
    m = build(scale, sd, table)
    out, ref_img = m(x)
    m.check()
    idx = m.refmodel.codebook.last_idx.clone()
    res = dict(flips=int((idx != idx0).sum()) if idx0 is not None else 0)
    if ref is not None:
        res['out_err'] = float((out - ref[0]).abs().max())
        res['out_mean_abs'] = float((out - ref[0]).abs().mean())
        res['ref_img_err'] = float((ref_img - ref[1]).abs().max())
    del m
    torch.cuda.empty_cache()
    return res, idx


def verify(cands):
    """The candidate groups (alone and accumulated) on several parameter / input seeds and both shapes: a plan must hold on every one."""
    rep = {}
    for scale, lr, seeds in ((16, 80, (1, 11, 101, 21)), (8, 156, (2, 93, 12, 22)), (8, 16, (400 + 8, 5)), (16, 16, (400 + 16, 6))):
        for seed in seeds:
            probe = gpemsr_b200.GPEMSR(None, None, **network_kwargs(scale))
            sd = W.fill_state({k: tuple(v.shape) for k, v in probe.state_dict().items()}, seed=seed)
            del probe
            x = torch.rand(1, 5, 1, lr, lr, generator=torch.Generator().manual_seed(1000 + seed)).cuda()
            _, idx0 = errors_only(scale, sd, x, {}, None, None)
            sd_dev = GE.to_device(sd)
            ref = GE.forward(x, sd_dev, scale, idx_override=idx0)
            key = f'x{scale}_{lr}_seed{seed}'
            rep[key] = {'all_split3': errors_only(scale, sd, x, {}, ref, idx0)[0]}
            acc = {}
            for g in cands:
                rep[key][g] = errors_only(scale, sd, x, {p: 1 for p in GROUPS[g]}, ref, idx0)[0]
                acc.update({p: 1 for p in GROUPS[g]})
                rep[key]['+'.join(cands[:cands.index(g) + 1])] = errors_only(scale, sd, x, acc, ref, idx0)[0]
            print(key, json.dumps(rep[key]), flush=True)
            del sd_dev, ref
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, 'gpurun_out', 'r02_precision_verify.json'), 'w'), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--budget', type=float, default=3e-4)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--verify', default='', help='comma-separated groups: errors only, on several seeds and both shapes')
    a = ap.parse_args()
    if a.verify:
        verify(a.verify.split(','))
        return
    report = {'budget_hr_max_abs': a.budget, 'criterion': 'HR image vs PyTorch eager on the same GPU (TF32 off) following the same codebook '
              'indices; tolerance of BASELINE.json: 1e-3 max-abs, PSNR delta < 0.01 dB, indices unchanged'}
    scale, lr = 16, 80
    probe = gpemsr_b200.GPEMSR(None, None, **network_kwargs(scale))
    sd = W.fill_state({k: tuple(v.shape) for k, v in probe.state_dict().items()}, seed=1)
    del probe
    x = torch.rand(1, 5, 1, lr, lr, generator=torch.Generator().manual_seed(100)).cuda()
    base, idx0, _ = measure(scale, sd, x, {}, None, None, a.steps)
    sd_dev = GE.to_device(sd)
    ref = GE.forward(x, sd_dev, scale, idx_override=idx0)
    base, idx0, _ = measure(scale, sd, x, {}, ref, idx0, a.steps)
    report['x16_80'] = {'all_split3': base, 'groups': {}}
    print('baseline', base, flush=True)
    names = list(GROUPS) if not a.quick else ['vgg', 'spynet', 'tail.hr']
    for gname in names:
        t0 = time.time()
        r, _, _ = measure(scale, sd, x, {p: 1 for p in GROUPS[gname]}, ref, idx0, a.steps)
        r['ms_saved'] = base['ms'] - r['ms']
        r['out_err_added'] = r['out_err'] - base['out_err']
        report['x16_80']['groups'][gname] = r
        print(gname, r, f'({time.time() - t0:.1f} s)', flush=True)
    # greedy accumulation: most time saved first, among the groups that keep the indices and cost little error on their own
    cand = sorted((g for g, r in report['x16_80']['groups'].items() if r['flips'] == 0 and r['ms_saved'] > 0.05 and r['out_err'] <= a.budget),
                  key=lambda g: -report['x16_80']['groups'][g]['ms_saved'])
    table, chosen, trail = {}, [], []
    for gname in cand:
        trial = dict(table, **{p: 1 for p in GROUPS[gname]})
        r, _, _ = measure(scale, sd, x, trial, ref, idx0, a.steps)
        ok = r['flips'] == 0 and r['out_err'] <= a.budget and r['ref_img_err'] <= 1e-3
        trail.append(dict(add=gname, accepted=ok, **r))
        print('greedy', gname, ok, r, flush=True)
        if ok:
            table, chosen = trial, chosen + [gname]
    report['x16_80']['greedy'] = trail
    report['chosen_groups'] = chosen
    report['chosen_table'] = table
    final, _, _ = measure(scale, sd, x, table, ref, idx0, a.steps)
    report['x16_80']['chosen'] = final
    del sd_dev, ref
    torch.cuda.empty_cache()
    # the chosen plan on the CREMI x8 window
    scale, lr = 8, 156
    probe = gpemsr_b200.GPEMSR(None, None, **network_kwargs(scale))
    sd = W.fill_state({k: tuple(v.shape) for k, v in probe.state_dict().items()}, seed=2)
    del probe
    x = torch.rand(1, 5, 1, lr, lr, generator=torch.Generator().manual_seed(101)).cuda()
    base8, idx8, _ = measure(scale, sd, x, {}, None, None, a.steps)
    sd_dev = GE.to_device(sd)
    ref8 = GE.forward(x, sd_dev, scale, idx_override=idx8)
    base8, _, _ = measure(scale, sd, x, {}, ref8, idx8, a.steps)
    plan8, _, _ = measure(scale, sd, x, table, ref8, idx8, a.steps)
    report['x8_156'] = {'all_split3': base8, 'chosen': plan8}
    print('x8', base8, plan8, flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, 'gpurun_out', 'r02_precision_plan.json'), 'w'), indent=1)
    print(json.dumps({'chosen': chosen, 'x16': final, 'x8': plan8}))


if __name__ == '__main__':
    main()
