#!/usr/bin/env python
"""bench.py -- GPEMSR inference hot path on B200: HR megapixels / s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--no-cpu-baseline] [--no-micro] [--no-graph]
    python bench.py [--gpus N] --volume          # BASELINE configs[4]: the 125-slice x8 volume, strong scaling over N
    python bench.py --profile-step               # one eager step between cudaProfilerStart/Stop (for ncu --profile-from-start off)

A "step" is ONE WHOLE FORWARD of the model on one slice window of BASELINE.json configs[1] (GPEMSR x16, N_frames = 5 per
option/output_GPEMSR_x16.yml, 80x80 LR -> 1280x1280 HR; 78x78 cannot run through the reference):
``gpemsr_b200.GPEMSR.forward(x[1, 5, 1, 80, 80]) -> (out[1, 1, 1280, 1280], ref_img[1, 5, 1, 1280, 1280])``, i.e. everything
model/GPEMSR.py:323-456 does:

    per-frame LR features (conv_first + 5 residual blocks) and their ConvTranspose pyramid
    f-4  Indexer16 conv stack -> a-2 head + softmax/top-1 + codebook gather -> a-3 Decoder.multi_scale_feat_calculate
    f-1  VGG19 relu1_2 patch-similarity mask (5 x 1280 x 1280, 16 x 16 patch cosine) -> refmaskconv1..3
    reference-feature fusion at 640^2 / 320^2 / 160^2 / 80^2 (reffusionconv, fusion blocks, mask multiply, strided convs)
    POD alignment of the 5 (neighbour, centre) pairs: f-3 SpyNet at 320^2 (with a-5 flow_warp inside), strided flow convs,
         3-level offset pyramid, four DCNv2Pack deformable convolutions
    ThreeDA fusion (temporal attention, Conv3d frame mixing, 3-level spatial attention)
    a-4  the SR tail (10 residual blocks, 4 x [conv -> PixelShuffle -> LeakyReLU], HRconv, conv_last, + bilinear base)

`value` times the step with inputs resident in HBM; `e2e` re-times it through the same public API with every step
input coming from pinned host memory and the HR slice read back.  N > 1: one process per GPU (torchrun), each rank
processes its own slice windows (weak scaling, no collective on the hot path); the HR slices are all-gathered (NCCL)
once per step.  `--impl reference` times the CPU restatement of the reference's whole forward (oracle/gpemsr_model.py, pinned
bit-exact to model/GPEMSR.py; PyTorch fp32, all host cores) on a bounded LR crop of the same window.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCALE, NFRAMES, LR = 16, 5, 80
METRIC, UNIT = 'hr_megapixels_per_s', 'MP/s'
ARGREF = {'Indexer16': dict(channel_list=[64, 64, 128, 256, 512], im_channel=1, num_resblock_per_scale=2, num_output_resblck=3,
                            latent_dim=512, use_non_local=True),
          'Codebook': dict(num_codebook_vectors=1024, latent_dim=512, beta=1),
          'Decoder': dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=3,
                          latent_dim=512, use_non_local=True)}
NET = dict(argref=ARGREF, nf=64, nframes=NFRAMES, groups=8, front_RBs=5, back_RBs=10, w_ref=True, ref_fusion_feat_RBs=1,
           align_mode='POD', fusion_mode='ThreeDA', mode='16to1', scale=SCALE)       # option/output_GPEMSR_x16.yml network block


def make_inputs(lr, nframes, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    ins = {'x': torch.rand(1, nframes, 1, lr, lr, generator=g)}            # the LR slice window (EM intensities in [0, 1])
    if pin:
        ins = {k: v.pin_memory() for k, v in ins.items()}
    return ins


_MODEL = []


def native_model():
    """The mirror module (CPU, random init); its state_dict names / shapes define the synthetic parameters of BOTH arms."""
    if not _MODEL:
        import gpemsr_b200
        _MODEL.append(gpemsr_b200.GPEMSR(None, None, **NET).eval())
    return _MODEL[0]


def make_weights(seed=1):
    from gpemsr_b200 import synth_weights as W      # deterministic random-init parameters (no checkpoints offline)
    return W.fill_state({k: tuple(v.shape) for k, v in native_model().state_dict().items()}, seed=seed)


# ----------------------------------------------------------------------------------------------- native arm
class NativeHotPath:
    def __init__(self, wts, device):
        self.model = native_model()
        self.model.load_state_dict(wts, strict=True)
        self.model.to(device)

    def step(self, d):
        out, ref_img = self.model(d['x'])
        return out, [ref_img]


# ----------------------------------------------------------------------------------------------- reference (CPU) arm
def cpu_step(wts, ins):
    from oracle import gpemsr_model as GM           # the restatement of model/GPEMSR.py pinned bit-exact to the reference
    with torch.no_grad():
        out, ref_img = GM.forward(ins['x'], wts, SCALE)
    return out, [ref_img]


def cpu_time(wts, lr, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    ins = make_inputs(lr, NFRAMES, seed=7)
    for _ in range(warmup):
        cpu_step(wts, ins)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(wts, ins)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, (SCALE * lr) ** 2 / 1e6 / dt


def pick_cpu_crop(wts, budget_s):
    """Largest LR crop (multiple of 4, 16..80) whose forward fits `budget_s`, from a 16x16 probe (cost ~ pixels)."""
    dt, _ = cpu_time(wts, 16, 1, 1)
    per_px = dt / 256.0
    lr = int((budget_s / per_px) ** 0.5) // 4 * 4
    return max(16, min(LR, lr))


# ----------------------------------------------------------------------------------------------- GPU-eager baseline
def gpu_eager_baseline(dev, steps=5, warmup=2):
    """the reference's own device path -- PyTorch eager on the same B200 -- timed beside the native
    arm (a baseline leg: it may execute oracle/).  `oracle/gpemsr_model.py` is the functional restatement of model/GPEMSR.py
    (pinned bit-exact on CPU) and dispatches to the same cuDNN / cuBLAS / ATen kernels as the reference's nn.Modules.  Two
    numerics: TF32 off (the parity oracle's setting) and PyTorch's defaults (cuDNN may use TF32: what a user of the reference
    gets).  Workloads: the configs[1] window and one window of configs[4] (x8, 5 x 156 x 156)."""
    from oracle import gpu_eager as GE
    import gpemsr_b200
    from gpemsr_b200 import synth_weights as W
    out = {'what': 'oracle/gpemsr_model.py (restatement of model/GPEMSR.py) in PyTorch eager on cuda: cuDNN / cuBLAS / ATen kernels; '
                   f'CUDA events, {warmup} warm-up + {steps} timed forwards each'}
    for tag, scale, lr, net in (('configs1_x16_5x80x80', SCALE, LR, NET),
                                ('configs4_window_x8_5x156x156', VOL_SCALE, VOL_LR, vol_net())):
        m = gpemsr_b200.GPEMSR(None, None, **net)
        sd = W.fill_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=1 if scale == SCALE else 2)
        del m
        sd_dev = GE.to_device(sd, dev)
        x = torch.rand(1, NFRAMES, 1, lr, lr, generator=torch.Generator().manual_seed(100)).to(dev)
        ent = {}
        for name, tf32 in (('tf32_off', False), ('pytorch_default_tf32_conv', True)):
            for _ in range(warmup):
                GE.forward(x, sd_dev, scale, tf32=tf32)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                GE.forward(x, sd_dev, scale, tf32=tf32)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / steps
            ent[name] = {'ms_per_step': ms, 'value': (scale * lr) ** 2 / 1e6 / (ms / 1e3), 'unit': UNIT, 'tf32': tf32}
        out[tag] = ent
        del sd_dev
        torch.cuda.empty_cache()
    return out


def vol_net():
    return dict(NET, mode='8to1', scale=VOL_SCALE, argref={'Indexer8': ARGREF['Indexer16'], 'Codebook': ARGREF['Codebook'],
                                                           'Decoder': ARGREF['Decoder']})


# ----------------------------------------------------------------------------------------------- helpers
_REAL_STDOUT = []


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL's version banner ignores NCCL_DEBUG_FILE
    here): keep a private duplicate of fd 1 for that line and point fd 1 at stderr for everything else."""
    if not _REAL_STDOUT:
        sys.stdout.flush()
        _REAL_STDOUT.append(os.fdopen(os.dup(1), 'w'))
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT[0] if _REAL_STDOUT else sys.stdout
    print(json.dumps(line), file=out, flush=True)


class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '25',
                                          '-i', str(self.index)], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.path or not os.path.exists(self.path):
            return None
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            p = [x.strip() for x in line.split(',')]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return None
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': mx, 'samples': len(sm), 'reasons': sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf=d.get('bf16_tflops_sustained', d['bf16_tflops']), tf_burst=d['bf16_tflops'], src='measured')
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src='fallback')


def instrumented_step(hp, dev_in):
    """One extra step with CUDA events around every GEMM launch (on the launching stream): per-kernel-class time and
    algorithmic FLOPs (2 * valid rows * cols * k * taps -- one pass, no padding, no split factor)."""
    from gpemsr_b200 import igemm as G
    rec = []
    orig = G.igemm

    def wrapped(a, w, err, **kw):
        geom = kw.get('a_geom') or a.geom
        rows = geom.n * geom.h * geom.w
        if w is not None:
            n, k, taps = (kw.get('n_cols') or w.n), w.k, len(w.taps)
        else:
            n, k, taps = kw['n_cols'], kw['k_pad'], 1
        bn = 16 if (n <= 16 and not kw.get('pixel_shuffle')) else 64 if n <= 64 else 128 if n <= 128 else 256
        # the kernel template gpemsr_igemm() launches: tap-/dy-fused (recorded on the weights), else the streaming kernel -- its
        # CTA-pair form (cta_group::2) for every tile width >= 64 unless GPEMSR_PAIR=0 / GPEMSR_TMA=0 / GPEMSR_CLUSTER=0
        name = getattr(w, 'kernel', None) or f'gemm_kernel<{bn}>'
        if name.startswith('gemm_kernel<') and int(name[12:-1]) >= 64 and \
                all(os.environ.get(v, '1') != '0' for v in ('GPEMSR_PAIR', 'GPEMSR_TMA', 'GPEMSR_CLUSTER')):
            name = 'gemm_pair_kernel<' + name[12:]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(a, w, err, **kw)
        e.record()
        # merged ConvTranspose phases / space-to-depth stride-2 convs launch 16 (tap, phase) blocks where the reference layer has 9 taps
        rec.append((name, 2.0 * rows * n * k * taps * getattr(w, 'flop_scale', 1.0), s, e, (rows, n, k, taps)))

    G.igemm = wrapped
    import gpemsr_b200.decoder as D
    import gpemsr_b200.sr_tail as S
    try:
        hp.step(dev_in)
        torch.cuda.synchronize()
    finally:
        G.igemm = orig
    agg = {}
    dump = os.environ.get('GPEMSR_BENCH_DUMP_LAUNCHES')
    rows_out = []
    for name, fl, s, e, shape in rec:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += fl; a[1] += s.elapsed_time(e); a[2] += 1
        rows_out.append(dict(kernel=name, rows=shape[0], n=shape[1], k=shape[2], taps=shape[3], ms=s.elapsed_time(e),
                             tflops=fl / max(s.elapsed_time(e), 1e-6) / 1e9))
    if dump:
        with open(dump, 'w') as f:
            json.dump(rows_out, f, indent=0)
    return agg


def micro_rooflines(peaks):
    """The two kernel-level numbers BASELINE.json's metric also names: VQ lookup TC fraction and flow_warp HBM fraction."""
    import gpemsr_b200
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def med(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]

    n, d = 1 << 20, 512
    z = torch.randn(1, d, n, 1, device='cuda')
    emb = torch.randn(1024, d, device='cuda')
    ms = med(lambda: gpemsr_b200.vq_lookup(z, emb))
    fl = 2.0 * n * 1024 * d
    out['vq_lookup'] = {'bound': 'tensor', 'shape': f'N=2^20 x D={d} vs 1024 codes', 'ms': ms, 'achieved': fl / ms / 1e9,
                        'peak': peaks['tf_burst'], 'unit': 'TFLOP/s', 'frac': fl / ms / 1e9 / peaks['tf_burst'],
                        'note': 'whole lookup (prep + GEMM + re-score + gather), L2 flushed'}
    del z
    c, s = 64, 1250
    x = torch.randn(1, c, s, s, device='cuda')
    f = torch.nn.functional.avg_pool2d(2.0 * torch.randn(1, 2, s, s, device='cuda'), 5, 1, 2).permute(0, 2, 3, 1).contiguous()
    ms = med(lambda: gpemsr_b200.flow_warp(x, f, 'bilinear', 'border'))
    by = 8.0 * c * s * s + 8.0 * s * s
    out['flow_warp'] = {'bound': 'hbm', 'shape': f'{c} x {s}^2', 'ms': ms, 'achieved': by / ms / 1e6, 'peak': peaks['hbm'],
                        'unit': 'GB/s', 'frac': by / ms / 1e6 / peaks['hbm']}
    # SpyNet alone, on the window's 10 (neighbour, centre) pairs at 320^2
    try:
        from gpemsr_b200.spynet import SpyNet, resize_bilinear
        from gpemsr_b200 import synth_weights as W
        spy = SpyNet().cuda()
        spy.load_state_dict({**W.fill(W.spynet_spec(), 6, gain=2.0), 'mean': spy.mean, 'std': spy.std}, strict=True)
        fr = torch.rand(NFRAMES, 1, LR, LR, device='cuda')
        x4 = resize_bilinear(fr, 4 * LR, 4 * LR, False, scale=4)
        idx = torch.arange(NFRAMES, device='cuda').repeat(2)
        ref, supp = x4.index_select(0, idx), x4[NFRAMES // 2:NFRAMES // 2 + 1].expand(2 * NFRAMES, -1, -1, -1).contiguous()
        ms = med(lambda: spy(ref, supp), iters=3, warm=2)
        fl = 2.0 * 49 * (8 * 32 + 32 * 64 + 64 * 32 + 32 * 16 + 16 * 2) * ref.shape[0] * (4 * LR) ** 2 * (4.0 / 3.0)
        out['spynet'] = {'bound': 'tensor', 'shape': f'{ref.shape[0]} pairs x {4 * LR}^2, 6 levels', 'ms': ms, 'achieved': fl / ms / 1e9,
                         'peak': peaks['tf'], 'unit': 'TFLOP/s', 'frac': fl / ms / 1e9 / peaks['tf'],
                         'note': 'dy-fused 49-tap GEMMs + resize / pool / pack / flow_warp helpers, 6 levels'}
    except Exception as e:                                # a diagnostic extra must never take the bench line down
        out['spynet'] = {'error': repr(e)[:200]}
    return out


# ----------------------------------------------------------------------------------------------- config 5: whole volume
VOL_SLICES, VOL_LR, VOL_SCALE = 125, 156, 8          # BASELINE configs[4]: 125 slices, x8, 156^2 LR -> 1248^2 HR


def volume_bench(dev, rank, world, dist, passes=2, n_slices=VOL_SLICES):
    """BASELINE configs[4]: the reference's output loop (output_GPEMSR.py:54-128) over a synthetic 125-slice x8 volume,
    slice-sharded over `world` ranks (strong scaling).  Timed end to end per pass: H2D of the LR volume from pinned host
    memory, per-slice encoding (each slice ONCE: the per-frame cache of SURVEY.md 8f-2), the 125 window fusions, the
    all-gather of the HR slices (N > 1) and the D2H of this rank's HR block."""
    import gpemsr_b200
    from gpemsr_b200 import synth_weights as W
    from gpemsr_b200.volume import gather_slices, shard_range
    # (Indexer8 takes the same argument block as Indexer16 in option/output_GPEMSR_x8.yml; only the class differs)
    model = gpemsr_b200.GPEMSR(None, None, **vol_net()).eval()
    model.load_state_dict(W.fill_state({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=2), strict=True)
    model.to(dev)
    vol_host = torch.rand(n_slices, 1, VOL_LR, VOL_LR, generator=torch.Generator().manual_seed(4)).pin_memory()
    lo, hi = shard_range(n_slices, world, rank)
    hr = VOL_SCALE * VOL_LR
    out_host = torch.empty(hi - lo, 1, hr, hr).pin_memory()
    out_dev = torch.empty(hi - lo, 1, hr, hr, device=dev)

    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # halo: the 2 boundary slices' features come from the neighbour ranks (GPEMSR_HALO=recompute: encoded again locally, round 1)
    exchange = world > 1 and os.environ.get('GPEMSR_HALO', 'exchange') == 'exchange'

    def one_pass():
        vol = vol_host.to(dev, non_blocking=True)
        model.forward_volume(vol, lo, hi, out=out_dev, halo_exchange=(dist, rank, world) if exchange else None)
        g0.record()
        if world > 1:
            gather_slices(out_dev, n_slices, world, rank, dist)
        g1.record()
        out_host.copy_(out_dev, non_blocking=True)

    one_pass()                                          # warm-up: builds plans, packs weights
    model.check()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = gpemsr_b200.kernel_launches()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(passes):
        one_pass()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / passes
    halo = 0 if exchange else (min(hi + NFRAMES // 2, n_slices) - max(lo - NFRAMES // 2, 0)) - (hi - lo)
    return {'value': n_slices * hr * hr / 1e6 / (ms / 1e3), 'unit': UNIT, 'ms_per_volume': ms, 'ms_per_slice': ms / n_slices * world,
            'slices_rank0': hi - lo, 'halo_slices_encoded_rank0': halo,
            'halo': 'exchanged with the neighbour ranks (P2P, 17 MB of features per slice)' if exchange else 'recomputed locally' if world > 1 else 'none', 'gather_ms_rank0': g0.elapsed_time(g1) if world > 1 else 0.0,
            'workload': f'{n_slices} slices x{VOL_SCALE}, {VOL_LR}^2 LR -> {hr}^2 HR, windows of output_GPEMSR.py:54-128, every slice '
                        f'encoded once (per-frame cache), slice blocks over {world} GPU(s), HR slices all-gathered',
            'h2d_bytes': vol_host.numel() * 4, 'd2h_bytes': out_host.numel() * 4, 'scaling': 'strong',
            'gpu_launches_per_pass': int((gpemsr_b200.kernel_launches() - l0) / passes), 'passes': passes}


# ----------------------------------------------------------------------------------------------- main
def run_reference(args, rank, world):
    """The CPU arm.  `kind: "port"`: oracle/gpemsr_model.py, the restatement pinned bit-exact to the reference's model/GPEMSR.py
    (the reference itself is Python under /root/reference, which does not exist on the GPU box and may not be read at run
    time).  A step is the whole forward on the FULL configs[1] window when steps + warm-up fit ~5 minutes of host time,
    otherwise on the largest LR crop that does -- `config` names what was really timed (`lr`, `cpu_crop`)."""
    if rank != 0:
        return
    wts = make_weights()
    budget = 300.0 / max(args.steps + args.warmup, 1)
    lr = pick_cpu_crop(wts, budget)
    dt, mps = cpu_time(wts, lr, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = (f'the full {NFRAMES}x{LR}x{LR} window' if lr == LR else f'{NFRAMES}x{lr}x{lr} LR crop of the {NFRAMES}x{LR}x{LR} window') + \
             f' ({SCALE * lr}^2 HR px per step), oracle/gpemsr_model.py on {cores} host cores'
    cfg = config_block(1, lr=lr)
    cfg['parallelism'] = f'CPU only: {cores} host threads, no GPU used (n_gpus echoes --gpus for the driver)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': mps, 'unit': UNIT, 'n_gpus': args.gpus, 'gpus_used': 0, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': mps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': mps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def config_block(world, lr=LR):
    crop = '' if lr == LR else f' -- TIMED ON A {lr}x{lr} LR CROP ({SCALE * lr}x{SCALE * lr} HR), MP/s normalises the size'
    return {'workload': f'GPEMSR x16 whole forward (gpemsr_b200.GPEMSR.forward = model/GPEMSR.py:323-456: LR features, Indexer16 + codebook '
                        f'lookup + VQ decoder, VGG similarity mask, reference fusion, POD alignment incl. SpyNet / flow_warp / DCNv2, ThreeDA, SR tail), '
                        f'{NFRAMES}-slice window {LR}x{LR} LR -> {SCALE * LR}x{SCALE * LR} HR, random-init weights' + crop,
            'lr': lr, 'cpu_crop': lr != LR, 'n_frames': NFRAMES, 'scale': SCALE, 'units_per_step': 'one output slice per GPU',
            'parallelism': f'slice-sharded x{world}, outputs all-gathered', 'l2': 'working set per step (>2 GB of activations) '
            'exceeds the 126 MB L2; no explicit flush',
            'precision': 'default precision plan (gpemsr_b200.gpemsr.DEFAULT_PLAN, profiles/r02_precision_plan.json): bf16 x3 split '
                         '(fp32-faithful) on tcgen05 everywhere except the VGG similarity branch and SpyNet (one bf16 pass)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-micro', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step kernel by kernel instead of replaying a CUDA graph')
    ap.add_argument('--volume', action='store_true', help='time BASELINE configs[4] instead (125-slice x8 volume, strong scaling over --gpus)')
    ap.add_argument('--profile-step', action='store_true', help='warm up, run ONE eager step between cudaProfilerStart/Stop and exit '
                    '(for `ncu --profile-from-start off ...`: the launch list of exactly one step; prints no bench line)')
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')      # keep NCCL's version banner off stdout: ONE JSON line there
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import gpemsr_b200
    if args.volume:
        v = volume_bench(dev, rank, world, dist)
        if rank == 0:
            emit({'metric': METRIC, 'value': v['value'], 'unit': UNIT, 'n_gpus': world, 'steps': v['passes'], 'warmup': 1,
                              'ms_per_step': v['ms_per_volume'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                              'dtype': 'bf16x3->f32', 'data': 'synthetic', 'config': {'workload': v['workload']}, 'volume': v,
                  'impl': 'native'})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    wts = make_weights()
    hp = NativeHotPath(wts, dev)
    host_in = make_inputs(LR, NFRAMES, seed=100 + rank, pin=True)
    dev_in = {k: v.to(dev) for k, v in host_in.items()}
    hr_px = (SCALE * LR) ** 2
    gathered = torch.empty(world, 1, 1, SCALE * LR, SCALE * LR, device=dev) if world > 1 else None

    if args.profile_step:
        for _ in range(3):
            hp.step(dev_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        hp.step(dev_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    graphed = {}

    def step(d):
        if args.no_graph:
            out, feats = hp.step(d)
        else:                                           # one CUDA graph per static input set, captured on first use
            key = id(d)
            if key not in graphed:
                from gpemsr_b200.graph import GraphedStep
                graphed[key] = GraphedStep(hp.step, d)
            out, feats = graphed[key]()
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.unsqueeze(0))
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    # nvidia-smi needs ~100 ms to start sampling: it runs from the warm-up on (same kernels, same load) through the timed region
    with ClockSampler(local) as cs:
        for _ in range(W):
            step(dev_in)
        hp.model.check()

        # ---- value: inputs resident in HBM
        barrier()
        l0 = gpemsr_b200.kernel_launches()
        hp.step(dev_in)                                 # one eager step: counts the kernels a step launches (graph replays
        launches_per_step = gpemsr_b200.kernel_launches() - l0      # re-issue the same kernels without passing the C ABI)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            step(dev_in)
        e.record()
        barrier()
    launches = launches_per_step * args.steps
    ms = s.elapsed_time(e)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps * hr_px / 1e6 / (ms_total / 1e3)

    # ---- e2e: every step input from pinned host memory, HR slice read back.  The volume driver's schedule: the H2D copy
    # of step i+1 runs on a copy stream while step i computes (two staging sets), the HR slice is read back per step.
    stage = [{k: torch.empty_like(v) for k, v in dev_in.items()} for _ in range(2)]
    out_host = [torch.empty(1, 1, SCALE * LR, SCALE * LR).pin_memory() for _ in range(2)]
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    d2h = out_host[0].numel() * 4
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]     # staging set filled
    freed = [torch.cuda.Event() for _ in range(2)]     # staging set consumed by its step

    def e2e_run(n_steps):
        for b in range(2):
            freed[b].record(main_stream)
        for i in range(n_steps + 1):
            if i < n_steps:                            # prefetch step i
                b = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[b])
                    for k in stage[b]:
                        stage[b][k].copy_(host_in[k], non_blocking=True)
                    ready[b].record(copy_stream)
            if i >= 1:                                 # compute step i-1
                b = (i - 1) & 1
                main_stream.wait_event(ready[b])
                out = step(stage[b])
                freed[b].record(main_stream)
                out_host[b].copy_(out, non_blocking=True)

    e2e_run(2)
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    e2e_run(args.steps)
    e2.record()
    barrier()
    t2 = torch.tensor([s2.elapsed_time(e2)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = world * args.steps * hr_px / 1e6 / (float(t2.item()) / 1e3)

    if rank == 0:
        peaks = load_peaks()
        agg = instrumented_step(hp, dev_in)
        name, (fl, tms, cnt) = max(agg.items(), key=lambda kv: kv[1][1])
        step_ms = ms_total / args.steps
        # DRAM bytes per launch of that kernel class: from the committed ncu pass over one eager step (profiles/README.md)
        prof = os.path.join(ROOT, 'profiles', 'r02_full_traffic.json')
        traffic = None
        if os.path.exists(prof):
            ent = json.load(open(prof))['per_step'].get(name)
            traffic = ent['traffic_per_launch_bytes'] if ent else None
        roof = {'bound': 'tensor', 'kernel': name, 'achieved': fl / tms / 1e9, 'peak': peaks['tf'], 'unit': 'TFLOP/s',
                'frac': fl / tms / 1e9 / peaks['tf'], 'traffic': traffic, 'launches_per_step': cnt,
                'ms_per_launch': tms / cnt, 'share_of_step': tms / step_ms, 'peak_source': peaks['src'] + ' (bf16 sustained)',
                'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02_full_traffic.json)',
                'note': 'algorithmic FLOPs = 2*rows*cols*k*taps (one pass); the kernel issues 3 bf16 MMAs per product '
                        '(fp32-faithful split), so frac <= 1/3 by construction',
                'by_kernel': {k: {'tflops': v[0] / v[1] / 1e9, 'ms': v[1], 'launches': v[2]} for k, v in agg.items()}}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': W,
                'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16x3->f32',
                'data': 'synthetic', 'config': config_block(world), 'roofline': roof,
                'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
                'gpu_launches': int(launches), 'clocks': cs.summary(), 'impl': 'native',
                'cuda_graph': not args.no_graph}
        if world == 1 and not args.no_micro:
            line['micro'] = micro_rooflines(peaks)
    # BASELINE configs[4] (the 125-slice x8 volume, STRONG scaling over the ranks) rides in the default line at every N, so the
    # driver's scaling series carries it; every rank takes part (slice blocks + one all-gather), rank 0 reports
    vol = None
    if not args.no_micro:
        try:
            del hp, graphed
            torch.cuda.empty_cache()
            vol = volume_bench(dev, rank, world, dist, passes=1 if world == 1 else 2)
        except Exception as ex:                                  # an extra must never take the bench line down
            vol = {'error': repr(ex)[:300]}
    if rank == 0:
        if vol is not None:
            line['volume'] = vol
        if world == 1 and not args.no_micro:
            try:
                torch.cuda.empty_cache()
                line['gpu_eager_baseline'] = gpu_eager_baseline(dev)
                ge = line['gpu_eager_baseline']['configs1_x16_5x80x80']
                line['speedup_vs_gpu_eager'] = {k: ge[k]['ms_per_step'] / step_ms for k in ge}
            except Exception as ex:
                line['gpu_eager_baseline'] = {'error': repr(ex)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            lr = pick_cpu_crop(wts, 20.0)
            dt, mps = cpu_time(wts, lr, 1, 0)
            line['cpu_baseline'] = {'value': mps, 'unit': UNIT, 'cores': os.cpu_count() or 1, 'kind': 'port',
                                    'sample': f'1 step on a {NFRAMES}x{lr}x{lr} LR crop of the window ({dt:.1f} s of CPU work), '
                                              'oracle/gpemsr_model.py (whole forward) on all host cores'}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
