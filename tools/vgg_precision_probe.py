"""Probe: the VGG relu1_2 similarity-mask branch (model/GPEMSR.py:344-353) in ONE bf16 pass while everything else keeps the
fp32-faithful split.  The mask is a cosine over 64 x 16 x 16 = 16384 products per patch: independent rounding errors average out.
Prints mask error, HR-image error vs the CPU oracle (16 x 16 window, x16) and the time per forward on the 80 x 80 window."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import gpemsr_b200  # noqa: E402
from full_model_util import network_kwargs  # noqa: E402
from gpemsr_b200 import synth_weights as W  # noqa: E402
from oracle import gpemsr_model as GM  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

res = {}
for tag, split in (('vgg_split3', 3), ('vgg_split1', 1)):
    m = gpemsr_b200.GPEMSR(None, None, **network_kwargs(16)).eval()
    sd = W.fill_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=916)
    m.load_state_dict(sd, strict=True)
    m.cuda()
    m.vgg.split = split
    errs, merrs = [], []
    for seed in (926, 927, 928):
        x = torch.rand(1, 5, 1, 16, 16, generator=torch.Generator().manual_seed(seed))
        m.debug = {}
        out, ref_img = m(x.cuda())
        logit = m.debug['mask_logit'].cpu()
        m.debug = None
        taps = {}
        with torch.no_grad():
            want, _ = GM.forward(x, sd, 16, taps)
        errs.append(float((out.cpu() - want).abs().max()))
        merrs.append(float((torch.sigmoid(logit) - taps['mask']).abs().max()))
        # the raw cosine itself
        cos = m.vgg.similarity_mask(ref_img[0].contiguous(), x[0].cuda(), 16).cpu()
        with torch.no_grad():
            cos_w = R.similarity_mask(ref_img[0].cpu(), x[0], {k[4:]: v for k, v in sd.items() if k.startswith('vgg.')}, 16)
        merrs.append(float((cos - cos_w).abs().max()))
    xb = torch.rand(1, 5, 1, 80, 80, device='cuda')
    for _ in range(3):
        m(xb)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        m(xb)
    e.record()
    torch.cuda.synchronize()
    m.check()
    res[tag] = {'out_err': errs, 'mask_err_then_cos_err': merrs, 'ms_per_forward_80x80_eager': s.elapsed_time(e) / 5}
    del m
    torch.cuda.empty_cache()
print(json.dumps(res))
