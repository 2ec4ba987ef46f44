"""What the fp32-faithful (hi, lo) split costs: the whole model with precision='bf16' (one bf16 MMA per product) vs the default
('fp32': three) -- HR-image error against the CPU oracle on a 16 x 16 window and time per forward on the x16 80 x 80 window."""
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

res = {}
for prec in ('fp32', 'bf16'):
    m = gpemsr_b200.GPEMSR(None, None, precision=prec, **network_kwargs(16)).eval()
    sd = W.fill_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=916)
    m.load_state_dict(sd, strict=True)
    m.cuda()
    x = torch.rand(1, 5, 1, 16, 16, generator=torch.Generator().manual_seed(926))
    out, _ = m(x.cuda())
    with torch.no_grad():
        want, _ = GM.forward(x, sd, 16)
    err = float((out.cpu() - want).abs().max())
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
    res[prec] = {'max_abs_err_vs_oracle_16x16': err, 'ms_per_forward_80x80_eager': s.elapsed_time(e) / 5}
    del m
    torch.cuda.empty_cache()
print(json.dumps(res))
