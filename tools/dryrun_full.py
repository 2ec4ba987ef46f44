"""CPU dry run of gpemsr_b200.GPEMSR.forward with the C library mocked out (every entry point returns OK after checking its
argument count against gpemsr_b200._lib.SIGNATURES): catches host-side mistakes (names, shapes, argument lists) without
a GPU.  Development aid only; computes nothing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _mock_lib  # noqa: E402
from _mock_lib import calls  # noqa: E402

_mock_lib.install()

from full_model_util import build  # noqa: E402

for scale, hw in ((8, (16, 16)), (16, (20, 24))):
    m, sd = build(scale)
    m.debug = {}
    out, ref = m(torch.rand(1, 5, 1, *hw))
    print(scale, tuple(out.shape), tuple(ref.shape), len(calls), 'calls;', sorted(m.debug))
    calls.clear()

m, sd = build(8)
vol = torch.rand(7, 1, 16, 16)
o = m.forward_volume(vol)
print('volume', tuple(o.shape), len(calls), 'calls')
o = m.forward_volume(vol, 2, 5)
print('volume block', tuple(o.shape))
