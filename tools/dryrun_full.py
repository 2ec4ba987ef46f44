"""CPU dry run of gpemsr_b200.GPEMSR.forward with the C library mocked out (every entry point returns OK after checking its
argument count against gpemsr_b200._lib.SIGNATURES): catches host-side mistakes (names, shapes, argument lists) without
a GPU.  Development aid only; computes nothing."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from gpemsr_b200 import _lib  # noqa: E402

calls = []


class Fake:
    def __getattr__(self, name):
        res, args = _lib.SIGNATURES[name]

        def fn(*a):
            assert len(a) == len(args), (name, len(a), len(args))
            calls.append(name)
            if name == 'gpemsr_igemm_plan':
                d = a[0]._obj
                a[1]._obj.value = 16 if d.n_cols <= 16 else 64 if d.n_cols <= 64 else 128 if d.n_cols <= 128 else 256
                a[2]._obj.value = int(d.n_cols <= 64 and d.taps <= 9 and d.k_pad <= 128)
            if name.endswith('_bytes'):
                return 1 << 20
            if name == 'gpemsr_last_error_string':
                return b''
            return 0
        return fn


_lib._lib = Fake()
_lib.stream_ptr = lambda: None
torch.Tensor.is_cuda = property(lambda self: True)
torch.cuda.current_device = lambda: 0
torch.cuda.current_stream = lambda device=None: type('S', (), {'cuda_stream': 0})()

from gpemsr_b200 import igemm as _G  # noqa: E402
_G.post_error_check = lambda device: None          # (the read-back of the device error flag needs a real device)
_G.poll_error = lambda device, wait=False: None

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
