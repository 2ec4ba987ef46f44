"""A stand-in for libgpemsr_b200.so that computes nothing: every entry point checks its argument count against
gpemsr_b200._lib.SIGNATURES and returns OK.  With it (and `is_cuda` forced on) the HOST side of the mirrors runs in a container
without a GPU: names, shapes, buffer plans, checkpoint handling and call lists.  Development / CPU-test aid only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, 'tests')) if p not in sys.path]
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


def install():
    _lib._lib = Fake()
    _lib.stream_ptr = lambda: None
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.cuda.current_device = lambda: 0
    torch.cuda.current_stream = lambda device=None: type('S', (), {'cuda_stream': 0})()
    from gpemsr_b200 import igemm as G
    G.post_error_check = lambda device: None          # (the read-back of the device error flag needs a real device)
    G.poll_error = lambda device, wait=False: None
