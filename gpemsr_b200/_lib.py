"""ctypes binding of libgpemsr_b200.so (the C ABI declared in include/gpemsr_b200.h).

There is no fallback: if the shared library is missing, or the device is not
sm_100-class, calls raise.  PyTorch is used above this layer only for device
memory and streams (``tensor.data_ptr()``, ``torch.cuda.current_stream()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPEMSR_B200_LIB: another build of the same library (A/B timing, profiling builds of tools/ablate.py); default = the in-tree build
LIB_PATH = os.path.abspath(os.environ['GPEMSR_B200_LIB']) if os.environ.get('GPEMSR_B200_LIB') else os.path.join(_HERE, 'lib', 'libgpemsr_b200.so')

ERRORS = {0: 'OK', -1: 'BAD_SHAPE', -2: 'BAD_ALIGN', -3: 'UNSUPPORTED_ARCH', -4: 'CUDA', -5: 'WORKSPACE',
          -6: 'UNSUPPORTED'}


class GpemsrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'gpemsr_b200: {ERRORS.get(code, code)}: {msg}')
        self.code = code


_p, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
_f = C.c_float

# name -> (restype, argtypes); must list every symbol include/gpemsr_b200.h declares
SIGNATURES = {
    'gpemsr_version': (_i, []),
    'gpemsr_last_error_string': (C.c_char_p, []),
    'gpemsr_device_check': (_i, [_i]),
    'gpemsr_kernel_launches': (_i64, []),
    'gpemsr_tensor_map_stats': (None, [_p, _p]),
    'gpemsr_flow_warp': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    'gpemsr_flow_warp_ex': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    'gpemsr_vq_workspace_bytes': (_sz, [_i64, _i, _i]),
    'gpemsr_vq_lookup_nchw': (_i, [_p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _sz, _p]),
    'gpemsr_logits_argmax_gather': (_i, [_p, _p, _p, _p, _i, _i, _i64, _i, _i, _p, _p, _p, _p, _sz, _p]),
    'gpemsr_argmax_gather': (_i, [_p, _p, _i, _i64, _i, _i, _p, _p, _p]),
    'gpemsr_igemm': (_i, [_p, _p]),
    'gpemsr_act_pack_nchw': (_i, [_p, _i, _p, _i, _p, _p, _p, _p]),
    'gpemsr_act_unpack_nchw': (_i, [_p, _i, _p, _i, _p, _p]),
    'gpemsr_deform_im2col': (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p]),
    'gpemsr_resize_bilinear': (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _i, _p, _p, _p, _i, _p, _p, _p]),
    'gpemsr_avg_pool2': (_i, [_p, _i64, _i, _i, _p, _p]),
    'gpemsr_pack_concat3': (_i, [_p, _i, _p, _i, _p, _i, _p, _p, _p, _p]),
    'gpemsr_conv3x3_c1_relu': (_i, [_p, _p, _p, _i, _p, _p, _p, _p]),
    'gpemsr_patch_cosine': (_i, [_p, _i64, _f, _p, _p]),
    'gpemsr_space_to_depth': (_i, [_p, _p, _p, _i, _p, _p, _p, _p]),
    'gpemsr_pack_weights': (_i, [_p, _i, _i, _i64, _i64, _i, _p, _i, _i, _p, _p, _p]),
    'gpemsr_igemm_plan': (_i, [_p, _p, _p]),
    'gpemsr_pack_weights_tiled_bytes': (_sz, [_i, _i, _i, _i, _i]),
    'gpemsr_pack_weights_tiled': (_i, [_p, _i, _i, _i64, _i64, _i, _p, _i, _i, _i, _p, _p]),
    'gpemsr_gn_stats': (_i, [_p, _i, _p, _p, _p]),
    'gpemsr_gn_scale_shift': (_i, [_p, _i, _p, _p, _i, _i, _i, C.c_double, _f, _p, _p]),
    'gpemsr_affine_act': (_i, [_p, _i, _p, _p, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    'gpemsr_softmax_cells_blocked': (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    'gpemsr_border_phase_conv': (_i, [_p, _i, _p, _p, _p, _i, _p, _p]),
    'gpemsr_add_bilinear_base': (_i, [_p, _i, _i, _i, _i, _p, _p]),
    'gpemsr_cells_upsample2x': (_i, [_p, _p, _i, _f, _p, _i, _p, _p, _p, _p]),
    'gpemsr_cells_mul_mask': (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    'gpemsr_cells_pool3x3s2': (_i, [_p, _p, _i, _p, _p, _p, _p]),
    'gpemsr_cells_copy': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p]),
    'gpemsr_temporal_attn_scale': (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    'gpemsr_threeda_combine': (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    'gpemsr_tap_gather_sum': (_i, [_p, _p, _i, _i, _i, _p, _i, _f, _p, _i, _i, _i, _p, _p]),
    'gpemsr_conv3x3_direct': (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p]),
    'gpemsr_selftest_gemm_workspace_bytes': (_sz, [_i64, _i, _i]),
    'gpemsr_selftest_gemm': (_i, [_p, _p, _i64, _i, _i, _i, _i, _p, _p, _sz, _p]),
    'gpemsr_selftest_gemm_status': (_i, [_p, _i64, _i, _i, _p]),
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpemsrError(-6, f'{LIB_PATH} is missing: build it with `python -m gpemsr_b200.build` '
                                  '(there is no CPU or PyTorch fallback)')
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name, None)
            if fn is None:      # stale build: calling the symbol raises AttributeError; tests/test_capi.py checks the full list
                continue
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise GpemsrError(rc, lib().gpemsr_last_error_string().decode())


def ptr(t):
    """Device pointer of a tensor for the C ABI.  The kernels take raw pointers, so a host tensor (a model that was never moved
    to the GPU, a checkpoint loaded with map_location='cpu') or a strided view would fault inside the launch: refuse here."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    if not t.is_cuda:
        raise GpemsrError(-3, f'expected a CUDA tensor, got one on {t.device} (move the module and its inputs to the GPU: '
                              'there is no CPU fallback)')
    if not t.is_contiguous():
        raise GpemsrError(-2, f'expected a contiguous tensor, got strides {tuple(t.stride())} for shape {tuple(t.shape)}')
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def kernel_launches():
    return int(lib().gpemsr_kernel_launches())
