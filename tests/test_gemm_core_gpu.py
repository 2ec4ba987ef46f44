"""GPU: the tcgen05 GEMM core (descriptors, TMEM, bulk-copy pipeline) against fp64 matmul."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a, b, split, block_n):
    from gpemsr_b200 import _lib
    L = _lib.lib()
    m, k = a.shape
    n = b.shape[0]
    wsb = L.gpemsr_selftest_gemm_workspace_bytes(m, n, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    d = torch.full((m, n), float('nan'), device='cuda')
    _lib.check(L.gpemsr_selftest_gemm(_lib.ptr(a), _lib.ptr(b), m, n, k, split, block_n, _lib.ptr(d), _lib.ptr(ws), wsb,
                                      _lib.stream_ptr()))
    _lib.check(L.gpemsr_selftest_gemm_status(_lib.ptr(ws), m, n, k, _lib.stream_ptr()))
    return d


@pytest.mark.parametrize('shape', [(128, 64, 64), (300, 200, 100), (1000, 256, 512), (4096, 1024, 512), (77, 40, 24)])
@pytest.mark.parametrize('block_n', [64, 128, 256])
def test_single_pass_bf16(shape, block_n, cuda_dev):
    m, n, k = shape
    g = torch.Generator(device='cuda').manual_seed(m + n + k)
    a = torch.randn(m, k, device='cuda', generator=g)
    b = torch.randn(n, k, device='cuda', generator=g)
    d = _gemm(a, b, 1, block_n)
    want = a.bfloat16().double() @ b.bfloat16().double().t()
    err = (d.double() - want).abs().max().item()
    assert err < 1e-3 * (k ** 0.5), err


@pytest.mark.parametrize('shape', [(128, 64, 64), (300, 200, 100), (1000, 256, 512), (2048, 512, 1152)])
@pytest.mark.parametrize('block_n', [64, 256])
def test_split3_fp32_faithful(shape, block_n, cuda_dev):
    m, n, k = shape
    g = torch.Generator(device='cuda').manual_seed(m * 3 + n + k)
    a = torch.randn(m, k, device='cuda', generator=g)
    b = torch.randn(n, k, device='cuda', generator=g)
    d = _gemm(a, b, 3, block_n)
    want = a.double() @ b.double().t()
    bound = (a.abs().double() @ b.abs().double().t()) * 2.0 ** -15 + 1e-6
    assert bool(((d.double() - want).abs() <= bound).all()), ((d.double() - want).abs() / bound).max().item()


def test_integer_exact(cuda_dev):
    g = torch.Generator(device='cuda').manual_seed(9)
    a = torch.randint(-4, 5, (515, 136), device='cuda', generator=g).float()
    b = torch.randint(-4, 5, (333, 136), device='cuda', generator=g).float()
    for split in (1, 3):
        d = _gemm(a, b, split, 128)
        assert torch.equal(d, a @ b.t())
