"""CPU: the host-side weight transforms the CUDA path relies on, each against the torch operator it re-expresses
(float64, exact up to rounding): stride-2 conv = space-to-depth + 2x2 taps, ConvTranspose2d = four merged parity phases,
Conv3d(k=1) over the frame axis = 1x1 conv with a Kronecker weight, reffusionconv's input-channel permutation."""
import torch
import torch.nn.functional as F

from gpemsr_b200.igemm import convT_merged_weight, down_conv_weight


def _space_to_depth(x):
    """gpemsr_space_to_depth: channel block (p*2 + q) of output pixel (y, x) = input pixel (2y + p, 2x + q), zero past an odd edge."""
    n, c, h, w = x.shape
    ho, wo = (h + 1) // 2, (w + 1) // 2
    xp = F.pad(x, (0, 2 * wo - w, 0, 2 * ho - h))
    return torch.cat([xp[:, :, p::2, q::2] for p in (0, 1) for q in (0, 1)], dim=1)


def test_down_conv_weight_equals_stride2_conv():
    torch.manual_seed(0)
    for ci, co, h, w in ((3, 5, 8, 10), (4, 2, 7, 9)):
        x = torch.randn(2, ci, h, w, dtype=torch.float64)
        wt = torch.randn(co, ci, 3, 3, dtype=torch.float64)
        m, taps = down_conv_weight(wt.float())
        m = m.double()
        s2d = _space_to_depth(x)
        ho, wo = s2d.shape[2], s2d.shape[3]
        sp = F.pad(s2d, (1, 1, 1, 1))
        out = torch.zeros(2, co, ho, wo, dtype=torch.float64)
        for src, dy, dx in taps:                                   # tap (a, b) reads the s2d pixel at offset (a - 1, b - 1)
            a, b = src // 2, src % 2
            out += torch.einsum('ok,nkyx->noyx', m[:, :, a, b], sp[:, :, 1 + dy:1 + dy + ho, 1 + dx:1 + dx + wo])
        want = F.conv2d(x, wt, None, 2, 1)
        assert out.shape == want.shape and (out - want).abs().max().item() < 1e-5


def test_convT_merged_weight_equals_conv_transpose():
    torch.manual_seed(1)
    ci, co, h, w = 4, 3, 5, 6
    x = torch.randn(1, ci, h, w, dtype=torch.float64)
    wt = torch.randn(ci, co, 3, 3, dtype=torch.float64)
    m = convT_merged_weight(wt.float()).double()                   # [4*co, ci, 2, 2]: phase p = py*2 + px, taps = input offsets {0,1}^2
    xp = F.pad(x, (0, 1, 0, 1))
    out = torch.zeros(1, co, 2 * h, 2 * w, dtype=torch.float64)
    for py in (0, 1):
        for px in (0, 1):
            p = py * 2 + px
            acc = torch.zeros(1, co, h, w, dtype=torch.float64)
            for dy in (0, 1):
                for dx in (0, 1):
                    acc += torch.einsum('ok,nkyx->noyx', m[p * co:(p + 1) * co, :, dy, dx], xp[:, :, dy:dy + h, dx:dx + w])
            out[:, :, py::2, px::2] = acc
    want = F.conv_transpose2d(x, wt, None, 2, 1, 1)
    assert (out - want).abs().max().item() < 1e-5


def test_conv3d_frame_mixing_is_a_kronecker_1x1_conv():
    """ThreeDA.conv3D_1/2 (model/GPEMSR.py:160-161, 204-207): Conv3d(t, t, k=1) on [b, t, c, h, w] == 1x1 Conv2d on [b, t*c, h, w]
    with weight kron(W, I_c) and bias repeated c times -- the form gpemsr_b200.GPEMSR runs on the tensor cores."""
    torch.manual_seed(2)
    b, t, c, h, w = 2, 5, 6, 4, 3
    al = torch.randn(b, t * c, h, w, dtype=torch.float64)
    w3, b3 = torch.randn(t, t, 1, 1, 1, dtype=torch.float64), torch.randn(t, dtype=torch.float64)
    want = F.conv3d(al.view(b, t, c, h, w), w3, b3).view(b, t * c, h, w)
    k = torch.kron(w3.reshape(t, t), torch.eye(c, dtype=torch.float64)).reshape(t * c, t * c, 1, 1)
    got = F.conv2d(al, k, b3.repeat_interleave(c))
    assert (got - want).abs().max().item() < 1e-12


def test_reffusion_channel_permutation():
    """gpemsr_b200.GPEMSR keeps one operand buffer per level laid out [R | carried | LR feat | decoder]; reffusionconv reads
    (carried, LR feat, decoder) with its input channels permuted from the reference order (LR feat, decoder, carried)."""
    torch.manual_seed(3)
    d, cc = 8, 12                                                  # decoder / carried widths of some level
    lr, dec, car = torch.randn(1, 64, 5, 5), torch.randn(1, d, 5, 5), torch.randn(1, cc, 5, 5)
    w = torch.randn(7, 64 + d + cc, 3, 3)
    want = F.conv2d(torch.cat((lr, dec, car), 1), w, None, 1, 1)
    wp = torch.cat([w[:, 64 + d:], w[:, :64], w[:, 64:64 + d]], dim=1)
    got = F.conv2d(torch.cat((car, lr, dec), 1), wp, None, 1, 1)
    assert (got - want).abs().max().item() < 1e-4


def test_cached_follows_parameter_identity_and_version():
    """igemm.cached: derived weights are rebuilt when the parameter is updated in place (load_state_dict) or replaced (.to())."""
    import torch
    from gpemsr_b200 import igemm as G
    p = torch.nn.Parameter(torch.zeros(4))
    cache, builds = {}, []
    get = lambda: G.cached(cache, 'w', (p,), lambda: builds.append(1) or float(p.sum()))
    assert get() == 0.0 and get() == 0.0 and len(builds) == 1
    with torch.no_grad():
        p.copy_(torch.ones(4))                                     # what load_state_dict does
    assert get() == 4.0 and len(builds) == 2
    p.data = torch.full((4,), 2.0)                                 # what .to(device) does
    assert get() == 8.0 and len(builds) == 3


def test_precision_table_longest_prefix_and_submodules():
    """igemm.Precision: 'fp32' / 'bf16' / {prefix: split}; longest prefix wins; sub() hands a sub-module its slice of the table."""
    import pytest
    from gpemsr_b200 import igemm as G
    assert G.Precision('fp32').split('anything') == 3 and G.Precision('bf16').split('x') == 1
    p = G.Precision({'vgg': 1, 'decoder': 3, 'decoder.feat_extract.0': 1, 'tail.up': 1, 'default': 3})
    assert p.split('vgg') == 1 and p.split('enc.conv_first') == 3
    assert p.split('decoder.feat_extract.0.q') == 1 and p.split('decoder.feat_extract.1.block.0') == 3
    assert p.split('tail.up3') == 1 and p.split('tail.hr') == 3 and p.planes() == 3
    d = p.sub('decoder')
    assert d.split('feat_extract.0.scores') == 1 and d.split('input_layer.0') == 3
    assert G.Precision({'default': 1}).planes() == 1
    with pytest.raises(ValueError):
        G.Precision({'vgg': 2})
    from gpemsr_b200.gpemsr import DEFAULT_PLAN
    q = G.Precision(DEFAULT_PLAN)
    assert q.split('vgg') == 1 and q.split('spynet') == 1 and q.split('indexer.feat_extract.0') == 3 and q.split('enc.mask.conv3') == 3
