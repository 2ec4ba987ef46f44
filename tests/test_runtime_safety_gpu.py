"""Host-runtime safety on the GPU: packed weights follow the live parameters, a pipeline time-out raises (and is reset)
instead of returning invalid slices, host tensors are refused before a launch."""
import pytest
import torch

from full_model_util import build

pytestmark = pytest.mark.gpu


def test_packed_weights_follow_reloaded_parameters(cuda_dev):
    """The reference nn.Module always reads its live parameters: a load_state_dict / in-place update after the first forward
    (warm-up then load, a checkpoint swap) must change the result here too."""
    from oracle import weights as W
    model, sd = build(8, seed=301, device=cuda_dev)
    x = torch.rand(1, 5, 1, 16, 16, generator=torch.Generator().manual_seed(302)).cuda()
    out_a, ref_a = model(x)
    sd_b = W.fill_state({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=303)
    model.load_state_dict(sd_b, strict=True)                       # in place: same storages, new values
    out_b, ref_b = model(x)
    fresh, _ = build(8, seed=303, device=cuda_dev)
    out_f, ref_f = fresh(x)
    assert float((out_b - out_f).abs().max()) <= 1e-4 and float((ref_b - ref_f).abs().max()) <= 1e-4
    assert float((out_b - out_a).abs().max()) > 1e-3               # ... and it really is a different model
    # an in-place edit of one parameter is picked up as well
    with torch.no_grad():
        model.conv_last.bias.add_(0.25)
    out_c, _ = model(x)
    assert abs(float((out_c - out_b).mean()) - 0.25) <= 1e-4
    # moving the module (new storages) too
    model.cpu(); model.cuda()
    out_d, _ = model(x)
    assert float((out_d - out_c).abs().max()) <= 1e-4


def test_pipeline_timeout_raises_and_resets(cuda_dev):
    import gpemsr_b200
    from gpemsr_b200 import igemm as G
    model, _ = build(8, seed=304, device=cuda_dev)
    x = torch.rand(1, 5, 1, 16, 16, generator=torch.Generator().manual_seed(305)).cuda()
    good, _ = model(x)
    model.check()
    G.err_flag(x.device).fill_(3)                                  # what a bounded mbarrier wait leaves behind
    with pytest.raises(gpemsr_b200.GpemsrError):
        model.strict_errors = True
        model(x)                                                   # kernels drain early; the call itself reports it
    model.strict_errors = False
    assert int(G.err_flag(x.device).item()) == 0                   # reset: later launches are not poisoned
    again, _ = model(x)
    model.check()
    assert float((again - good).abs().max()) <= 1e-4
    # deferred mode: the NEXT call (or check()) raises
    G.err_flag(x.device).fill_(4)
    model(x)
    torch.cuda.synchronize()
    with pytest.raises(gpemsr_b200.GpemsrError):
        model(x)
    model(x)
    model.check()


def test_host_tensors_are_refused(cuda_dev):
    import gpemsr_b200
    model, _ = build(8, seed=306)                                  # never moved to the GPU
    with pytest.raises(gpemsr_b200.GpemsrError):
        model(torch.rand(1, 5, 1, 16, 16, device='cuda'))
    torch.cuda.synchronize()                                       # no launch happened: the context is healthy
    assert float(torch.ones(4, device='cuda').sum()) == 4.0
