"""CPU: the whole-model drop-in keeps the reference's construction / checkpoint-loading surface (output_GPEMSR.py:5,36-52)."""
import sys

import pytest
import torch

from full_model_util import network_kwargs


def test_install_registers_model_gpemsr():
    import gpemsr_b200
    from gpemsr_b200 import dropin
    saved = {k: sys.modules.get(k) for k in ('model', 'model.GPEMSR')}
    try:
        dropin.install()
        from model.GPEMSR import GPEMSR                      # the reference's import line
        assert GPEMSR is gpemsr_b200.GPEMSR
        kw = network_kwargs(16)
        m = GPEMSR(ref_path_G=None, ref_path_Indexer=None, argref=kw['argref'], nf=kw['nf'], nframes=kw['nframes'], groups=kw['groups'],
                   front_RBs=kw['front_RBs'], back_RBs=kw['back_RBs'], w_ref=kw['w_ref'], ref_fusion_feat_RBs=kw['ref_fusion_feat_RBs'],
                   align_mode=kw['align_mode'], fusion_mode=kw['fusion_mode'], mode=kw['mode'], scale=16)     # output_GPEMSR.py:36-43
        assert m.center == 2 and m.scale == 16
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_checkpoint_keys_load_strict():
    """A stage-3 checkpoint also holds refmodel.encoder.* (training only) and vgg.slice2..5 (never reach an output): they are
    dropped, everything else must match exactly (strict=True, output_GPEMSR.py:52)."""
    import gpemsr_b200
    m = gpemsr_b200.GPEMSR(None, None, **network_kwargs(8))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd['refmodel.encoder.input_layer.0.weight'] = torch.zeros(64, 1, 3, 3)
    sd['vgg.slice2.5.weight'] = torch.zeros(128, 64, 3, 3)
    m.load_state_dict(sd, strict=True)
    del sd['conv_first.weight']
    with pytest.raises(RuntimeError):
        m.load_state_dict(sd, strict=True)


def test_rejects_unsupported_configurations():
    import gpemsr_b200
    kw = network_kwargs(8)
    with pytest.raises(gpemsr_b200.GpemsrError):
        gpemsr_b200.GPEMSR(None, None, **dict(kw, align_mode='none'))
    with pytest.raises(ValueError):
        gpemsr_b200.GPEMSR(None, None, **dict(kw, scale=4, mode='4to1'))
