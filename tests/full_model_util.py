"""Shared by the whole-model tests: the network blocks of option/output_GPEMSR_x{8,16}.yml restated (the tests may not read
/root/reference at run time) and the synthetic parameters of the golden fixtures."""
import torch

from oracle import weights as W


def argref(scale):
    """``network.argref`` of option/output_GPEMSR_x{8,16}.yml (the Encoder block is training-only and omitted)."""
    key = 'Indexer16' if scale == 16 else 'Indexer8'
    return {key: dict(channel_list=[64, 64, 128, 256, 512], im_channel=1, num_resblock_per_scale=2, num_output_resblck=3,
                      latent_dim=512, use_non_local=True),
            'Codebook': dict(num_codebook_vectors=1024, latent_dim=512, beta=1),
            'Decoder': dict(channel_list=[512, 256, 128, 64, 64], im_channel=1, num_resblock_per_scale=1, num_input_resblck=3,
                            latent_dim=512, use_non_local=True)}


def network_kwargs(scale, nframes=5):
    """``network`` block of the yml: nf 64, nframes 5, groups 8, front_RBs 5, back_RBs 10, ref_fusion_feat_RBs 1, POD, ThreeDA."""
    return dict(argref=argref(scale), nf=64, nframes=nframes, groups=8, front_RBs=5, back_RBs=10, w_ref=True, ref_fusion_feat_RBs=1,
                align_mode='POD', fusion_mode='ThreeDA', mode='16to1' if scale == 16 else '8to1', scale=scale)


def build(scale, seed=None, device=None, nframes=5, precision='plan'):
    """(mirror module with the fixture's synthetic parameters loaded, the same parameters as a CPU state dict).
    precision: 'plan' = the default precision plan (what a user gets), 'fp32' = every GEMM in the fp32-faithful split."""
    import gpemsr_b200
    m = gpemsr_b200.GPEMSR(None, None, precision=precision, **network_kwargs(scale, nframes))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = W.fill_state(shapes, seed=900 + scale if seed is None else seed)
    m.load_state_dict(sd, strict=True)
    if device is not None:
        m = m.to(device)
    return m.eval(), sd
