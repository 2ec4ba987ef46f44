"""Run the reference's REAL, unmodified entry point -- ``output_GPEMSR.main()`` (output_GPEMSR.py:18-128) -- on the mirror.

Authoring container only (needs /root/reference; there is no GPU here, so the C library is the argument-checking stand-in of
tools/_mock_lib.py and the written images carry no meaningful pixels).  What this proves is everything AROUND the kernels, with
the reference's own code driving it:

  yml (a copy of option/output_GPEMSR_x8.yml with the paths rewritten) -> the script's CREMIDataset over a synthetic PNG stack ->
  ``from model.GPEMSR import GPEMSR`` resolved by ``gpemsr_b200.dropin.install()`` -> the constructor call of :36-43 with
  ``ref_path_G`` / ``ref_path_Indexer`` pointing at synthetic stage-1 / stage-2 checkpoints -> ``.eval().to(device)`` ->
  ``load_state_dict(torch.load(stage3), strict=True)`` with the REFERENCE's full key set (incl. ``refmodel.encoder.*`` and
  ``vgg.slice2..5``, produced by the unmodified reference model) -> the 2 padded head windows, the loader loop, the 2 padded tail
  windows -> ``tensor2img`` -> ``cv2.imwrite``.

Checked: one forward per slice, every window equal to ``gpemsr_b200.volume.window_indices`` (what ``forward_volume`` and the
GPU test of the same loop use), the parameters that reached the mirror equal the checkpoint, one PNG per slice.
The numerical half (real kernels, PNGs against the oracle) is tests/test_entry_loop_gpu.py.

    python tools/run_entry_point.py [--ref /root/reference/GPEMSR-CREMI/GPEMSR] [--scale 8]
"""
import argparse
import os
import shutil
import sys
import tempfile

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.dirname(os.path.abspath(__file__))]


def reference_state(ref, scale, seed):
    """state_dict of the unmodified reference model (names / shapes incl. the dead sub-modules), live entries = fill_state."""
    from oracle import basicsr_shim
    from oracle import weights as W
    from gpemsr_b200.gpemsr import DEAD_PREFIXES
    with open(os.path.join(ref, 'option', f'output_GPEMSR_x{scale}.yml')) as f:
        net = yaml.safe_load(f)['network']
    gp = basicsr_shim.install(ref)
    kw = {k: net[k] for k in ('argref', 'nf', 'nframes', 'groups', 'front_RBs', 'back_RBs', 'w_ref', 'ref_fusion_feat_RBs',
                              'align_mode', 'fusion_mode', 'mode')}
    model = gp.GPEMSR(ref_path_G=None, ref_path_Indexer=None, scale=scale, **kw)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gp._oracle_restore()
    for m in [k for k in sys.modules if k == 'model' or k.startswith('model.')]:
        del sys.modules[m]                       # the drop-in registers its own model.GPEMSR below
    live = {k: tuple(v.shape) for k, v in sd.items() if not k.startswith(DEAD_PREFIXES)}
    sd.update(W.fill_state(live, seed=seed))
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference/GPEMSR-CREMI/GPEMSR')
    ap.add_argument('--scale', type=int, default=8)
    ap.add_argument('--slices', type=int, default=9)
    ap.add_argument('--lr', type=int, default=16)
    a = ap.parse_args()
    work = tempfile.mkdtemp(prefix='gpemsr_entry_')
    S, lr, s = a.slices, a.lr, a.scale
    try:
        # ---- synthetic data set: S LR slices + S HR slices (the script lists the HR directory to find its centre frames)
        import cv2
        rng = np.random.default_rng(5)
        d_lr, d_hr, d_out, d_ck = (os.path.join(work, p) for p in ('LR', 'HR', 'SR', 'ckpt'))
        for d in (d_lr, d_hr, d_ck):
            os.makedirs(d)
        vol = rng.integers(0, 256, (S, lr, lr), dtype=np.uint8)
        for i in range(S):
            cv2.imwrite(os.path.join(d_lr, f'{i}.png'), vol[i])
            cv2.imwrite(os.path.join(d_hr, f'{i}.png'), np.zeros((s * lr, s * lr), np.uint8))
        # ---- synthetic checkpoints under the reference's key names
        sd = reference_state(a.ref, s, seed=41)
        torch.save(sd, os.path.join(d_ck, 'stage3.pth'))
        torch.save({k[len('refmodel.'):]: v for k, v in sd.items() if k.startswith('refmodel.') and not k.startswith('refmodel.indexer.')},
                   os.path.join(d_ck, 'stage1.pth'))
        torch.save({k[len('refmodel.indexer.'):]: v for k, v in sd.items() if k.startswith('refmodel.indexer.')},
                   os.path.join(d_ck, 'stage2.pth'))
        # ---- the reference's yml with the paths rewritten
        with open(os.path.join(a.ref, 'option', f'output_GPEMSR_x{s}.yml')) as f:
            opt = yaml.safe_load(f)
        opt['save_path'], opt['pretrain_path'] = d_out, os.path.join(d_ck, 'stage3.pth')
        opt['dataset']['dataroot_GT'], opt['dataset']['dataroot_LQ'] = d_hr, d_lr
        opt['network']['ref_path_G'], opt['network']['ref_path_Indexer'] = os.path.join(d_ck, 'stage1.pth'), os.path.join(d_ck, 'stage2.pth')
        yml = os.path.join(work, 'opt.yml')
        with open(yml, 'w') as f:
            yaml.safe_dump(opt, f)

        # ---- no GPU here: argument-checking stand-in for the C library, `cuda` moves become no-ops
        import _mock_lib
        _mock_lib.install()
        real_mod_to, real_t_to = torch.nn.Module.to, torch.Tensor.to
        is_cuda = lambda x: (isinstance(x, torch.device) and x.type == 'cuda') or (isinstance(x, str) and x.startswith('cuda'))
        torch.nn.Module.to = lambda self, *aa, **kk: self if any(is_cuda(v) for v in aa) else real_mod_to(self, *aa, **kk)
        torch.Tensor.to = lambda self, *aa, **kk: self if any(is_cuda(v) for v in aa) else real_t_to(self, *aa, **kk)

        import gpemsr_b200
        from gpemsr_b200 import dropin
        from gpemsr_b200.volume import window_indices
        dropin.install()
        seen = []
        real_forward = gpemsr_b200.GPEMSR.forward

        def spy(self, x):
            seen.append((self, x.detach().clone()))
            return real_forward(self, x)
        gpemsr_b200.GPEMSR.forward = spy

        sys.path.insert(0, a.ref)
        os.chdir(a.ref)
        sys.argv = ['output_GPEMSR.py', '-opt', yml]
        import output_GPEMSR                                    # the reference's file, unmodified
        assert output_GPEMSR.GPEMSR is gpemsr_b200.GPEMSR        # `from model.GPEMSR import GPEMSR` took the mirror
        output_GPEMSR.main()

        # ---- checks
        assert len(seen) == S, (len(seen), S)
        lrf = torch.from_numpy(vol.astype(np.float32) / 255.0)
        for i, (_, x) in enumerate(seen):
            want = lrf[window_indices(i, S)].view(1, 5, 1, lr, lr)
            assert tuple(x.shape) == (1, 5, 1, lr, lr) and torch.equal(x, want), f'window {i} differs from the replicate-padded window'
        model = seen[0][0]
        got = model.state_dict()
        from gpemsr_b200.gpemsr import DEAD_PREFIXES
        live = [k for k in sd if not k.startswith(DEAD_PREFIXES)]
        assert sorted(got) == sorted(live), set(got) ^ set(live)
        assert all(torch.equal(got[k], sd[k]) for k in live)
        pngs = sorted(os.listdir(d_out), key=lambda n: int(n[:-4]))
        assert pngs == [f'{i}.png' for i in range(S)], pngs
        for n in pngs:
            im = cv2.imread(os.path.join(d_out, n), cv2.IMREAD_UNCHANGED)
            assert im.shape == (s * lr, s * lr) and im.dtype == np.uint8
        print(f'entry point ok: output_GPEMSR.main() ran unmodified on gpemsr_b200.GPEMSR (x{s}): {S} windows, {len(live)} live + '
              f'{len(sd) - len(live)} dropped checkpoint entries, {len(_mock_lib.calls)} C-ABI calls, {len(pngs)} PNGs')
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()
