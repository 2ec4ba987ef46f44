"""Restatement of the reference's inference entry point ``output_GPEMSR.main()`` (output_GPEMSR.py:18-128) around a model callable.

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference does not exist on the GPU box, so the GPU test of the whole loop
(tests/test_entry_loop_gpu.py) drives THIS restatement; the real, unmodified script is run on the mirror in the authoring
container by tools/run_entry_point.py (tests/test_entry_point.py).  Followed line by line:

  * option file (:23-28): ``scale``, ``save_path``, ``dataset`` block, ``network`` block
  * ``CREMIDataset`` (:132-214): the centre frames are the HR directory's sorted integer stems minus N//2 at each end (:148-154);
    item i = the N LR frames around centre i, read by ``read_img`` (data/util.py:75-88: cv2.IMREAD_UNCHANGED, float32 / 255,
    HW -> HWC) and stacked to [N, C, H, W] (:205-209); a missing LR file falls back to the previous index (``seek_path`` :216-222)
  * windows (:54-128): two head windows built from item 0 ((0,0,0,1,2), (0,0,1,2,3)), the loader loop (batch 1), two tail windows
    built from the last item ((-4,-3,-2,-1,-1), (-3,-2,-1,-1,-1)); ``SR, _ = model(LQ)``
  * ``tensor2img`` (util/util.py:139-163): squeeze, clamp to [0, 1], * 255, round, uint8; ``cv2.imwrite`` to ``{k}.png``.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def read_img(path):                                             # data/util.py:75-88
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED).astype(np.float32) / 255.
    if img.ndim == 2:
        img = np.expand_dims(img, axis=2)
    return img[:, :, :3]


def seek_path(idx, dir_path, center):                           # output_GPEMSR.py:216-222
    cur = center + idx
    p = os.path.join(dir_path, str(cur) + '.png')
    return p if os.path.exists(p) else seek_path(idx - 1, dir_path, center)


def tensor2img(t):                                              # util/util.py:139-163 (uint8, min_max (0, 1), 2-D case)
    t = t.squeeze().float().cpu().clamp_(0, 1)
    assert t.dim() == 2
    return (t.numpy() * 255.0).round().astype(np.uint8)


class Dataset:                                                  # output_GPEMSR.py:132-214 (phase 'val')
    def __init__(self, gt_root, lq_root, n_frames):
        half = (n_frames - 1) // 2
        stems = sorted(int(n[:-4]) for n in os.listdir(gt_root))
        self.centers = stems[half:-half]
        self.lq_root, self.offsets = lq_root, list(range(-half, half + 1))

    def __len__(self):
        return len(self.centers)

    def __getitem__(self, i):
        i = i if i >= 0 else len(self) + i
        imgs = np.stack([read_img(seek_path(o, self.lq_root, self.centers[i])) for o in self.offsets], axis=0)
        return torch.from_numpy(np.ascontiguousarray(np.transpose(imgs, (0, 3, 1, 2))).copy()).float()


def run(model, gt_root, lq_root, save_path, n_frames=5, device='cpu'):
    """The slice loop of output_GPEMSR.py:49-128; returns the list of uint8 images it wrote (0.png, 1.png, ...)."""
    import cv2
    os.makedirs(save_path, exist_ok=True)
    ds = Dataset(gt_root, lq_root, n_frames)
    assert n_frames == 5                                        # the head / tail windows below are written out for 5 frames
    first, last = ds[0], ds[-1]
    pick = lambda item, idx: torch.stack([item[j] for j in idx]).unsqueeze(0)
    windows = [pick(first, (0, 0, 0, 1, 2)), pick(first, (0, 0, 1, 2, 3))]          # :54-82
    windows += [ds[i].unsqueeze(0) for i in range(len(ds))]                        # :84-96 (DataLoader, batch 1)
    windows += [pick(last, (-4, -3, -2, -1, -1)), pick(last, (-3, -2, -1, -1, -1))]  # :98-128
    imgs = []
    with torch.no_grad():
        for k, lq in enumerate(windows):
            sr, _ = model(lq.to(device))
            img = tensor2img(sr)
            cv2.imwrite(os.path.join(save_path, f'{k}.png'), img)
            imgs.append(img)
    return imgs
