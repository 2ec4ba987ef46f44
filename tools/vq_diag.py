"""Diagnostic: candidate-list statistics and timing of the fused logits lookup on real Indexer features vs N(0,1) features."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gpemsr_b200
from gpemsr_b200 import codebook as CB


def stats(tag, feat, w, b, emb):
    for _ in range(3):
        CB.logits_argmax_gather(feat, w, b, emb)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        CB.logits_argmax_gather(feat, w, b, emb)
    e.record(); torch.cuda.synchronize()
    ws = list(CB._ws_cache.values())[0]
    B, D, H, W = feat.shape
    rows = B * H * W
    rows_pad = (rows + 127) // 128 * 128
    d_pad, k_pad = (D + 3 + 63) // 64 * 64, 1024
    r256 = lambda n: (n + 255) // 256 * 256
    off = r256(rows_pad * d_pad * 2) + r256(k_pad * d_pad * 2) + r256(k_pad * 4) + r256(rows_pad * 4)
    margin = ws[off:off + rows * 4].view(torch.float32)
    off += r256(rows_pad * 4)
    cnt = ws[off:off + rows * 8].view(torch.int32).view(rows, 2)
    over = (cnt == -1).any(1)
    c = cnt.clone(); c[c < 0] = 17
    print(tag, 'ms/call', s.elapsed_time(e) / 10, 'rows', rows, 'overflow rows', int(over.sum()), 'mean appended/half', float(c.float().mean()),
          'max', int(c.max()), 'margin mean', float(margin.mean()), 'row norm', float(feat.permute(0, 2, 3, 1).reshape(-1, D).norm(dim=1).mean()))


def main():
    dev = torch.device('cuda')
    wts = bench.make_weights()
    hp = bench.NativeHotPath(wts, dev)
    ins = {k: v.to(dev) for k, v in bench.make_inputs(bench.LR, bench.NFRAMES, seed=100).items()}
    feat = hp.idx.features(ins['lr_frames']).clone()
    w, b = hp.idx.embedding.weight.detach(), hp.idx.embedding.bias.detach()
    emb = hp.cb.embedding.weight.detach()
    stats('indexer feat', feat, w, b, emb)
    stats('randn feat  ', torch.randn_like(feat), w, b, emb)
    stats('randn*30    ', torch.randn_like(feat) * 1.3, w, b, emb)


if __name__ == '__main__':
    main()
