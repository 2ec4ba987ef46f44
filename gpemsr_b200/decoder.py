"""Host-side mirror of the reference VQ-GAN ``Decoder`` (model/decoder.py) and its blocks (model/blocks.py)
on the sm_100a implicit-GEMM kernels.

The module tree, constructor arguments and parameter names are those of the reference
(``input_layer.{0..}``, ``feat_extract.{i}.block.{0,1,3,4}``, ``feat_extract.{i}.upblock``, ``output_layer`` ...),
so reference checkpoints load with ``strict=True``; the nn.Modules below only HOLD parameters -- ``forward`` and
``multi_scale_feat_calculate`` (same signatures and return values as model/decoder.py:37-57) run the CUDA path.
Inference only.  ``precision='fp32'`` (default) runs every GEMM as a 3-term bf16 split (fp32-faithful);
``precision='bf16'`` runs single bf16 passes.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import igemm as G


def Normalize(in_channels):                                   # model/blocks.py:5-6
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class ResidualBlock(nn.Module):                               # parameter holder for model/blocks.py:8-23
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.block = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, 1, 1), Normalize(out_channels), nn.ReLU(inplace=True),
                                   nn.Conv2d(out_channels, out_channels, 3, 1, 1), Normalize(out_channels), nn.ReLU(inplace=True))
        if in_channels != out_channels:
            self.channel_up = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class UpBlock(nn.Module):                                     # model/blocks.py:32-38
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.upblock = nn.ConvTranspose2d(in_channels, out_channels, 3, 2, 1, 1)


class DownBlock(nn.Module):                                   # model/blocks.py:41-47
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.downblock = nn.Conv2d(in_channels, out_channels, 3, 2, 1)


class NonLocalBlock(nn.Module):                               # model/blocks.py:50-59
    def __init__(self, channels):
        super().__init__()
        self.in_channels = channels
        self.gn = Normalize(channels)
        self.q = nn.Conv2d(channels, channels, 1, 1, 0)
        self.k = nn.Conv2d(channels, channels, 1, 1, 0)
        self.v = nn.Conv2d(channels, channels, 1, 1, 0)
        self.proj_out = nn.Conv2d(channels, channels, 1, 1, 0)


class _Plan:
    """Packed weights + activation buffers for one input shape (built lazily, reused across calls)."""

    def __init__(self, dec, n, h, w, device):
        self.prec = G.Precision(dec.precision)
        self.split = self.prec.planes()          # activations carry lo planes whenever some launch is fp32-faithful
        self.device = device
        self.err = G.err_flag(device)           # one flag per device, read back once per forward (igemm.post_error_check)
        self.n, self.h, self.w = n, h, w
        self.bufs = {}
        self.wts = {}
        self.gn = {}

    def act(self, name, geom, c, f32, planes=True):
        key = name
        a = self.bufs.get(key)
        if a is None:
            a = G.Act(geom, c, self.device, f32=f32, planes=planes, split=self.split)
            self.bufs[key] = a
        return a

    def sp(self, name):
        """The split (1 | 3) layer `name` runs at (igemm.Precision)."""
        return self.prec.split(name)

    def weights(self, name, param, kind, taps=None, min_rows=0, split=None):
        """Packed B operand of `param`, re-packed whenever the live parameter changes (igemm.cached)."""
        return G.cached(self.wts, name, (param,),
                        lambda: G.Weights(param, kind, taps=taps, split=split or self.sp(name), min_rows=min_rows))

    def derived(self, name, params, build):
        """Anything computed on the host from parameters (merged phases, composed layers, Kronecker weights), same caching."""
        return G.cached(self.wts, name, tuple(params), build)

    def scratch(self, n, c):
        key = (n, c)
        s = self.gn.get(key)
        if s is None:
            s = G.GroupNormScratch(n, c, self.device)
            self.gn[key] = s
        return s


class _BlockNet(nn.Module):
    """Runs the reference's building blocks (model/blocks.py) on the implicit-GEMM kernels; shared by Decoder and Indexer*."""

    def _init_runner(self, precision):
        self.precision = G.Precision(precision)
        self._plans = {}

    def _plan_for(self, x):
        if not x.is_cuda:
            from ._lib import GpemsrError
            raise GpemsrError(-3, f'{type(self).__name__} needs CUDA tensors: there is no CPU fallback')
        n, _, h, w = x.shape
        key = (n, h, w, x.device.index)
        P = self._plans.get(key)
        if P is None:
            P = _Plan(self, n, h, w, x.device)
            self._plans[key] = P
        self._last_plan = P
        return P

    def check(self):
        """Synchronise and raise if a GEMM pipeline timed out."""
        G.check_pipeline(self._last_plan.err)

    # ------------------------------------------------------------------ building blocks
    def _conv(self, P, name, mod, x, out, **kw):
        wt = P.weights(name, mod.weight, 'conv')
        G.igemm(x, wt, P.err, split=P.sp(name), bias=mod.bias.detach(), out=out, **kw)

    def _res_block(self, P, name, rb, x, nchw_out=None, planes_into=None):
        """x + ReLU(GN(conv(ReLU(GN(conv(x))))))  (model/blocks.py:25-29); x carries fp32 master + planes.
        planes_into = (Act, c_off): the block's operand planes are written straight into that channel slot of a wider buffer
        of the same geometry (the consumer's concat operand) and the returned activation reads them from there."""
        g, c = x.geom, rb.out_channels
        raw = P.act(f'raw{g.key()}_{c}', g, c, f32=True, planes=False)
        h = P.act(f'h{g.key()}_{c}', g, c, f32=False)
        y = P.act(name + '.out', g, c, f32=True)
        sc = P.scratch(g.n, c)
        # GroupNorm statistics ride in the conv epilogue when a group spans whole 8-channel cells (c >= 256); for narrow
        # layers the per-group shuffles and atomics cost more than the separate reduction pass (measured), so they keep it
        cpg = c // 32
        fused = cpg >= 8
        st_kw = dict(gn_sums=sc.sums, gn_cpg=cpg) if fused else {}
        self._conv(P, name + '.block.0', rb.block[0], x, raw, out_planes=False, **st_kw)
        G.group_norm_act(raw, rb.block[1].weight.detach(), rb.block[1].bias.detach(), sc, h, act=G.ACT_RELU, out_f32=False,
                         fused_stats=fused)
        self._conv(P, name + '.block.3', rb.block[3], h, raw, out_planes=False, **st_kw)
        if rb.in_channels != rb.out_channels:
            short = P.act(name + '.short', g, c, f32=True, planes=False)
            self._conv(P, name + '.channel_up', rb.channel_up, x, short, out_planes=False)
            res = short.f32
        else:
            res = x.f32
        if planes_into is not None:
            dst, c_off = planes_into
            out = _View(dst.hi[c_off // 8:], None if dst.lo is None else dst.lo[c_off // 8:], g)
            out.f32, out.c = y.f32, c
            y = out
        G.group_norm_act(raw, rb.block[4].weight.detach(), rb.block[4].bias.detach(), sc, y, act=G.ACT_RELU, residual=res,
                         fused_stats=fused, out_nchw=nchw_out)
        return y

    def _up_block(self, P, name, ub, x, need_f32):
        """ConvTranspose2d(k3, s2, p1, op1) as four parity-phase GEMMs (model/blocks.py:32-38)."""
        g = x.geom
        og = G.Geom(g.n, 2 * g.h, 2 * g.w, True)
        cout = ub.upblock.weight.shape[1]
        y = P.act(name + '.out', og, cout, f32=need_f32)
        if cout % 32 == 0 and cout <= 64:
            # narrow up-blocks are bound by memory traffic: run the four phases as ONE GEMM with 4*cout columns
            def build():
                wt = G.Weights(G.convT_merged_weight(ub.upblock.weight.detach()), 'conv', taps='offsets01', split=P.sp(name),
                               flop_scale=9 / 16)
                wt.bias4 = ub.upblock.bias.detach().repeat(4).contiguous()
                return wt
            wt = P.derived(name + '.merged', (ub.upblock.weight, ub.upblock.bias), build)
            G.igemm(x, wt, P.err, split=P.sp(name), bias=wt.bias4, out=y, up=2, phase_cols=cout, out_f32=need_f32)
            return y
        for py in (0, 1):
            for px in (0, 1):
                wt = P.weights(f'{name}.p{py}{px}', ub.upblock.weight, 'convT', taps=G.convT_phase_taps(py, px))
                G.igemm(x, wt, P.err, split=P.sp(name), bias=ub.upblock.bias.detach(), out=y, up=2, py=py, px=px,
                        out_f32=need_f32)
        return y

    def _down_block(self, P, name, db, x):
        """Conv2d(k3, s2, p1) (model/blocks.py:41-47) = space-to-depth + a 2x2-tap stride-1 GEMM over 4*cin channels."""
        g = x.geom
        conv = db.downblock
        og = G.Geom(g.n, (g.h + 1) // 2, (g.w + 1) // 2, True)
        s2d = P.act(name + '.s2d', og, 4 * conv.in_channels, f32=False)
        G.space_to_depth(x, s2d)
        def build():
            m, taps = G.down_conv_weight(conv.weight.detach())
            return G.Weights(m, 'conv', taps=taps, split=P.sp(name), flop_scale=9 / 16)
        wt = P.derived(name, (conv.weight,), build)
        y = P.act(name + '.out', og, conv.out_channels, f32=True)
        G.igemm(s2d, wt, P.err, split=P.sp(name), bias=conv.bias.detach(), out=y)
        return y

    def _non_local(self, P, name, nl, x):
        """x + proj_out(softmax(q^T k / sqrt(c)) applied to v)  (model/blocks.py:61-83)."""
        g, c = x.geom, nl.in_channels
        t = g.h * g.w
        t_pad = G._round_up(t, 128)
        cg = G.Geom(g.n, g.h, g.w, padded=False, r_img=t_pad)          # compact token rows
        hn = P.act(name + '.h', cg, c, f32=False)
        sc = P.scratch(g.n, c)
        G.group_norm_act(x, nl.gn.weight.detach(), nl.gn.bias.detach(), sc, hn, act=G.ACT_NONE, out_f32=False)
        q = P.act(name + '.q', cg, c, f32=False)
        k = P.act(name + '.k', cg, c, f32=False)
        o = P.act(name + '.o', cg, c, f32=False)
        G.igemm(hn, P.weights(name + '.q', nl.q.weight, 'conv'), P.err, split=P.sp(name + '.q'), bias=nl.q.bias.detach(), out=q, out_f32=False)
        G.igemm(hn, P.weights(name + '.k', nl.k.weight, 'conv'), P.err, split=P.sp(name + '.k'), bias=nl.k.bias.detach(), out=k, out_f32=False)
        # v^T [tokens/8][channels][8]: weights as the A operand, tokens as columns, bias per row
        wv = P.weights(name + '.v', nl.v.weight, 'conv', min_rows=128)      # used as the A operand: >= one 128-row tile
        c_rows = G._round_up(wv.b_rows, 128)
        assert c_rows == wv.b_rows
        wgeom = G.Geom(1, 1, c, padded=False, r_img=c_rows, m0=0, rows_alloc=c_rows)
        wa = _View(wv.hi[0], wv.lo[0] if wv.lo is not None else None, wgeom)
        vt = P.bufs.get(name + '.vt')
        if vt is None:
            shape = (g.n, t_pad // 8, c_rows, 8)
            vt = (torch.zeros(shape, dtype=torch.bfloat16, device=P.device),
                  torch.zeros(shape, dtype=torch.bfloat16, device=P.device) if P.split == 3 else None)
            P.bufs[name + '.vt'] = vt
            # the softmax is fused into the two attention GEMMs: no fp32 score matrix.  rmax / rsum: per-query statistics;
            # p: the unnormalised probabilities exp(s - rmax) as (hi, lo) operand planes [t_pad/8][t_pad][8] (zero-initialised:
            # padding rows / key columns are never written), shared by the images of the batch (stream order)
            P.bufs[name + '.rmax'] = torch.zeros(t_pad, dtype=torch.float32, device=P.device)
            P.bufs[name + '.rsum'] = torch.zeros(t_pad, dtype=torch.float32, device=P.device)
            pshape = (t_pad // 8, t_pad, 8)
            P.bufs[name + '.p'] = (torch.zeros(pshape, dtype=torch.bfloat16, device=P.device),
                                   torch.zeros(pshape, dtype=torch.bfloat16, device=P.device) if P.split == 3 else None)
        rmax, rsum = P.bufs[name + '.rmax'], P.bufs[name + '.rsum']
        p_hi, p_lo = P.bufs[name + '.p']
        pg = G.Geom(1, g.h, g.w, padded=False, r_img=t_pad, m0=0, rows_alloc=t_pad)
        pa = _View(p_hi, p_lo, pg)
        row_bytes = 16                                                # one 8 x bf16 cell
        sm_scale = float(int(c) ** (-0.5))
        for i in range(g.n):
            vt_geom = G.Geom(1, 1, c, padded=False, r_img=wgeom.r_img, m0=0, rows_alloc=c_rows)
            vt_i = _View(vt[0][i], vt[1][i] if vt[1] is not None else None, vt_geom)
            G.igemm(wa, None, P.err, split=P.sp(name + '.v'), bias=nl.v.bias.detach(), bias_per_row=True, out=vt_i, out_f32=False,
                    n_cols=t, b_hi=hn.hi.data_ptr() + i * t_pad * row_bytes,
                    b_lo=(hn.lo.data_ptr() + i * t_pad * row_bytes) if hn.lo is not None else None,
                    b_rows=cg.rows_alloc, k_pad=hn.c_pad)
            k_hi = k.hi.data_ptr() + i * t_pad * row_bytes
            k_lo = (k.lo.data_ptr() + i * t_pad * row_bytes) if k.lo is not None else None
            # 1. row maxima of q_i^T k_i / sqrt(c) from ONE bf16 pass (a stabiliser only: any value near the maximum will do)
            G.igemm(q, None, P.err, split=1, scale=sm_scale, a_geom=cg.sample(i), n_cols=t, b_hi=k_hi, b_rows=cg.rows_alloc,
                    k_pad=k.c_pad, row_max_out=rmax)
            # 2. scores in the fp32-faithful split; the epilogue stores exp(s - max) as operand planes and sums the rows
            G.igemm(q, None, P.err, split=P.sp(name + '.scores'), scale=sm_scale, a_geom=cg.sample(i), n_cols=t, b_hi=k_hi, b_lo=k_lo,
                    b_rows=cg.rows_alloc, k_pad=k.c_pad, act=G.ACT_EXP, row_max=rmax, row_sum=rsum, out=pa, out_f32=False)
            # 3. o_i = (P v_i^T) / row sum
            G.igemm(pa, None, P.err, split=P.sp(name + '.pv'), n_cols=c, b_hi=vt_i.hi.data_ptr(),
                    b_lo=vt_i.lo.data_ptr() if vt_i.lo is not None else None, b_rows=c_rows, k_pad=t_pad,
                    out=o, o_geom=cg.sample(i), out_f32=False, row_div=rsum)
        y = P.act(name + '.out', g, c, f32=True)
        G.igemm(o, P.weights(name + '.proj_out', nl.proj_out.weight, 'conv'), P.err, split=P.sp(name + '.proj_out'),
                bias=nl.proj_out.bias.detach(), residual=x.f32, out=y)
        return y


class Decoder(_BlockNet):
    """Drop-in for ``model.decoder.Decoder`` (same ``args`` dict)."""

    def __init__(self, args, precision='fp32', compose_final=True):
        super().__init__()
        self.args = args
        self.compose_final = compose_final
        self.channel_list = args['channel_list']
        self.num_res_blocks = args['num_resblock_per_scale']
        self.num_input_resblck = args['num_input_resblck']
        self.latent_dim = args['latent_dim']
        self.use_non_local = args['use_non_local']
        self._init_runner(precision)

        layers = [nn.Conv2d(self.latent_dim, self.channel_list[0], 1)]
        for _ in range(self.num_input_resblck):
            layers.append(ResidualBlock(self.channel_list[0], self.channel_list[0]))
        self.input_layer = nn.Sequential(*layers)
        layers = []
        if self.use_non_local:
            layers.append(NonLocalBlock(self.channel_list[0]))
        for i in range(len(self.channel_list) - 1):
            cin, cout = self.channel_list[i], self.channel_list[i + 1]
            for _ in range(self.num_res_blocks):
                layers.append(ResidualBlock(cin, cin))
            layers.append(UpBlock(cin, cout))
        self.feat_extract = nn.Sequential(*layers)
        self.output_layer = nn.Conv2d(self.channel_list[-1], args['im_channel'], 3, 1, 1)

    def _final_stage(self, P, ub, x, img):
        """Last UpBlock + output conv (model/decoder.py:31,33) composed into ONE 4-phase GEMM on the up-block's input
        grid: the 64 x 16h x 16w intermediate (the largest tensor of the decoder, which the reference never returns) is not
        materialised.  The one-pixel ring of the image, where the conv's zero padding changes the weights, is then
        re-evaluated exactly with per-class weights."""
        co = self.output_layer.out_channels
        def build():
            wc, bias = G.compose_upblock_conv(ub.upblock.weight, ub.upblock.bias, self.output_layer.weight, self.output_layer.bias)
            dev = ub.upblock.weight.device
            w_int = wc[4].reshape(4 * co, wc.shape[3], 3, 3).contiguous().to(dev)          # phase-major output channels
            few = 4 * co <= 4            # nine taps as GEMM columns + a nine-point sum instead of nine N = 16 tensor-pipe tiles
            return dict(wt=G.Weights(G.taps_as_columns(w_int) if few else w_int, 'conv', split=P.sp('final')), few=few,
                        b_int=bias[4].repeat(4).contiguous().to(dev), wc=wc.contiguous().to(dev), bias=bias.contiguous().to(dev))
        st = P.derived('final', (ub.upblock.weight, ub.upblock.bias, self.output_layer.weight, self.output_layer.bias), build)
        if st['few']:
            key = f'final.taps{x.geom.key()}'
            taps = P.bufs.get(key)
            if taps is None:
                taps = P.bufs[key] = G.TapCells(x.geom, 4 * co, P.device)
            G.conv3x3_few_outputs(x, st['wt'], taps, P.err, P.sp('final'), st['b_int'], img, 4 * co, up=2, co=co)
        else:
            G.igemm(x, st['wt'], P.err, split=P.sp('final'), bias=st['b_int'], up=2, phase_cols=co, out_nchw=img, nchw_c=co)
        G.border_phase_conv(x, st['wc'], st['bias'], co, img)

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def multi_scale_feat_calculate(self, x):                  # model/decoder.py:40-57
        feats, img = self._run(x, want_feats=True)
        return feats + [img]

    @torch.no_grad()
    def forward(self, x):                                     # model/decoder.py:37-38
        return self._run(x, want_feats=False)[1]

    @torch.no_grad()
    def multi_scale_feat_into(self, x, sinks):
        """``multi_scale_feat_calculate`` for a caller that consumes the four features as convolution operands: sinks[i] =
        (Act, c_off) receives the operand planes of feature i (coarse to fine; None = not needed) directly from the block's
        last GroupNorm pass -- no NCHW copy, no re-pack.  Returns the decoded image (NCHW)."""
        return self._run(x, want_feats=False, sinks=sinks)[1]

    def _run(self, x, want_feats, sinks=None):
        P = self._plan_for(x)
        G.poll_error(x.device)
        with G.nested():
            res = self._run_body(P, x, want_feats, sinks)
        G.post_error_check(x.device)
        return res

    def _run_body(self, P, x, want_feats, sinks):
        n, c, h, w = x.shape
        g = G.Geom(n, h, w, True)
        xin = P.act('in', g, c, f32=False)
        G.pack_nchw(x.float(), xin)
        cur = P.act('input_layer.0.out', g, self.channel_list[0], f32=True)
        self._conv(P, 'input_layer.0', self.input_layer[0], xin, cur)
        for i in range(self.num_input_resblck):
            cur = self._res_block(P, f'input_layer.{i + 1}', self.input_layer[i + 1], cur)
        feats, n_feat = [], 0
        nlayers = len(self.feat_extract)
        for li, mod in enumerate(self.feat_extract):
            name = f'feat_extract.{li}'
            if isinstance(mod, NonLocalBlock):
                cur = self._non_local(P, name, mod, cur)
            elif isinstance(mod, ResidualBlock):
                nxt = self.feat_extract[li + 1] if li + 1 < nlayers else None
                nchw, sink = None, None
                if isinstance(nxt, UpBlock):                   # the tensors model/decoder.py:46/51 collects
                    if want_feats:
                        nchw = torch.empty(n, mod.out_channels, cur.geom.h, cur.geom.w, dtype=torch.float32, device=x.device)
                        feats.append(nchw)                     # written by the block's last GroupNorm pass, no extra copy
                    elif sinks is not None:
                        sink = sinks[n_feat]
                    n_feat += 1
                cur = self._res_block(P, name, mod, cur, nchw_out=nchw, planes_into=sink)
            else:
                last = li == nlayers - 1
                if last and self.compose_final and 4 * self.output_layer.out_channels <= 16:
                    img = torch.empty(n, self.output_layer.out_channels, 2 * cur.geom.h, 2 * cur.geom.w, dtype=torch.float32,
                                      device=x.device)
                    self._final_stage(P, mod, cur, img)
                    return feats, img
                cur = self._up_block(P, name, mod, cur, need_f32=not last)
        img = torch.empty(n, self.output_layer.out_channels, cur.geom.h, cur.geom.w, dtype=torch.float32, device=x.device)
        wt = P.weights('output_layer', self.output_layer.weight, 'conv')
        G.igemm(cur, wt, P.err, split=P.sp('output_layer'), bias=self.output_layer.bias.detach(), out_nchw=img,
                nchw_c=self.output_layer.out_channels)
        return feats, img


class _View:
    """Operand planes that are not a whole Act (weights used as the A operand, per-sample slices)."""

    def __init__(self, hi, lo, geom):
        self.hi, self.lo, self.geom, self.f32 = hi, lo, geom, None
