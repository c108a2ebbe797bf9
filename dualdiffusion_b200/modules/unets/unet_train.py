"""Train-step schedule of the B200 EDM2 UNet: the activation-saving forward and the hand-scheduled backward
that replace `loss.backward()` through the reference's module graph
(/root/reference/src/training/trainer.py:1022-1044 -> modules/unets/unet_edm2_b4.py:250-296, Block.forward
:110-158, MPConv.forward modules/mp_tools.py:357-373).

Nothing here is PyTorch autograd arithmetic: `UNetFunction` is one autograd node whose forward/backward are
fixed schedules of C-ABI launches (include/dualdiffusion_b200.h, "backward pass of the UNet train step").
Per MPConv the backward is
    dgrad  = dd_mpconv_forward on the transposed / tap-reversed effective weights (dd_weight_transpose),
    wgrad  = dd_mpconv_wgrad (tcgen05, pixel dimension as K) -> dL/dW_eff,
    dL/dW  = dd_weight_prep_bwd (weight-norm projection + gain), batched over all parameters of a bucket,
and the block glue (pixel-norm, mp_silu, mp_sum, mp_cat, clip, resample, attention, embedding heads) is the
dd_*_bwd kernels.  Parameter gradients land in one flat fp32 buffer laid out in backward-completion order so a
bucket can be all-reduced (NCCL, `dualdiffusion_b200.ddp.GradAllReducer`) while earlier layers are still
running their backward.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from ... import _lib as L
from ... import ops
from ..mp_tools import mp_cat_weights

Tensor = torch.Tensor

_CONV_OUT_PAD = 32      # conv_out rows (4) padded to the dgrad kernel's 32-channel input granule


def _mp_sum_coeffs(t: float) -> Tuple[float, float]:
    n = math.sqrt((1 - t) ** 2 + t ** 2)
    return (1 - t) / n, t / n


class _ParamSlot:
    """One trainable tensor: where its effective-weight gradient and its parameter gradient live."""
    __slots__ = ("name", "param", "O", "I_g", "taps", "row_stride", "rows_eff", "normalize", "perm", "head_dim",
                 "gain", "eff_off", "grad_off", "dweff", "grad", "t_cout_g")

    def __init__(self, name, param, *, row_stride=0, rows_eff=0, perm=0, head_dim=0, gain=None, normalize=True):
        # Grouped layers with more input than output channels per group (conv_res1: I_g = 2 * cout_g): the weight
        # gradient is computed with the operands exchanged (M = input channels), so that a 128-row MMA tile straddles
        # half as many groups of the block diagonal (twice the useful work per issued MMA); the gradient then lands
        # transposed and tap-reversed, [groups*I_g][taps][cout_g], which dd_weight_prep_bwd reads through t_cout_g.
        groups = getattr(param, "conv_groups", 1)
        cout_g = param.shape[0] // groups
        self.t_cout_g = cout_g if (groups > 1 and param.ndim == 4 and param.shape[1] > cout_g and cout_g < 128
                                   and cout_g % 32 == 0) else 0
        self.name, self.param = name, param
        self.O, self.I_g = param.shape[0], param.shape[1]
        taps = 1
        for s in param.shape[2:]:
            taps *= s
        self.taps = taps
        self.row_stride = row_stride or self.I_g * taps
        self.rows_eff = max(rows_eff, self.O)
        self.normalize, self.perm, self.head_dim, self.gain = normalize, perm, head_dim, gain


class TrainState:
    """Persistent device state of the train step for one UNet: transposed weights for dgrad, the flat
    effective-weight-gradient and parameter-gradient buffers, and the descriptor tables of the batched kernels."""

    def __init__(self, net, plan) -> None:
        from .unet_edm2_b4 import Block
        self.net, self.plan = net, plan
        dev = plan.device
        for n, p in net.named_parameters():
            if p.dtype != torch.float32:
                raise RuntimeError(f"dualdiffusion_b200 UNet training needs fp32 master parameters ({n} is {p.dtype}); "
                                   "the reference trains fp32 parameters under bf16 autocast (trainer.py:114)")
        self.wt: Dict[str, Tensor] = {}          # dgrad operands
        self.conv_out_w32: Optional[Tensor] = None
        self._prep_descs = self._trans_descs = self._prep_entries = None
        self._prep_sig = None
        self.step_graphs: Dict[tuple, _StepGraphs] = {}

        # ---- gradient buckets in backward-completion order ----
        buckets: List[List[_ParamSlot]] = []
        slots: Dict[str, _ParamSlot] = {}

        def conv_slots(p: str, blk) -> List[_ParamSlot]:
            """Every weight of one block: its convolutions and its embedding projections.  The latter's gradients are
            complete as soon as the block's backward has run, so they travel in the block's bucket; left to the tail
            bucket their ~200 MB were all-reduced after the backward had ended, fully exposed."""
            out = [_ParamSlot(p + ".conv_res0.weight", blk.conv_res0.weight),
                   _ParamSlot(p + ".conv_res1.weight", blk.conv_res1.weight),
                   _ParamSlot(p + ".conv_skip.weight", blk.conv_skip.weight),
                   _ParamSlot(p + ".emb_linear.weight", blk.emb_linear.weight, gain=blk.emb_gain)]
            if blk.use_attention:
                out += [_ParamSlot(p + ".attn_qk.weight", blk.attn_qk.weight, perm=L.WPERM_QK,
                                   head_dim=blk.channels_per_head),
                        _ParamSlot(p + ".attn_v.weight", blk.attn_v.weight),
                        _ParamSlot(p + ".attn_proj.weight", blk.attn_proj.weight),
                        _ParamSlot(p + ".emb_linear_qk.weight", blk.emb_linear_qk.weight, gain=blk.emb_gain_qk),
                        _ParamSlot(p + ".emb_linear_v.weight", blk.emb_linear_v.weight, gain=blk.emb_gain_v)]
            return out

        cur: List[_ParamSlot] = [_ParamSlot("conv_out.weight", net.conv_out.weight, rows_eff=_CONV_OUT_PAD,
                                            gain=net.out_gain)]
        cur_bytes = 0
        # >= 64 MB of fp32 gradients per bucket: still at NVLink bandwidth, and the buckets that finish last (whose
        # all-reduce is what stays exposed) are smaller.  DD_DDP_BUCKET_MB: tuning experiments.
        target = int(os.environ.get("DD_DDP_BUCKET_MB", "64")) << 20
        for prefix, blocks in (("dec", net.dec), ("enc", net.enc)):
            for name, blk in reversed(list(blocks.items())):
                if not isinstance(blk, Block):
                    continue
                for s in conv_slots(f"{prefix}.{name}", blk):
                    cur.append(s)
                    cur_bytes += s.param.numel() * 4
                if cur_bytes >= target:
                    buckets.append(cur)
                    cur, cur_bytes = [], 0
        # tail bucket: the first encoder blocks (whatever did not fill a bucket), the stem and the noise embedding
        cur.append(_ParamSlot("enc.conv_in.weight", net.enc["conv_in"].weight, row_stride=64))
        self.block_order: List[Tuple[str, object]] = []
        for prefix, blocks in (("enc", net.enc), ("dec", net.dec)):
            for name, blk in blocks.items():
                if isinstance(blk, Block):
                    self.block_order.append((f"{prefix}.{name}", blk))
        cur.append(_ParamSlot("emb_noise.weight", net.emb_noise.weight))
        buckets.append(cur)
        self.buckets = buckets

        eff_total = grad_total = 0
        for b in buckets:
            for s in b:
                s.eff_off, s.grad_off = eff_total, grad_total
                eff_total += s.rows_eff * s.row_stride
                grad_total += s.param.numel()
                slots[s.name] = s
        self.slots = slots
        self.n_gains = len(plan.gain_params)
        self.gain_off = grad_total
        grad_total += self.n_gains
        self.dweff_flat = torch.zeros(eff_total, device=dev, dtype=torch.float32)
        self.grad_flat = torch.zeros(grad_total, device=dev, dtype=torch.float32)
        self.bucket_ranges: List[Tuple[int, int]] = []
        for i, b in enumerate(buckets):
            lo = b[0].grad_off
            hi = b[-1].grad_off + b[-1].param.numel()
            if i == len(buckets) - 1:
                hi = grad_total            # scalar gains ride in the tail bucket
            self.bucket_ranges.append((lo, hi))
        for s in slots.values():
            s.dweff = self.dweff_flat[s.eff_off:s.eff_off + s.rows_eff * s.row_stride].view(s.rows_eff, s.row_stride)
            s.grad = self.grad_flat[s.grad_off:s.grad_off + s.param.numel()].view_as(s.param)
        self.dgains = self.grad_flat[self.gain_off:self.gain_off + self.n_gains]

        # ---- weight-prep backward descriptor tables (one per bucket, overwrite / accumulate variants) ----
        self.wbwd: List[Dict[bool, Tuple[Tensor, int, int]]] = []
        for b in buckets:
            variants = {}
            for acc in (False, True):
                entries = []
                for s in b:
                    gi = None if s.gain is None else plan.gain_index[id(s.gain)]
                    entries.append(dict(w=s.param.detach(), dweff=s.dweff, dw=s.grad,
                                        gain=None if gi is None else plan.gains_f32[gi:gi + 1],
                                        dgain=None if gi is None else self.dgains[gi:gi + 1], gain_host=1.0,
                                        O=s.O, I_g=s.I_g, taps=s.taps, normalize=s.normalize, perm=s.perm,
                                        head_dim=s.head_dim, row_stride=s.row_stride, accumulate=acc,
                                        t_cout_g=s.t_cout_g))
                buf, rows = ops.make_wbwd_descs(entries, dev)
                variants[acc] = (buf, len(entries), rows)
            self.wbwd.append(variants)

        # ---- embedding projections backward (batched over blocks) ----
        self.affine_bwd: Dict[int, dict] = {}

    def refresh(self) -> None:
        """Train-mode weight preparation (mp_tools.py:359-364: normalise, scale by gain/sqrt(fan_in), cast) for every
        MPConv of the network plus the transposed dgrad operands: two batched launches, re-run whenever a parameter
        version changed (i.e. after every optimizer step)."""
        from .unet_edm2_b4 import _ver
        plan = self.plan
        plan.refresh_gains()
        items = plan._weights()
        if self._prep_descs is None and not all(w.is_contiguous() for _, w, *_ in items):
            raise RuntimeError("dualdiffusion_b200 UNet: parameters must be in the default contiguous (OIHW) layout; the "
                               "kernels read them through raw pointers (supports_channels_last is False for this reason)")
        sig = (sum(_ver(w) for _, w, *_ in items), plan.gain_version)
        if self._prep_descs is None:
            dev = plan.device
            prep, trans = [], []
            for key, w, gain, qk_dim, pad_rows, row_stride in items:
                O, I_g = w.shape[0], w.shape[1]
                taps = w.shape[2] * w.shape[3]
                groups = getattr(w, "conv_groups", 1)
                if key == "conv_out":
                    self.conv_out_w32 = torch.zeros((_CONV_OUT_PAD, taps * I_g), device=dev, dtype=torch.bfloat16)
                    out = self.conv_out_w32
                else:
                    out = plan.prepped.get(key)
                    if out is None:
                        out = (torch.zeros((max(O, pad_rows), row_stride or taps * I_g), device=dev, dtype=torch.bfloat16)
                               if (pad_rows or row_stride) else torch.empty((O, taps, I_g), device=dev, dtype=torch.bfloat16))
                        plan.prepped[key] = out
                prep.append(dict(w=w.detach(), out=out, gain=None if gain is None else plan.gain_ptr(gain), O=O, I_g=I_g,
                                 taps=taps, normalize=True, perm=L.WPERM_QK if qk_dim else L.WPERM_NONE, head_dim=qk_dim,
                                 row_stride=row_stride))
                if key == "enc.conv_in":
                    continue                          # no gradient flows to the network input
                rows = _CONV_OUT_PAD if key == "conv_out" else O
                self.wt[key] = torch.empty((groups * I_g, taps, rows // groups), device=dev, dtype=torch.bfloat16)
                trans.append(dict(src=out, dst=self.wt[key], cout_g=rows // groups, cin_g=I_g, taps=taps, groups=groups))
            self._prep_entries = (prep, trans)        # keeps every tensor the descriptors point at alive
            self._prep_descs = ops.make_wprep_descs(prep, dev) + (len(prep),)
            self._trans_descs = ops.make_wtrans_descs(trans, dev) + (len(trans),)
        if sig == self._prep_sig and plan.training is True:
            return
        buf, rows, n = self._prep_descs
        ops.weight_prep_batched(buf, n, rows)
        buf, tiles, n = self._trans_descs
        ops.weight_transpose_batched(buf, n, tiles)
        self._prep_sig = sig
        plan.training = True                          # eval-mode refresh_weights() re-prepares (un-normalised) on mode flip
        plan.versions.clear()

    def affine_bwd_for(self, B: int, st_fwd: dict) -> dict:
        st = self.affine_bwd.get(B)
        if st is not None and st["fwd"] is st_fwd:
            return st
        dev = self.plan.device
        total_o = sum(e["w"].shape[0] for e in st_fwd["entries"])
        dc_flat = torch.zeros(B * total_o, device=dev, dtype=torch.float32)
        rowscale = torch.empty(total_o, device=dev, dtype=torch.float32)
        douts: Dict[str, Tensor] = {}
        entries = []
        off = 0
        keys = list(st_fwd["outs"].keys())
        for key, e in zip(keys, st_fwd["entries"]):
            conv = e["conv"]
            O, I = conv.weight.shape[0], conv.weight.shape[1]
            dout = dc_flat[B * off:B * (off + O)].view(B, O)
            douts[key] = dout
            pfx, tag = key.rsplit(".", 1)
            pname = pfx + {"c": ".emb_linear.weight", "c_qk": ".emb_linear_qk.weight", "c_v": ".emb_linear_v.weight"}[tag]
            slot = self.slots[pname]
            entries.append(dict(w=conv.weight.detach().view(O, I), gain=e["gain"], dout=dout, dweff=slot.dweff,
                                rowscale=rowscale[off:off + O], groups=conv.groups, normalize=True))
            off += O
        descs, max_o, max_cols = ops.make_affine_bwd_descs(entries, dev)
        # descriptors of the blocks of each gradient bucket: a contiguous range of the (forward-ordered) table, because a
        # bucket is a contiguous run of the backward order and the last encoder blocks precede the first decoder blocks
        bucket_of = {s.name: i for i, b in enumerate(self.buckets) for s in b}
        ranges: List[Optional[Tuple[int, int]]] = [None] * len(self.buckets)
        for j, key in enumerate(keys):
            pfx, tag = key.rsplit(".", 1)
            i = bucket_of[pfx + {"c": ".emb_linear.weight", "c_qk": ".emb_linear_qk.weight", "c_v": ".emb_linear_v.weight"}[tag]]
            lo, n = ranges[i] if ranges[i] is not None else (j, 0)
            if lo + n != j:
                raise RuntimeError("embedding descriptors of a gradient bucket are not contiguous")
            ranges[i] = (lo, n + 1)
        st = dict(fwd=st_fwd, dc_flat=dc_flat, douts=douts, descs=descs, n=len(entries), max_o=max_o, max_cols=max_cols,
                  rowscale=rowscale, entries=entries, ranges=ranges)
        self.affine_bwd[B] = st
        return st


def get_train_state(net, plan) -> TrainState:
    ts = getattr(plan, "train_state", None)
    if ts is None:
        ts = TrainState(net, plan)
        plan.train_state = ts
    return ts


# ---------------------------------------------------------------------------------------------------------
# forward (train mode): the inference schedule plus the tensors the backward needs
# ---------------------------------------------------------------------------------------------------------
def train_forward(net, plan, x_in: Tensor, net_in: Tensor, sigma: Tensor, embeddings: Tensor, ln_freqs: Tensor,
                  x_ref: Optional[Tensor]) -> Tuple[Tensor, dict]:
    from .unet_edm2_b4 import Block
    cfg = net.config
    B = x_in.shape[0]
    W = plan.prepped
    aux = net._aux()
    g = cfg.mlp_groups
    ca_r, cb_r = _mp_sum_coeffs(cfg.res_balance)
    ca_a, cb_a = _mp_sum_coeffs(cfg.attn_balance)

    emb = ops.noise_embedding(sigma, aux["emb_freqs"], aux["emb_phases"], net.emb_noise.weight.detach(), embeddings,
                              cfg.label_balance, normalize=True)
    st = plan.affine_for(B)
    descs, max_o = plan.affine_descs(st)
    ops.emb_affine(descs, len(st["entries"]), max_o, emb)
    # the projections are overwritten by the next forward: the backward needs this call's values
    cvec = {k: v.clone() for k, v in st["outs"].items()}

    patches = ops.stem_patches(net_in, sigma, cfg.sigma_data, ln_freqs)
    x = ops.mpconv(patches, W["enc.conv_in"], 1)
    saved = dict(B=B, sigma=sigma, embeddings=embeddings, emb=emb, cvec=cvec, patches=patches, x_ref=x_ref, blocks={},
                 affine_fwd=st)
    skips = [x]

    def residual_branch(p: str, blk, s: Tensor, sv: dict, resid: Tensor) -> Tensor:
        y0, pre = ops.mpconv(s, W[p + ".conv_res0"], 3, g, epi=L.EPI_SCALE_SILU, scale=cvec[p + ".c"], epi2=L.EPI2_RAW)
        sv.update(s=s, pre=pre, y0=y0)
        if not blk.use_attention:
            return ops.mpconv(y0, W[p + ".conv_res1"], 3, g, epi=L.EPI_RESIDUAL, alpha=cb_r, beta=ca_r, clip=blk.clip_act,
                              residual=resid)
        x2, xs = ops.mpconv(y0, W[p + ".conv_res1"], 3, g, epi=L.EPI_RESIDUAL, alpha=cb_r, beta=ca_r, residual=resid,
                            epi2=L.EPI2_SCALE, scale2=cvec[p + ".c_qk"])
        v = ops.mpconv(x2, W[p + ".attn_v"], 1)
        qk = ops.mpconv(xs, W[p + ".attn_qk"], 1)
        y, a_raw = ops.attention_train(qk, v, cvec[p + ".c_v"], blk.num_heads, blk.channels_per_head)
        sv.update(x2=x2, xs=xs, v=v, qk=qk, y=y, a_raw=a_raw)
        return ops.mpconv(y, W[p + ".attn_proj"], 1, epi=L.EPI_RESIDUAL, alpha=cb_a, beta=ca_a, clip=blk.clip_act,
                          residual=x2)

    for name, blk in net.enc.items():
        if not isinstance(blk, Block):
            continue
        p = "enc." + name
        sv = dict(down=blk.resample_mode == "down", in_shape=tuple(x.shape))
        if sv["down"]:
            x = ops.avgpool2(x)
        t0 = ops.mpconv(x, W[p + ".conv_skip"], 1)
        xn, s = ops.pixnorm_silu(t0)
        sv.update(x_in=x, t0=t0)
        x = residual_branch(p, blk, s, sv, xn)
        sv["out"] = x
        saved["blocks"][p] = sv
        skips.append(x)

    for name, blk in net.dec.items():
        p = "dec." + name
        sv = dict(a_prev=x, skip_idx=None, up=False, wa=1.0, wb=0.0, Ca=x.shape[-1], Cb=0)
        if "layer" in name:
            skip = skips.pop()
            sv["skip_idx"] = len(skips)
            wa, wb = mp_cat_weights(x.shape[-1], skip.shape[-1], cfg.concat_balance)
            sv.update(wa=wa, wb=wb, Cb=skip.shape[-1])
            xc, s = ops.cat_silu(x, skip, wa, wb, False)
        elif blk.resample_mode == "up":
            sv["up"] = True
            xc, s = ops.cat_silu(x, None, 1.0, 0.0, True)
        else:
            xc = x
            _, s = ops.cat_silu(x, None, 1.0, 0.0, False, need_cat=False)
        t0 = ops.mpconv(xc, W[p + ".conv_skip"], 1)
        sv["xc"] = xc
        x = residual_branch(p, blk, s, sv, t0)
        sv["out"] = x
        saved["blocks"][p] = sv

    saved["x_last"] = x
    saved["n_skips"] = len(net.enc)
    d = ops.conv_out(x, get_train_state(net, plan).conv_out_w32, x_in, sigma, cfg.sigma_data, x_ref)
    return d, saved


# ---------------------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------------------
def train_backward(net, plan, saved: dict, dD: Tensor, accumulate: bool = False, bucket_done=None) -> Tensor:
    """Runs the backward schedule; parameter gradients are written to (or, with `accumulate`, added into)
    `TrainState.grad_flat`.  `bucket_done(i)` is called right after bucket i's gradients were enqueued.
    Returns dL/d(embeddings) [B, cemb] fp32."""
    from .unet_edm2_b4 import Block
    ts = get_train_state(net, plan)
    cfg = net.config
    WT = ts.wt
    slots = ts.slots
    g8 = cfg.mlp_groups
    B = saved["B"]
    ca_r, cb_r = _mp_sum_coeffs(cfg.res_balance)
    ca_a, cb_a = _mp_sum_coeffs(cfg.attn_balance)
    cvec = saved["cvec"]
    ab = ts.affine_bwd_for(B, saved["affine_fwd"])
    ab["dc_flat"].zero_()
    dcv = ab["douts"]
    if not accumulate:
        ts.dgains.zero_()

    def wgrad(key: str, x: Tensor, dy: Tensor, k: int, groups: int = 1, scale: float = 1.0) -> None:
        s = slots[key + ".weight"]
        if s.t_cout_g:      # operands exchanged: [Cin][taps][cout_g], taps reversed (see _ParamSlot)
            ops.mpconv_wgrad(dy, x, k, groups, scale, out=s.dweff.view(groups * s.I_g, s.taps * s.t_cout_g))
        else:
            ops.mpconv_wgrad(x, dy, k, groups, scale, out=s.dweff)

    slot_bucket = {}
    for i, b in enumerate(ts.buckets):
        for s in b:
            slot_bucket[s.name] = i
    pending = [len(b) for b in ts.buckets]
    finished = set()

    def finish(names: List[str]) -> None:
        """Mark parameters complete; when a bucket fills, project its gradients (weight-norm backward)."""
        for n in names:
            if n in finished:
                continue
            finished.add(n)
            i = slot_bucket[n]
            pending[i] -= 1
            if pending[i] == 0:
                if ab["ranges"][i] is not None and i != len(ts.buckets) - 1:      # weight gradients of the bucket's embedding
                    # projections (the tail bucket's ran before the embedding-vector gradient below, which needs them)
                    lo, cnt_e = ab["ranges"][i]
                    ops.emb_affine_bwd(ab["descs"], cnt_e, ab["max_o"], ab["max_cols"], saved["emb"], None, first=lo)
                buf, cnt, rows = ts.wbwd[i][accumulate]
                ops.weight_prep_bwd(buf, cnt, rows)
                if bucket_done is not None:
                    bucket_done(i)

    def block_tail_bwd(p: str, blk, sv: dict, gout: Tensor) -> Tuple[Tensor, Tensor]:
        """From the (clip-masked) gradient of the block output to (g at the mp_sum input, ds): everything
        downstream of the residual add, i.e. the attention sub-path and conv_res1/conv_res0."""
        done = []
        if blk.use_attention:
            dy = ops.mpconv(gout, WT[p + ".attn_proj"], 1)
            wgrad(p + ".attn_proj", sv["y"], gout, 1, 1, cb_a)
            da = ops.silu_scale_bwd(dy, cb_a, sv["a_raw"], cvec[p + ".c_v"], dcv[p + ".c_v"])
            dqk, dv = ops.attention_bwd(sv["qk"], sv["v"], sv["a_raw"], da, blk.num_heads, blk.channels_per_head)
            dxs = ops.mpconv(dqk, WT[p + ".attn_qk"], 1)
            wgrad(p + ".attn_qk", sv["xs"], dqk, 1)
            dxv = ops.mpconv(dv, WT[p + ".attn_v"], 1)
            wgrad(p + ".attn_v", sv["x2"], dv, 1)
            gout = ops.attn_in_bwd(gout, ca_a, dxv, dxs, sv["x2"], cvec[p + ".c_qk"], dcv[p + ".c_qk"])
            done += [p + ".attn_proj.weight", p + ".attn_qk.weight", p + ".attn_v.weight",
                     p + ".emb_linear_qk.weight", p + ".emb_linear_v.weight"]     # their dc rows are complete
        dy0 = ops.mpconv(gout, WT[p + ".conv_res1"], 3, g8)
        wgrad(p + ".conv_res1", sv["y0"], gout, 3, g8, cb_r)
        dpre = ops.silu_scale_bwd(dy0, cb_r, sv["pre"], cvec[p + ".c"], dcv[p + ".c"])
        ds = ops.mpconv(dpre, WT[p + ".conv_res0"], 3, g8)
        wgrad(p + ".conv_res0", sv["s"], dpre, 3, g8)
        done += [p + ".conv_res1.weight", p + ".conv_res0.weight", p + ".emb_linear.weight"]
        sv["_done"] = done
        return gout, ds

    # ---- head ----
    x_last = saved["x_last"]
    dF = ops.head_grad(dD, saved["sigma"], cfg.sigma_data, saved["x_ref"], _CONV_OUT_PAD)
    dx = ops.mpconv(dF, WT["conv_out"], 3)
    wgrad("conv_out", x_last, dF, 3)
    finish(["conv_out.weight"])
    dec_blocks = list(net.dec.items())
    clip_last = dec_blocks[-1][1].clip_act
    gcur = ops.enc_grad_combine(dx, False, None, x_last, clip_last, tuple(x_last.shape))

    # ---- decoder, last block first ----
    dskips: Dict[int, Tensor] = {}
    for name, blk in reversed(dec_blocks):
        p = "dec." + name
        sv = saved["blocks"][p]
        gmid, ds = block_tail_bwd(p, blk, sv, gcur)
        d_xc = ops.mpconv(gmid, WT[p + ".conv_skip"], 1)
        wgrad(p + ".conv_skip", sv["xc"], gmid, 1, 1, ca_r)
        gcur, db = ops.cat_silu_bwd(d_xc, ca_r, ds, sv["xc"], sv["a_prev"], blk.clip_act, sv["wa"], sv["wb"], sv["up"],
                                    sv["Ca"], sv["Cb"])
        if sv["skip_idx"] is not None:
            dskips[sv["skip_idx"]] = db
        finish(sv["_done"] + [p + ".conv_skip.weight"])

    # ---- encoder, last block first; gcur is the gradient that reached the last encoder output through in0 ----
    enc_blocks = [(n, b) for n, b in net.enc.items() if isinstance(b, Block)]
    k = len(enc_blocks)
    last_sv = saved["blocks"]["enc." + enc_blocks[-1][0]]
    gcur = ops.enc_grad_combine(gcur, False, dskips.get(k), last_sv["out"], enc_blocks[-1][1].clip_act,
                                tuple(last_sv["out"].shape))
    for idx in range(k - 1, -1, -1):
        name, blk = enc_blocks[idx]
        p = "enc." + name
        sv = saved["blocks"][p]
        gmid, ds = block_tail_bwd(p, blk, sv, gcur)
        dt0 = ops.pixnorm_silu_bwd(gmid, ca_r, ds, sv["t0"])
        dx_in = ops.mpconv(dt0, WT[p + ".conv_skip"], 1)
        wgrad(p + ".conv_skip", sv["x_in"], dt0, 1)
        finish(sv["_done"] + [p + ".conv_skip.weight"])
        if idx > 0:
            prev = saved["blocks"]["enc." + enc_blocks[idx - 1][0]]["out"]
            gcur = ops.enc_grad_combine(dx_in, sv["down"], dskips.get(idx), prev, enc_blocks[idx - 1][1].clip_act,
                                        tuple(prev.shape))
        else:
            gcur = ops.enc_grad_combine(dx_in, sv["down"], dskips.get(0), None, 0.0, sv["in_shape"])

    # ---- stem ----
    wgrad("enc.conv_in", saved["patches"], gcur, 1)

    # ---- embedding side ----
    emb = saved["emb"]
    demb = torch.zeros_like(emb)
    if ab["ranges"][-1] is not None:            # stage 1 of the blocks left in the tail bucket (the first encoder blocks)
        lo, cnt_e = ab["ranges"][-1]
        ops.emb_affine_bwd(ab["descs"], cnt_e, ab["max_o"], ab["max_cols"], emb, None, first=lo)
    ops.emb_affine_bwd(ab["descs"], ab["n"], ab["max_o"], ab["max_cols"], None, demb)      # stage 2: every row scale exists now
    aux = net._aux()
    s_noise = slots["emb_noise.weight"]
    _, dlabel = ops.noise_embedding_bwd(saved["sigma"], aux["emb_freqs"], aux["emb_phases"], net.emb_noise.weight.detach(),
                                        saved["embeddings"], cfg.label_balance, demb, True, dweff=s_noise.dweff)
    finish([s.name for s in ts.buckets[-1]])
    return dlabel


# ---------------------------------------------------------------------------------------------------------
# autograd nodes (the boundary: torch sees three opaque differentiable functions)
# ---------------------------------------------------------------------------------------------------------
class _StepGraphs:
    """CUDA graphs of one train-step shape: the forward schedule, and the backward schedule cut into one graph per
    gradient bucket so the NCCL all-reduce of a finished bucket can be enqueued between two replays.  All graphs of a
    step share one memory pool and are replayed in capture order (forward, then the backward segments)."""

    def __init__(self) -> None:
        self.calls = 0                  # the first call of a shape runs eagerly (lazy descriptor tables, allocator)
        self.pool = None
        self.static: Optional[dict] = None
        self.fwd = None
        self.d: Optional[Tensor] = None
        self.saved: Optional[dict] = None
        self.dD: Optional[Tensor] = None
        self.bwd: Dict[bool, tuple] = {}     # accumulate -> ([(graph, bucket index or None)], dlabel, launches)
        self.fwd_launches = 0
        self.busy = False               # a forward whose backward has not run yet owns the saved activations


def _capture(fn, pool, device):
    """Capture fn() into a new graph on a side stream; returns (graph, fn's result)."""
    g = torch.cuda.CUDAGraph()
    cur = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        g.capture_begin(pool=pool)
        out = fn()
        g.capture_end()
    cur.wait_stream(side)
    return g, out


class UNetFunction(torch.autograd.Function):
    """D = UNet(x_in, sigma, embeddings[, x_ref, perturbed_input]); differentiable in `embeddings` and the
    parameters.  Parameter order: TrainState.slots order, then the scalar gains (plan.gain_params)."""

    @staticmethod
    def forward(ctx, net, x32, net_in, sg, em, lf, xr, embeddings_in, *params):
        plan = net._get_plan()
        ts = get_train_state(net, plan)
        ts.refresh()
        ctx.net = net
        ctx.emb_dtype = embeddings_in.dtype
        ctx.graphs = None
        key = (tuple(x32.shape), xr is not None, net_in is not x32, lf.data_ptr())
        sg_ = ts.step_graphs.setdefault(key, _StepGraphs()) if net.use_cuda_graphs else None
        if sg_ is None or sg_.busy or sg_.calls == 0:
            if sg_ is not None:
                sg_.calls += 1
            d, ctx.saved = train_forward(net, plan, x32, net_in, sg, em, lf, xr)
            return d
        if sg_.fwd is None:
            sg_.pool = torch.cuda.graph_pool_handle()
            sg_.static = dict(x=x32.clone(), n=net_in.clone() if net_in is not x32 else None, s=sg.clone(), e=em.clone(),
                              r=None if xr is None else xr.clone())
            st = sg_.static
            before = ops.launch_count
            sg_.fwd, (sg_.d, sg_.saved) = _capture(
                lambda: train_forward(net, plan, st["x"], st["n"] if st["n"] is not None else st["x"], st["s"], st["e"], lf,
                                      st["r"]), sg_.pool, plan.device)
            sg_.fwd_launches = ops.launch_count - before
            ops.launch_count = before                 # captured, not executed: replays are counted below
        st = sg_.static
        st["x"].copy_(x32)
        if st["n"] is not None:
            st["n"].copy_(net_in)
        st["s"].copy_(sg)
        st["e"].copy_(em)
        if st["r"] is not None:
            st["r"].copy_(xr)
        sg_.fwd.replay()
        ops.launch_count += sg_.fwd_launches
        sg_.busy = True
        ctx.graphs = sg_
        ctx.saved = sg_.saved
        return sg_.d.clone()

    @staticmethod
    def backward(ctx, dD):
        net, saved, sg_ = ctx.net, ctx.saved, ctx.graphs
        ctx.saved = None
        plan = net._get_plan()
        ts = get_train_state(net, plan)
        sync = getattr(net, "grad_sync", None)
        dD = dD.detach().to(torch.float32).contiguous()

        def run(accumulate: bool, bucket_done) -> Tensor:
            if sg_ is None:
                return train_backward(net, plan, saved, dD, accumulate, bucket_done)
            if sg_.dD is None:
                sg_.dD = torch.empty_like(dD)
            sg_.dD.copy_(dD)
            if accumulate not in sg_.bwd:
                # capture: every bucket boundary closes the current graph and opens the next one
                segs: list = []
                before = ops.launch_count
                cur = torch.cuda.current_stream(plan.device)
                side = torch.cuda.Stream(device=plan.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    state = {"g": torch.cuda.CUDAGraph()}
                    state["g"].capture_begin(pool=sg_.pool)

                    def cut(i: int) -> None:
                        state["g"].capture_end()
                        segs.append((state["g"], i))
                        state["g"] = torch.cuda.CUDAGraph()
                        state["g"].capture_begin(pool=sg_.pool)

                    dl = train_backward(net, plan, saved, sg_.dD, accumulate, cut)
                    state["g"].capture_end()
                    segs.append((state["g"], None))
                cur.wait_stream(side)
                sg_.bwd[accumulate] = (segs, dl, ops.launch_count - before)
                ops.launch_count = before
            segs, dl, n_launch = sg_.bwd[accumulate]
            ops.launch_count += n_launch
            for g, i in segs:
                g.replay()
                if i is not None and bucket_done is not None:
                    bucket_done(i)
            sg_.busy = False
            return dl.clone()

        if sync is not None:
            dlabel = sync.run_backward(net, plan, ts, saved, dD, backward_fn=lambda n_, p_, s_, d_, accumulate=False,
                                       bucket_done=None: run(accumulate, bucket_done))
            grads = [None] * (len(ts.slots) + ts.n_gains)
        else:
            dlabel = run(False, None)
            flat = ts.grad_flat.clone()       # autograd owns what it is handed; grad_flat is reused next step
            grads = [flat[s.grad_off:s.grad_off + s.param.numel()].view_as(s.param) for s in ts.slots.values()]
            grads += [flat[ts.gain_off + i] for i in range(ts.n_gains)]
        return (None, None, None, None, None, None, None, dlabel.to(ctx.emb_dtype), *grads)


def unet_params(net, plan) -> List[Tensor]:
    ts = get_train_state(net, plan)
    return [s.param for s in ts.slots.values()] + list(plan.gain_params)


class LabelEmbeddingFunction(torch.autograd.Function):
    """UNet.get_embeddings (unet_edm2_b4.py:232-235), differentiable in emb_label / emb_label_unconditional."""

    @staticmethod
    def forward(ctx, e, mask, w_label, w_uncond):
        ctx.save_for_backward(e, mask, w_label, w_uncond)
        return ops.label_embedding(e, w_label.detach().contiguous(), w_uncond.detach().contiguous(), mask, normalize=True)

    @staticmethod
    def backward(ctx, dout):
        e, mask, w_label, w_uncond = ctx.saved_tensors
        dwl_eff, dwu_eff = ops.label_embedding_bwd(e, mask, dout.detach().float().contiguous())
        dwl = torch.empty_like(w_label, dtype=torch.float32)
        dwu = torch.empty_like(w_uncond, dtype=torch.float32)
        entries = [dict(w=w_label.detach(), dweff=dwl_eff, dw=dwl, O=w_label.shape[0], I_g=w_label.shape[1], taps=1,
                        normalize=True),
                   dict(w=w_uncond.detach(), dweff=dwu_eff, dw=dwu, O=w_uncond.shape[0], I_g=w_uncond.shape[1], taps=1,
                        normalize=True)]
        buf, rows = ops.make_wbwd_descs(entries, dout.device)
        ops.weight_prep_bwd(buf, len(entries), rows)
        return None, None, dwl, dwu


class SigmaLogvarFunction(torch.autograd.Function):
    """UNet.get_sigma_loss_logvar (unet_edm2_b4.py:237-238), differentiable in logvar_linear.weight."""

    @staticmethod
    def forward(ctx, s, freqs, phases, w):
        ctx.save_for_backward(s, freqs, phases)
        return ops.sigma_logvar(s, freqs, phases, w.detach().contiguous())

    @staticmethod
    def backward(ctx, dout):
        s, freqs, phases = ctx.saved_tensors
        dw = ops.sigma_logvar_bwd(s, freqs, phases, dout.detach().float().contiguous().flatten())
        return None, None, None, dw
