"""Tensor-level wrappers over the C ABI.  PyTorch is plumbing here (device memory + streams):
every function allocates its outputs with torch and enqueues exactly the C-ABI call(s) named in
include/dualdiffusion_b200.h on torch's current CUDA stream.  Activations are NHWC bf16 tensors
of logical shape [B, H, W, C]."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor

launch_count = 0   # number of C-ABI kernel launches issued (bench.py reports it as gpu_launches)


trace = None       # set to a list to record (op, detail) per launch (profiling scripts only)
timing = None      # set to a list to record (flops, start_event, end_event, detail) per MPConv launch (bench.py)


def _count(n: int = 1, what: str = "", detail=None) -> None:
    global launch_count
    launch_count += n
    if trace is not None:
        import sys
        trace.append((what or sys._getframe(1).f_code.co_name, detail))


def _is_bf16(t: Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise TypeError(f"weights must be fp32 or bf16, got {t.dtype}")


def weight_prep(w: Tensor, gain: Optional[Tensor] = None, gain_host: float = 1.0, normalize: bool = False,
                fmt: int = L.WFMT_BF16_OTI, qk_head_dim: int = 0, out: Optional[Tensor] = None,
                pad_rows: int = 0, row_stride: int = 0, qkv_head_dim: int = 0) -> Tensor:
    """MPConv weight path (reference modules/mp_tools.py:359-364) fused into one pass."""
    L.require_cuda(w)
    w = w.contiguous()
    O = w.shape[0]
    I_g = w.shape[1]
    taps = 1
    for s in w.shape[2:]:
        taps *= s
    if out is None:
        if pad_rows or row_stride:      # zero-initialised padded buffer [max(O,pad_rows)][row_stride or taps*I_g]
            out = torch.zeros((max(O, pad_rows), row_stride or taps * I_g), device=w.device,
                              dtype=torch.bfloat16 if fmt == L.WFMT_BF16_OTI else torch.float32)
        elif fmt == L.WFMT_BF16_OTI:
            out = torch.empty((O, taps, I_g), device=w.device, dtype=torch.bfloat16)
        else:
            out = torch.empty((O, I_g, taps), device=w.device, dtype=torch.float32)
    perm = L.WPERM_QK if qk_head_dim else (L.WPERM_QKV if qkv_head_dim else L.WPERM_NONE)
    L.check(L.load().dd_weight_prep(L.ptr(w), _is_bf16(w), L.ptr(out), fmt, O, I_g, taps, L.ptr(gain), gain_host,
                                    int(normalize), perm, qk_head_dim or qkv_head_dim, row_stride, L.stream_ptr()))
    _count()
    return out


def mpconv(x: Tensor, w_prepped: Tensor, ksize: int, groups: int = 1, *, epi: int = L.EPI_NONE,
           epi2: int = L.EPI2_NONE, alpha: float = 1.0, beta: float = 0.0, clip: float = 0.0,
           scale: Optional[Tensor] = None, scale2: Optional[Tensor] = None, residual: Optional[Tensor] = None,
           out: Optional[Tensor] = None, out2: Optional[Tensor] = None):
    """tcgen05 implicit-GEMM MPConv (reference modules/mp_tools.py:369) with fused block epilogue."""
    B, H, W, Cin = x.shape
    Cout = w_prepped.shape[0]
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.bfloat16)
    if epi2 != L.EPI2_NONE and out2 is None:
        out2 = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.bfloat16)
    e = L.ConvEpilogue(epi, epi2, alpha, beta, clip, L.ptr(scale), L.ptr(scale2), L.ptr(residual), L.ptr(out2))
    if timing is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    L.check(L.load().dd_mpconv_forward(L.ptr(x), L.ptr(w_prepped), L.ptr(out), B, H, W, Cin, Cout, ksize, groups,
                                       C.byref(e), L.stream_ptr()))
    if timing is not None:
        ev1.record()
        timing.append((2.0 * B * H * W * Cout * (Cin // groups) * ksize * ksize, ev0, ev1,
                       (B, H, W, Cin, Cout, ksize, groups, epi, epi2)))
    _count(1, "mpconv", (B, H, W, Cin, Cout, ksize, groups, epi, epi2))
    return (out, out2) if epi2 != L.EPI2_NONE else out


def mpconv_cat(x1: Tensor, x2: Tensor, w_prepped: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """1x1 MPConv over [x1 | x2] without materialising the concatenation (mp_cat weights folded into w_prepped)."""
    B, H, W, C1 = x1.shape
    C2 = x2.shape[-1]
    Cout = w_prepped.shape[0]
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x1.device, dtype=torch.bfloat16)
    L.check(L.load().dd_mpconv_forward_cat(L.ptr(x1), C1, L.ptr(x2), C2, L.ptr(w_prepped), L.ptr(out), B, H, W, Cout,
                                           L.stream_ptr()))
    _count(1, "mpconv_cat", (B, H, W, C1 + C2, Cout, 1, 1, 0, 0))
    return out


def mpconv_naive(x: Tensor, w_prepped: Tensor, ksize: int, groups: int = 1) -> Tensor:
    B, H, W, Cin = x.shape
    Cout = w_prepped.shape[0]
    out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.bfloat16)
    L.check(L.load().dd_mpconv_forward_naive(L.ptr(x), L.ptr(w_prepped), L.ptr(out), B, H, W, Cin, Cout, ksize, groups,
                                             L.stream_ptr()))
    _count()
    return out


def stem_patches(x_in: Tensor, sigma: Tensor, sigma_data: float, ln_freqs: Tensor,
                 out: Optional[Tensor] = None, cols: int = 64) -> Tensor:
    """Stem input as zero-padded 3x3 patches [B, H, W, cols] (conv_in then runs as a K=cols tensor-core GEMM)."""
    B, Cin, H, W = x_in.shape
    if out is None:
        out = torch.empty((B, H, W, cols), device=x_in.device, dtype=torch.bfloat16)
    L.check(L.load().dd_stem_patches_cols(L.ptr(x_in), L.ptr(sigma), sigma_data, L.ptr(ln_freqs), L.ptr(out), B, Cin, H, W,
                                          cols, L.stream_ptr()))
    _count()
    return out


def conv_out(x: Tensor, w16: Tensor, x_in: Tensor, sigma: Tensor, sigma_data: float,
             x_ref: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """w16: bf16 [16, 9*C] from weight_prep(conv_out.weight, gain=out_gain, pad_rows=16)."""
    B, H, W, Cc = x.shape
    Cout = x_in.shape[1]
    if out is None:
        out = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
    L.check(L.load().dd_conv_out(L.ptr(x), L.ptr(w16), L.ptr(x_in), L.ptr(sigma), sigma_data, L.ptr(x_ref), L.ptr(out),
                                 B, Cc, H, W, Cout, L.stream_ptr()))
    _count()
    return out


def noise_embedding(sigma: Tensor, freqs: Tensor, phases: Tensor, w_noise: Tensor, label_emb: Tensor,
                    label_balance: float, normalize: bool = False, out: Optional[Tensor] = None) -> Tensor:
    B = sigma.numel()
    cemb, cnoise = w_noise.shape
    if out is None:
        out = torch.empty((B, cemb), device=sigma.device, dtype=torch.float32)
    L.check(L.load().dd_noise_embedding(L.ptr(sigma), L.ptr(freqs), L.ptr(phases), cnoise, L.ptr(w_noise),
                                        _is_bf16(w_noise), int(normalize), L.ptr(label_emb), label_balance, L.ptr(out),
                                        B, cemb, L.stream_ptr()))
    _count()
    return out


def make_affine_descs(entries: Sequence[dict], device) -> Tuple[Tensor, int]:
    """Pack dd_affine_desc records into a device buffer.  Each entry: w, gain, out, groups, bias, normalize."""
    arr = (L.AffineDesc * len(entries))()
    max_o = 0
    for i, e in enumerate(entries):
        w = e["w"]
        O, I = w.shape[0], w.shape[1]
        arr[i] = L.AffineDesc(L.ptr(w), L.ptr(e.get("gain")), L.ptr(e["out"]), O, I, e.get("groups", 1), _is_bf16(w),
                              e.get("bias", 0.0), int(e.get("normalize", False)))
        max_o = max(max_o, O)
    raw = bytes(arr)
    buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    return buf, max_o


def emb_affine(descs: Tensor, n: int, max_o: int, emb: Tensor) -> None:
    B, cemb = emb.shape
    L.check(L.load().dd_emb_affine(L.ptr(descs), n, max_o, L.ptr(emb), B, cemb, L.stream_ptr()))
    _count()


def pixnorm_silu(t: Tensor, x_out: Optional[Tensor] = None, s_out: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    Cc = t.shape[-1]
    npix = t.numel() // Cc
    if x_out is None:
        x_out = torch.empty_like(t)
    if s_out is None:
        s_out = torch.empty_like(t)
    L.check(L.load().dd_pixnorm_silu(L.ptr(t), L.ptr(x_out), L.ptr(s_out), npix, Cc, L.stream_ptr()))
    _count()
    return x_out, s_out


def cat_silu(a: Tensor, b: Optional[Tensor], wa: float, wb: float, upsample: bool, need_cat: bool = True):
    B, Ha, Wa, Ca = a.shape
    H, W = (Ha * 2, Wa * 2) if upsample else (Ha, Wa)
    Cb = 0 if b is None else b.shape[-1]
    xcat = torch.empty((B, H, W, Ca + Cb), device=a.device, dtype=torch.bfloat16) if need_cat else None
    s = torch.empty((B, H, W, Ca + Cb), device=a.device, dtype=torch.bfloat16)
    L.check(L.load().dd_cat_silu(L.ptr(a), Ca, L.ptr(b), Cb, wa, wb, int(upsample), L.ptr(xcat), L.ptr(s), B, H, W,
                                 L.stream_ptr()))
    _count()
    return xcat, s


def avgpool2(x: Tensor) -> Tensor:
    B, H, W, Cc = x.shape
    out = torch.empty((B, H // 2, W // 2, Cc), device=x.device, dtype=torch.bfloat16)
    L.check(L.load().dd_avgpool2(L.ptr(x), L.ptr(out), B, H, W, Cc, L.stream_ptr()))
    _count()
    return out


def attention(qk: Tensor, v: Tensor, scale_v: Tensor, heads: int, head_dim: int = 64) -> Tensor:
    B, H, W, Cc = v.shape
    out = torch.empty_like(v)
    L.check(L.load().dd_attention(L.ptr(qk), L.ptr(v), L.ptr(scale_v), L.ptr(out), B, H * W, heads, head_dim,
                                  L.stream_ptr()))
    _count()
    return out


def attention_qkv(qkv: Tensor, heads: int, head_dim: int = 64) -> Tensor:
    """Fused q|k|v attention of the b4_2 lineage: qkv [B, H, W, 3C] -> raw attention output [B, H, W, C]."""
    B, H, W, C3 = qkv.shape
    out = torch.empty((B, H, W, C3 // 3), device=qkv.device, dtype=torch.bfloat16)
    L.check(L.load().dd_attention_qkv(L.ptr(qkv), L.ptr(out), B, H * W, heads, head_dim, L.stream_ptr()))
    _count()
    return out


def sampler_cfg_lerp(d_2b: Tensor, sample: Tensor, cfg_scale: float, t_hat: float, cfg_out: Tensor,
                     x_hat_out: Optional[Tensor], dup: bool = False) -> None:
    """`sample` may be the first half of a duplicated [2B] buffer; n is taken from cfg_out."""
    L.check(L.load().dd_sampler_cfg_lerp(L.ptr(d_2b), L.ptr(sample), cfg_scale, t_hat, L.ptr(cfg_out), L.ptr(x_hat_out),
                                         int(dup), cfg_out.numel(), L.stream_ptr()))
    _count()


def sampler_update(cfg1: Tensor, d2_2b: Optional[Tensor], cfg_scale: float, use_heun: bool, t: float, p: float,
                   noise: Optional[Tensor], sample: Tensor, cfg_out: Optional[Tensor], dup: bool = False) -> None:
    L.check(L.load().dd_sampler_update(L.ptr(cfg1), L.ptr(d2_2b), cfg_scale, int(use_heun), t, p, L.ptr(noise),
                                       L.ptr(sample), L.ptr(cfg_out), int(dup), cfg1.numel(), L.stream_ptr()))
    _count()


def label_embedding(emb_in: Tensor, w_label: Tensor, w_uncond: Tensor, mask: Tensor, normalize: bool = False) -> Tensor:
    Bc, I = emb_in.shape
    Bm = mask.numel()
    cemb = w_label.shape[0]
    out = torch.empty((Bm, cemb), device=emb_in.device, dtype=torch.float32)
    L.check(L.load().dd_label_embedding(L.ptr(emb_in), Bc, I, L.ptr(w_label), L.ptr(w_uncond), _is_bf16(w_label),
                                        L.ptr(mask), Bm, int(normalize), L.ptr(out), cemb, L.stream_ptr()))
    _count()
    return out


def sigma_logvar(sigma: Tensor, freqs: Tensor, phases: Tensor, w: Tensor) -> Tensor:
    out = torch.empty(sigma.numel(), device=sigma.device, dtype=torch.float32)
    L.check(L.load().dd_sigma_logvar(L.ptr(sigma), sigma.numel(), L.ptr(freqs), L.ptr(phases), freqs.numel(), L.ptr(w),
                                     _is_bf16(w), L.ptr(out), L.stream_ptr()))
    _count()
    return out


def axpby(a: Tensor, b: Tensor, alpha: float, beta: float, clip: float = 0.0) -> Tensor:
    out = torch.empty_like(a)
    L.check(L.load().dd_axpby(L.ptr(a), L.ptr(b), alpha, beta, clip, L.ptr(out), a.numel(), L.stream_ptr()))
    _count()
    return out


def roll_pad_w(x: Tensor, shift: int, pad: int, copies: int = 1, out: Optional[Tensor] = None) -> Tensor:
    """torch.roll(x, shift, -1) -> circular pad by `pad` columns per side (-> repeat `copies` times along dim 0);
    fp32 [..., W] -> [copies * x.shape[0], ..., W + 2 pad] (pipeline.py:651-656)."""
    L.require_cuda(x)
    W = x.shape[-1]
    rows = x.numel() // W
    if out is None:
        out = torch.empty((copies * x.shape[0],) + tuple(x.shape[1:-1]) + (W + 2 * pad,), device=x.device,
                          dtype=torch.float32)
    L.check(L.load().dd_roll_pad_w(L.ptr(x), L.ptr(out), rows, W, int(shift), int(pad), int(copies), L.stream_ptr()))
    _count()
    return out


def crop_unroll_w(xp: Tensor, shift: int, pad: int, out: Optional[Tensor] = None) -> Tensor:
    """torch.roll(xp[..., pad:-pad], -shift, -1) (pipeline.py:729-732); fp32 [..., W + 2 pad] -> [..., W]."""
    L.require_cuda(xp)
    W = xp.shape[-1] - 2 * pad
    rows = xp.numel() // xp.shape[-1]
    if out is None:
        out = torch.empty(tuple(xp.shape[:-1]) + (W,), device=xp.device, dtype=torch.float32)
    L.check(L.load().dd_crop_unroll_w(L.ptr(xp), L.ptr(out), rows, W, int(shift), int(pad), L.stream_ptr()))
    _count()
    return out


def stereo_fix_noise(noise: Tensor, fresh: Tensor, t: float) -> Tensor:
    """pipeline.py:638-640: even channels take their odd neighbour's noise, then mp_sum(fresh, noise, t)."""
    L.require_cuda(noise, fresh)
    B, Cc = noise.shape[0], noise.shape[1]
    out = torch.empty_like(noise)
    L.check(L.load().dd_stereo_fix_noise(L.ptr(noise), L.ptr(fresh), float(t), L.ptr(out), B, Cc,
                                         noise.numel() // (B * Cc), L.stream_ptr()))
    _count()
    return out


def mp_fourier(x: Tensor, freqs: Tensor, phases: Tensor) -> Tensor:
    out = torch.empty((x.numel(), freqs.numel()), device=x.device, dtype=torch.float32)
    L.check(L.load().dd_mp_fourier(L.ptr(x), x.numel(), L.ptr(freqs), L.ptr(phases), freqs.numel(), L.ptr(out),
                                   L.stream_ptr()))
    _count()
    return out


def stft_mel(raw: Tensor, window: Tensor, tw: Tensor, tw_half: Tensor, n_fft: int, hop: int, fb: dict,
             exponent: float, mean: float, scale: float, window2: Optional[Tensor] = None,
             coef1: Optional[Tensor] = None, coef2: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """raw [S, L] fp32 -> [S, n_filters, 1 + L // hop] fp32 (mel-STFT encode)."""
    S, Ln = raw.shape
    T = 1 + Ln // hop
    nf = fb["start"].numel()
    if out is None:
        out = torch.empty((S, nf, T), device=raw.device, dtype=torch.float32)
    L.check(L.load().dd_stft_mel(L.ptr(raw), S, Ln, L.ptr(window), L.ptr(window2), L.ptr(coef1), L.ptr(coef2), L.ptr(tw),
                                 L.ptr(tw_half), n_fft, hop,
                                 L.ptr(fb["start"]), L.ptr(fb["count"]), L.ptr(fb["offset"]), L.ptr(fb["weight"]), nf,
                                 exponent, mean, scale, L.ptr(out), T, L.stream_ptr()))
    _count()
    return out


def fgla_istft(state: Optional[Tensor], mag_tk: Tensor, stereo: bool, interp_t: float, window: Tensor, tw: Tensor,
               tw_half: Tensor, n_fft: int, hop: int, ola: Tensor) -> None:
    S, T, _ = mag_tk.shape
    L.check(L.load().dd_fgla_istft(L.ptr(state), L.ptr(mag_tk), S, T, int(stereo), interp_t, L.ptr(window), L.ptr(tw),
                                   L.ptr(tw_half), n_fft, hop, L.ptr(ola), ola.shape[-1], L.stream_ptr()))
    _count()


def fgla_stft_update(ola: Tensor, env: Tensor, state: Tensor, momentum: float, first: bool, window: Tensor, tw: Tensor,
                     tw_half: Tensor, n_fft: int, hop: int) -> None:
    S, T, _ = state.shape[:3]
    L.check(L.load().dd_fgla_stft_update(L.ptr(ola), L.ptr(env), S, T, hop * (T - 1), L.ptr(window), L.ptr(tw),
                                         L.ptr(tw_half), n_fft, hop, L.ptr(state), momentum, int(first), L.stream_ptr()))
    _count()


def ola_finalize(ola: Tensor, env: Tensor, n_fft: int, length: int) -> Tensor:
    S = ola.shape[0]
    out = torch.empty((S, length), device=ola.device, dtype=torch.float32)
    L.check(L.load().dd_ola_finalize(L.ptr(ola), L.ptr(env), S, ola.shape[-1], n_fft, length, L.ptr(out), L.stream_ptr()))
    _count()
    return out


def attention_axis(qkv: Tensor, heads: int, axis: int, head_dim: int = 64) -> Tensor:
    """Axis attention on a channels_last_3d tensor [B, Z, H, W, 3C] (q|k|v thirds) -> [B, Z, H, W, C]."""
    B, Z, H, W, C3 = qkv.shape
    out = torch.empty((B, Z, H, W, C3 // 3), device=qkv.device, dtype=torch.bfloat16)
    L.check(L.load().dd_attention_axis(L.ptr(qkv), L.ptr(out), B, Z, H, W, heads, head_dim, axis, L.stream_ptr()))
    _count()
    return out


# ---------------------------------------------------------------------------------------------------------
# backward pass (train step): thin wrappers over the dd_*_bwd / dd_mpconv_wgrad entry points
# ---------------------------------------------------------------------------------------------------------
def attention_train(qk: Tensor, v: Tensor, scale_v: Tensor, heads: int, head_dim: int = 64) -> Tuple[Tensor, Tensor]:
    """Train-mode attention: returns (mp_silu(a * scale_v), a) with a = softmax(qk^T/8) v."""
    B, H, W, Cc = v.shape
    out = torch.empty_like(v)
    raw = torch.empty_like(v)
    L.check(L.load().dd_attention_train(L.ptr(qk), L.ptr(v), L.ptr(scale_v), L.ptr(out), L.ptr(raw), B, H * W, heads,
                                        head_dim, L.stream_ptr()))
    _count()
    return out, raw


def mpconv_wgrad(x: Tensor, dy: Tensor, ksize: int, groups: int = 1, scale: float = 1.0,
                 out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    """dW_eff [Cout, taps, Cin/groups] fp32 of an MPConv (tcgen05, MN-major operands)."""
    B, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    if out is None:
        out = torch.empty((Cout, ksize * ksize, Cin // groups), device=x.device, dtype=torch.float32)
    if timing is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    L.check(L.load().dd_mpconv_wgrad(L.ptr(x), L.ptr(dy), L.ptr(out), B, H, W, Cin, Cout, ksize, groups, scale,
                                     int(accumulate), L.stream_ptr()))
    if timing is not None:
        ev1.record()
        timing.append((2.0 * B * H * W * Cout * (Cin // groups) * ksize * ksize, ev0, ev1,
                       ("wgrad", B, H, W, Cin, Cout, ksize, groups)))
    _count(1, "mpconv_wgrad", (B, H, W, Cin, Cout, ksize, groups))
    return out


def weight_transpose(w_prepped: Tensor, cout: int, cin_g: int, taps: int, groups: int,
                     out: Optional[Tensor] = None) -> Tensor:
    """bf16 [Cout, taps*cin_g] -> [Cin, taps*cout_g] with reversed taps (dgrad operand of dd_mpconv_forward)."""
    if out is None:
        out = torch.empty((groups * cin_g, taps, cout // groups), device=w_prepped.device, dtype=torch.bfloat16)
    L.check(L.load().dd_weight_transpose(L.ptr(w_prepped), L.ptr(out), cout, cin_g, taps, groups, L.stream_ptr()))
    _count()
    return out


def make_wprep_descs(entries: Sequence[dict], device) -> Tuple[Tensor, int]:
    """Pack dd_wprep_desc records (dicts with w, out, gain, gain_host, O, I_g, taps, normalize, perm, head_dim,
    row_stride); returns (device buffer, total_rows)."""
    arr = (L.WprepDesc * len(entries))()
    rows = 0
    for i, e in enumerate(entries):
        arr[i] = L.WprepDesc(L.ptr(e["w"]), L.ptr(e["out"]), L.ptr(e.get("gain")), e.get("gain_host", 1.0), _is_bf16(e["w"]),
                             e["O"], e["I_g"], e["taps"], int(e.get("normalize", False)), e.get("perm", 0),
                             e.get("head_dim", 0), e.get("row_stride", 0) or e["I_g"] * e["taps"], rows)
        rows += e["O"]
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device), rows


def weight_prep_batched(descs: Tensor, n: int, total_rows: int) -> None:
    L.check(L.load().dd_weight_prep_batched(L.ptr(descs), n, total_rows, L.stream_ptr()))
    _count()


def weight_normalize_batched(descs: Tensor, n: int, total_rows: int) -> None:
    L.check(L.load().dd_weight_normalize_batched(L.ptr(descs), n, total_rows, L.stream_ptr()))
    _count()


def make_wtrans_descs(entries: Sequence[dict], device) -> Tuple[Tensor, int]:
    """Pack dd_wtrans_desc records (dicts with src, dst, cout_g, cin_g, taps, groups); returns (buffer, total_tiles)."""
    arr = (L.WtransDesc * len(entries))()
    tiles = 0
    for i, e in enumerate(entries):
        arr[i] = L.WtransDesc(L.ptr(e["src"]), L.ptr(e["dst"]), e["cout_g"], e["cin_g"], e["taps"], e["groups"], tiles)
        tiles += ((e["cin_g"] + 31) // 32) * ((e["cout_g"] + 31) // 32) * e["groups"] * e["taps"]
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device), tiles


def weight_transpose_batched(descs: Tensor, n: int, total_tiles: int) -> None:
    L.check(L.load().dd_weight_transpose_batched(L.ptr(descs), n, total_tiles, L.stream_ptr()))
    _count()


def make_wbwd_descs(entries: Sequence[dict], device) -> Tuple[Tensor, int]:
    """Pack dd_wbwd_desc records (dicts with w, dweff, dw, gain, dgain, gain_host, O, I_g, taps, normalize, perm,
    head_dim, row_stride, accumulate, t_cout_g) into a device buffer; returns (buffer, total_rows)."""
    arr = (L.WbwdDesc * len(entries))()
    rows = 0
    for i, e in enumerate(entries):
        arr[i] = L.WbwdDesc(L.ptr(e["w"]), L.ptr(e["dweff"]), L.ptr(e["dw"]), L.ptr(e.get("gain")), L.ptr(e.get("dgain")),
                            e.get("gain_host", 1.0), e["O"], e["I_g"], e["taps"], int(e.get("normalize", False)),
                            e.get("perm", 0), e.get("head_dim", 0), e.get("row_stride", 0) or e["I_g"] * e["taps"],
                            int(e.get("accumulate", False)), rows, int(e.get("t_cout_g", 0)))
        rows += e["O"]
    buf = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    return buf, rows


def weight_prep_bwd(descs: Tensor, n: int, total_rows: int) -> None:
    L.check(L.load().dd_weight_prep_bwd(L.ptr(descs), n, total_rows, L.stream_ptr()))
    _count()


# ---------------------------------------------------------------------------------------------------------
# optimizer-side sweep: clip_grad_norm_ + AdamW + EMA lerps + normalize_weights in two launches
# ---------------------------------------------------------------------------------------------------------
OPTIM_FLAT_ROW = 4096       # row length used for tensors without weight-norm (scalars, gains, un-normalised weights)


def pack_gnorm_descs(grads: Sequence[Tensor]):
    """dd_gnorm_desc records for a list of fp32 gradient tensors; returns (ctypes array, total_chunks).  Host only."""
    arr = (L.GnormDesc * len(grads))()
    chunks = 0
    for i, g in enumerate(grads):
        n = g.numel()
        arr[i] = L.GnormDesc(L.ptr(g), n, chunks, 0)
        chunks += (n + L.GNORM_CHUNK - 1) // L.GNORM_CHUNK
    return arr, chunks


def pack_optim_descs(entries: Sequence[dict]):
    """dd_optim_desc records (dicts with p, g, m, v, emas (list, <= 4), fan_in (0 = no weight-norm)); returns
    (ctypes array, total_rows).  A weight-normalised tensor [O, ...] is viewed as O rows of fan_in elements (one CTA
    per output channel, as mp_tools.normalize does); anything else is cut into rows of OPTIM_FLAT_ROW.  Host only."""
    arr = (L.OptimDesc * len(entries))()
    rows_total = 0
    for i, e in enumerate(entries):
        n = e["p"].numel()
        fan_in = int(e.get("fan_in", 0))
        if fan_in > 0:
            if n % fan_in != 0:
                raise ValueError(f"pack_optim_descs: numel {n} is not a multiple of fan_in {fan_in}")
            rows, row_len, normalize = n // fan_in, fan_in, 1
        else:
            rows, row_len, normalize = (n + OPTIM_FLAT_ROW - 1) // OPTIM_FLAT_ROW, OPTIM_FLAT_ROW, 0
        emas = list(e.get("emas", ()))
        if len(emas) > L.OPTIM_MAX_EMA:
            raise ValueError(f"pack_optim_descs: at most {L.OPTIM_MAX_EMA} EMA copies per launch")
        ema_ptrs = (C.c_void_p * L.OPTIM_MAX_EMA)(*[L.ptr(t) for t in emas] + [None] * (L.OPTIM_MAX_EMA - len(emas)))
        arr[i] = L.OptimDesc(L.ptr(e["p"]), L.ptr(e["g"]), L.ptr(e["m"]), L.ptr(e["v"]), ema_ptrs, n, rows, row_len,
                             normalize, rows_total)
        rows_total += rows
    return arr, rows_total


def descs_to_device(arr, device) -> Tensor:
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


def grad_norm_clip(descs: Tensor, n: int, total_chunks: int, partials: Tensor, max_norm: float, out: Tensor) -> None:
    """out[0] = ||g||_2 over every described tensor, out[1] = clip_grad_norm_'s coefficient (2 launches)."""
    L.require_cuda(descs, partials, out)
    L.check(L.load().dd_grad_norm_clip(L.ptr(descs), n, total_chunks, L.ptr(partials), float(max_norm), L.ptr(out),
                                       L.stream_ptr()))
    _count(2)


def optim_step_batched(descs: Tensor, n: int, total_rows: int, hyper: "L.OptimHyper",
                       norm_coef: Optional[Tensor]) -> None:
    L.require_cuda(descs, norm_coef)
    L.check(L.load().dd_optim_step_batched(L.ptr(descs), n, total_rows, C.byref(hyper), L.ptr(norm_coef),
                                           L.stream_ptr()))
    _count()


def silu_scale_bwd(dy: Tensor, coef: float, pre: Tensor, scale: Tensor, dscale: Tensor,
                   out: Optional[Tensor] = None) -> Tensor:
    B = dy.shape[0]
    Cc = dy.shape[-1]
    npix = dy.numel() // (B * Cc)
    if out is None:
        out = torch.empty_like(dy)
    L.check(L.load().dd_silu_scale_bwd(L.ptr(dy), coef, L.ptr(pre), L.ptr(scale), L.ptr(out), L.ptr(dscale), B, npix, Cc,
                                       L.stream_ptr()))
    _count()
    return out


def pixnorm_silu_bwd(g: Tensor, ca: float, ds: Tensor, t0: Tensor) -> Tensor:
    Cc = t0.shape[-1]
    out = torch.empty_like(t0)
    L.check(L.load().dd_pixnorm_silu_bwd(L.ptr(g), ca, L.ptr(ds), L.ptr(t0), L.ptr(out), t0.numel() // Cc, Cc,
                                         L.stream_ptr()))
    _count()
    return out


def cat_silu_bwd(d_xc: Tensor, c1: float, d_s: Tensor, xc: Tensor, a_prev: Optional[Tensor], clip: float, wa: float,
                 wb: float, upsample: bool, Ca: int, Cb: int) -> Tuple[Tensor, Optional[Tensor]]:
    B, H, W, _ = xc.shape
    Ha, Wa = (H // 2, W // 2) if upsample else (H, W)
    da = torch.empty((B, Ha, Wa, Ca), device=xc.device, dtype=torch.bfloat16)
    db = torch.empty((B, H, W, Cb), device=xc.device, dtype=torch.bfloat16) if Cb else None
    L.check(L.load().dd_cat_silu_bwd(L.ptr(d_xc), c1, L.ptr(d_s), L.ptr(xc), L.ptr(a_prev), clip, wa, wb, int(upsample),
                                     L.ptr(da), L.ptr(db), B, H, W, Ca, Cb, L.stream_ptr()))
    _count()
    return da, db


def enc_grad_combine(dx0: Tensor, down: bool, dskip: Optional[Tensor], x_prev: Optional[Tensor], clip: float,
                     shape: Tuple[int, int, int, int]) -> Tensor:
    B, H, W, Cc = shape
    out = torch.empty(shape, device=dx0.device, dtype=torch.bfloat16)
    L.check(L.load().dd_enc_grad_combine(L.ptr(dx0), int(down), L.ptr(dskip), L.ptr(x_prev), clip, L.ptr(out), B, H, W, Cc,
                                         L.stream_ptr()))
    _count()
    return out


def attn_in_bwd(g3: Tensor, ca: float, dxv: Tensor, dxs: Tensor, x2: Tensor, c_qk: Tensor, dc_qk: Tensor) -> Tensor:
    B = x2.shape[0]
    Cc = x2.shape[-1]
    out = torch.empty_like(x2)
    L.check(L.load().dd_attn_in_bwd(L.ptr(g3), ca, L.ptr(dxv), L.ptr(dxs), L.ptr(x2), L.ptr(c_qk), L.ptr(out), L.ptr(dc_qk),
                                    B, x2.numel() // (B * Cc), Cc, L.stream_ptr()))
    _count()
    return out


def attention_bwd(qk: Tensor, v: Tensor, a_raw: Tensor, d_a: Tensor, heads: int, head_dim: int = 64):
    B, H, W, Cc = v.shape
    N = H * W
    dqk = torch.empty_like(qk)
    dv = torch.empty_like(v)
    stats = torch.empty((B, heads, N, 2), device=v.device, dtype=torch.float32)
    L.check(L.load().dd_attention_bwd(L.ptr(qk), L.ptr(v), L.ptr(a_raw), L.ptr(d_a), L.ptr(dqk), L.ptr(dv), L.ptr(stats), B,
                                      N, heads, head_dim, L.stream_ptr()))
    _count(2)
    return dqk, dv


def make_affine_bwd_descs(entries: Sequence[dict], device) -> Tuple[Tensor, int, int]:
    arr = (L.AffineBwdDesc * len(entries))()
    max_o = max_cols = 0
    for i, e in enumerate(entries):
        w = e["w"]
        O, I = w.shape[0], w.shape[1]
        arr[i] = L.AffineBwdDesc(L.ptr(w), L.ptr(e.get("gain")), L.ptr(e["dout"]), L.ptr(e["dweff"]), L.ptr(e["rowscale"]),
                                 O, I, e.get("groups", 1), int(e.get("normalize", False)))
        max_o = max(max_o, O)
        max_cols = max(max_cols, I * e.get("groups", 1))
    buf = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    return buf, max_o, max_cols


def emb_affine_bwd(descs: Tensor, n: int, max_o: int, max_cols: int, emb: Optional[Tensor], demb: Optional[Tensor],
                   first: int = 0, B: int = 0, cemb: int = 0) -> None:
    """Stage 1 (emb given: weight gradients) and / or stage 2 (demb given: embedding gradient) over descriptors
    [first, first + n) of the table."""
    if emb is not None:
        B, cemb = emb.shape
    elif demb is not None:
        B, cemb = demb.shape
    base = descs.data_ptr() + first * C.sizeof(L.AffineBwdDesc)
    L.check(L.load().dd_emb_affine_bwd(base, n, max_o, max_cols, L.ptr(emb), L.ptr(demb), B, cemb, L.stream_ptr()))
    _count((emb is not None) + (demb is not None))


def noise_embedding_bwd(sigma: Tensor, freqs: Tensor, phases: Tensor, w_noise: Tensor, label_emb: Tensor,
                        label_balance: float, demb: Tensor, normalize: bool,
                        dweff: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    B = sigma.numel()
    cemb, cnoise = w_noise.shape
    if dweff is None:
        dweff = torch.empty((cemb, cnoise), device=sigma.device, dtype=torch.float32)
    dlabel = torch.empty((B, cemb), device=sigma.device, dtype=torch.float32)
    L.check(L.load().dd_noise_embedding_bwd(L.ptr(sigma), L.ptr(freqs), L.ptr(phases), cnoise, L.ptr(w_noise),
                                            int(normalize), L.ptr(label_emb), label_balance, L.ptr(demb), L.ptr(dweff),
                                            L.ptr(dlabel), B, cemb, L.stream_ptr()))
    _count()
    return dweff, dlabel


def label_embedding_bwd(emb_in: Tensor, mask: Tensor, dout: Tensor) -> Tuple[Tensor, Tensor]:
    Bc, I = emb_in.shape
    Bm, cemb = dout.shape
    dwl = torch.empty((cemb, I), device=dout.device, dtype=torch.float32)
    dwu = torch.empty((cemb, 1), device=dout.device, dtype=torch.float32)
    L.check(L.load().dd_label_embedding_bwd(L.ptr(emb_in), Bc, I, L.ptr(mask), Bm, L.ptr(dout), L.ptr(dwl), L.ptr(dwu),
                                            cemb, L.stream_ptr()))
    _count()
    return dwl, dwu


def sigma_logvar_bwd(sigma: Tensor, freqs: Tensor, phases: Tensor, dout: Tensor) -> Tensor:
    n = freqs.numel()
    dw = torch.empty((1, n), device=sigma.device, dtype=torch.float32)
    L.check(L.load().dd_sigma_logvar_bwd(L.ptr(sigma), sigma.numel(), L.ptr(freqs), L.ptr(phases), n, L.ptr(dout),
                                         L.ptr(dw), 0, L.stream_ptr()))
    _count()
    return dw


def head_grad(dD: Tensor, sigma: Tensor, sigma_data: float, x_ref: Optional[Tensor], cpad: int = 32) -> Tensor:
    B, Cout, H, W = dD.shape
    out = torch.empty((B, H, W, cpad), device=dD.device, dtype=torch.bfloat16)
    L.check(L.load().dd_head_grad(L.ptr(dD), L.ptr(sigma), sigma_data, L.ptr(x_ref), L.ptr(out), B, Cout, H, W, cpad,
                                  L.stream_ptr()))
    _count()
    return out


# ---------------------------------------------------------------------------------------------------------
# DAE_D3 decoder glue (stereo depth folded into channels, physical reflection halo along W)
# ---------------------------------------------------------------------------------------------------------
def weight_prep_z2(w: Tensor, gain: Optional[Tensor] = None, gain_host: float = 1.0, i_stride: int = 0,
                   out: Optional[Tensor] = None) -> Tensor:
    """MPConv3D weight [O, I, kz, kh, kw] -> bf16 [2*O, kh*kw, i_stride or (kz*I)] in the folded-stereo layout."""
    L.require_cuda(w)
    w = w.contiguous()
    O, I, kz = w.shape[0], w.shape[1], w.shape[2]
    taps = w.shape[3] * w.shape[4]
    n_in = 2 * I if kz == 2 else I
    if out is None:
        out = torch.empty((2 * O, taps, i_stride or n_in), device=w.device, dtype=torch.bfloat16)
    L.check(L.load().dd_weight_prep_z2(L.ptr(w), _is_bf16(w), L.ptr(out), O, I, kz, taps, L.ptr(gain), gain_host,
                                       i_stride, L.stream_ptr()))
    _count()
    return out


def reflect_fill_w(x: Tensor, pw: int) -> Tensor:
    B, H, Wp, Cc = x.shape
    L.check(L.load().dd_reflect_fill_w(L.ptr(x), B, H, Wp, Cc, pw, L.stream_ptr()))
    _count()
    return x


def dae_stem(latents: Tensor, latent_channels: int, pw: int, cpad: int = 32) -> Tensor:
    B, C2, H, W = latents.shape
    out = torch.empty((B, H, W + 2 * pw, cpad), device=latents.device, dtype=torch.bfloat16)
    L.check(L.load().dd_dae_stem(L.ptr(latents), L.ptr(out), B, latent_channels, H, W, pw, cpad, L.stream_ptr()))
    _count()
    return out


def up2_silu_pad(a: Tensor, pw: int) -> Tuple[Tensor, Tensor]:
    B, Ha, Wpa, Cc = a.shape
    Wa = Wpa - 2 * pw
    xc = torch.empty((B, 2 * Ha, 2 * Wa + 2 * pw, Cc), device=a.device, dtype=torch.bfloat16)
    s = torch.empty_like(xc)
    L.check(L.load().dd_up2_silu_pad(L.ptr(a), L.ptr(xc), L.ptr(s), B, Ha, Wa, Cc, pw, L.stream_ptr()))
    _count()
    return xc, s


def conv5x5_out(x: Tensor, w25: Tensor, gain: Optional[Tensor], pw: int) -> Tensor:
    B, H, Wp, C2 = x.shape
    W = Wp - 2 * pw
    out = torch.empty((B, 2, H, W), device=x.device, dtype=torch.float32)
    L.check(L.load().dd_conv5x5_out(L.ptr(x), L.ptr(w25), L.ptr(gain), L.ptr(out), B, H, W, C2 // 2, pw, L.stream_ptr()))
    _count()
    return out


def ddec_stem(x_in: Tensor, x_ref: Tensor, sigma: Tensor, sigma_data: float, k: int, pw: int, cpad: int = 64) -> Tensor:
    B, _, Fq, W = x_in.shape
    out = torch.empty((B, Fq, W + 2 * pw, cpad), device=x_in.device, dtype=torch.bfloat16)
    L.check(L.load().dd_ddec_stem(L.ptr(x_in), L.ptr(x_ref), L.ptr(sigma), sigma_data, L.ptr(out), B, Fq, W, k, pw, cpad,
                                  L.stream_ptr()))
    _count()
    return out


def avgpool2_pad(x: Tensor, pw: int) -> Tensor:
    B, H, Wp, Cc = x.shape
    W = Wp - 2 * pw
    out = torch.empty((B, H // 2, W // 2 + 2 * pw, Cc), device=x.device, dtype=torch.bfloat16)
    L.check(L.load().dd_avgpool2_pad(L.ptr(x), L.ptr(out), B, H, W, Cc, pw, L.stream_ptr()))
    _count()
    return out


def ddec_head(f: Tensor, x_in: Tensor, sigma: Tensor, sigma_data: float, pw: int) -> Tensor:
    B, H, Wp, Cst = f.shape
    out = torch.empty_like(x_in)
    L.check(L.load().dd_ddec_head(L.ptr(f), L.ptr(x_in), L.ptr(sigma), sigma_data, L.ptr(out), B, H, Wp - 2 * pw, pw, Cst,
                                  L.stream_ptr()))
    _count()
    return out


def q4_stem(x_in: Tensor, x_ref: Tensor, sigma: Tensor, sigma_data: float, wa: float, wb: float, k: int,
            cpad: int = 32) -> Tensor:
    B, Cc, Fq, W = x_in.shape
    out = torch.empty((B, Fq, W, cpad), device=x_in.device, dtype=torch.bfloat16)
    L.check(L.load().dd_q4_stem(L.ptr(x_in), L.ptr(x_ref), L.ptr(sigma), sigma_data, wa, wb, L.ptr(out), B, Cc, Fq, W, k,
                                cpad, L.stream_ptr()))
    _count()
    return out


# ---------------------------------------------------------------------------------------------------------
# MDCT side of the live format (framing / |.| / overlap-add / mel linearisation around a library GEMM)
# ---------------------------------------------------------------------------------------------------------
def gemm_f32(a: Tensor, a_strides, b: Tensor, b_strides, c: Tensor, c_strides, M: int, N: int, K: int, batch: int = 1, *,
             gather=None, a_transform=None, relu: bool = False) -> Tensor:
    """C[m][n] = sum_k A[m][k] B[k][n] on fp32 tensors addressed by element strides (m, k, batch) / (k, n, batch) /
    (m, n, batch).  gather = (hop, pad_left, length): B is the reflect-padded frames of the raw rows in `b`.
    a_transform = (scale, offset, power): A elements are read as clip(a * scale + offset, 0) ** power."""
    hop, pad, glen = gather if gather is not None else (0, 0, 0)
    sc, off, pw = a_transform if a_transform is not None else (1.0, 0.0, 0.0)
    L.check(L.load().dd_gemm_f32(L.ptr(a), *a_strides, L.ptr(b), *b_strides, L.ptr(c), *c_strides, M, N, K, batch, hop, pad,
                                 glen, sc, off, pw, int(relu), L.stream_ptr()))
    _count()
    return c


def frame_reflect(raw: Tensor, block_width: int, hop: int, pad_left: int, n_frames: int) -> Tensor:
    S, Ln = raw.shape
    out = torch.empty((S, n_frames, block_width), device=raw.device, dtype=torch.float32)
    L.check(L.load().dd_frame_reflect(L.ptr(raw), L.ptr(out), S, Ln, n_frames, block_width, hop, pad_left, L.stream_ptr()))
    _count()
    return out


def complex_abs(y: Tensor, scale: float) -> Tensor:
    S, N2, T = y.shape
    out = torch.empty((S, N2 // 2, T), device=y.device, dtype=torch.float32)
    L.check(L.load().dd_complex_abs(L.ptr(y), L.ptr(out), S, N2 // 2, T, scale, L.stream_ptr()))
    _count()
    return out


def mdct_ola(y: Tensor) -> Tensor:
    S, T, N2 = y.shape
    out = torch.empty((S, (T - 1) * (N2 // 2)), device=y.device, dtype=torch.float32)
    L.check(L.load().dd_mdct_ola(L.ptr(y), L.ptr(out), S, T, N2 // 2, L.stream_ptr()))
    _count()
    return out


def mel_linearize(mel: Tensor, offset: float, inv_exponent: float) -> Tensor:
    out = torch.empty_like(mel)
    L.check(L.load().dd_mel_linearize(L.ptr(mel), L.ptr(out), mel.numel(), offset, inv_exponent, L.stream_ptr()))
    _count()
    return out


def dae_enc_patches(mel: Tensor, pw: int) -> Tensor:
    B, _, H, W = mel.shape
    out = torch.empty((B, H, W + 2 * pw, 128), device=mel.device, dtype=torch.bfloat16)
    L.check(L.load().dd_dae_enc_patches(L.ptr(mel), L.ptr(out), B, H, W, pw, L.stream_ptr()))
    _count()
    return out


def dae_latents_pool(f: Tensor, latent_channels: int, pw: int, ratio: int) -> Tensor:
    B, H, Wp, Cst = f.shape
    W = Wp - 2 * pw
    out = torch.empty((B, 2 * latent_channels, H // ratio, W // ratio), device=f.device, dtype=torch.float32)
    L.check(L.load().dd_dae_latents_pool(L.ptr(f), L.ptr(out), B, latent_channels, H, W, pw, Cst, ratio, L.stream_ptr()))
    _count()
    return out


# ---- dae_edm2_q4.DAE ends (csrc/dae.cu) ----
def pack_nhwc(x: Tensor, cpad: int, ones_channel: int = -1) -> Tensor:
    """fp32 NCHW -> bf16 NHWC padded to `cpad` channels; `ones_channel` carries a convolution bias (dd_pack_nhwc)."""
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, cpad), device=x.device, dtype=torch.bfloat16)
    L.check(L.load().dd_pack_nhwc(L.ptr(x), L.ptr(out), B, C, H, W, cpad, ones_channel, L.stream_ptr()))
    _count()
    return out


def unpack_nchw(x: Tensor, channels: int) -> Tensor:
    """bf16 NHWC -> fp32 NCHW, the first `channels` channels (dd_unpack_nchw)."""
    B, H, W, Cpad = x.shape
    out = torch.empty((B, channels, H, W), device=x.device, dtype=torch.float32)
    L.check(L.load().dd_unpack_nchw(L.ptr(x), L.ptr(out), B, channels, H, W, Cpad, L.stream_ptr()))
    _count()
    return out


def patches5x5(x: Tensor, cols: int = 64) -> Tensor:
    """Zero-padded 5x5 patches of an fp32 NCHW image plus the bias column, as the K = cols operand of conv_in (dd_patches5x5)."""
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, cols), device=x.device, dtype=torch.bfloat16)
    L.check(L.load().dd_patches5x5(L.ptr(x), L.ptr(out), B, C, H, W, cols, L.stream_ptr()))
    _count()
    return out


def conv5x5_dense(x: Tensor, w: Tensor, gain: Optional[Tensor]) -> Tensor:
    """conv_out (5,5) C -> Cout <= 4 on bf16 NHWC, fp32 NCHW result; w fp32 [Cout][25][C] pre-scaled (dd_conv5x5_dense)."""
    B, H, W, C = x.shape
    cout = w.shape[0]
    out = torch.empty((B, cout, H, W), device=x.device, dtype=torch.float32)
    L.check(L.load().dd_conv5x5_dense(L.ptr(x), L.ptr(w), L.ptr(gain), L.ptr(out), B, H, W, C, cout, L.stream_ptr()))
    _count()
    return out


# ---- MS_MDCT_DualFormat, second lineage (csrc/mdct.cu) ----
def mel_blend(mels: Tensor, ww: Tensor, exponent: float, offset: float, inv_scale: float) -> Tensor:
    """mels [n_win][S][F][T], ww [F][n_win] -> ((sum_i mels[i] * ww[:, i]) ** exponent + offset) * inv_scale (dd_mel_blend)."""
    n_win, S, F, T = mels.shape
    out = torch.empty((S, F, T), device=mels.device, dtype=torch.float32)
    L.check(L.load().dd_mel_blend(L.ptr(mels), L.ptr(ww), n_win, S, F, T, exponent, offset, inv_scale, L.ptr(out), L.stream_ptr()))
    _count()
    return out


def mdct_phase_psd(y: Tensor, inv_density: Tensor, exponent: float, offset: float, inv_scale: float,
                   phase_mul: float) -> Tuple[Tensor, Tensor]:
    """MCLT rows y [S][2N][T] -> (phase, psd) [S][N][T] (dd_mdct_phase_psd)."""
    S, N2, T = y.shape
    phase = torch.empty((S, N2 // 2, T), device=y.device, dtype=torch.float32)
    psd = torch.empty_like(phase)
    L.check(L.load().dd_mdct_phase_psd(L.ptr(y), L.ptr(inv_density), S, N2 // 2, T, exponent, offset, inv_scale, phase_mul,
                                       L.ptr(phase), L.ptr(psd), L.stream_ptr()))
    _count()
    return phase, psd
