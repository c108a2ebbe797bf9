"""ctypes binding of the C-ABI library (include/dualdiffusion_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing it is
built in-tree (nvcc cross-compiles); if that fails, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import build as _build

_lib: Optional[C.CDLL] = None

c_void_p, c_int, c_long, c_float = C.c_void_p, C.c_int, C.c_long, C.c_float


class ConvEpilogue(C.Structure):
    """struct dd_conv_epilogue"""
    _fields_ = [("mode", c_int), ("mode2", c_int), ("alpha", c_float), ("beta", c_float), ("clip", c_float),
                ("scale", c_void_p), ("scale2", c_void_p), ("residual", c_void_p), ("out2", c_void_p)]


class AffineDesc(C.Structure):
    """struct dd_affine_desc"""
    _fields_ = [("w", c_void_p), ("gain", c_void_p), ("out", c_void_p), ("O", c_int), ("I", c_int),
                ("groups", c_int), ("w_is_bf16", c_int), ("bias", c_float), ("normalize", c_int)]


class WbwdDesc(C.Structure):
    """struct dd_wbwd_desc"""
    _fields_ = [("w", c_void_p), ("dweff", c_void_p), ("dw", c_void_p), ("gain", c_void_p), ("dgain", c_void_p),
                ("gain_host", c_float), ("O", c_int), ("I_g", c_int), ("taps", c_int), ("normalize", c_int),
                ("perm", c_int), ("head_dim", c_int), ("row_stride", c_int), ("accumulate", c_int),
                ("row_begin", c_int), ("t_cout_g", c_int)]


class WprepDesc(C.Structure):
    """struct dd_wprep_desc"""
    _fields_ = [("w", c_void_p), ("out", c_void_p), ("gain", c_void_p), ("gain_host", c_float), ("w_is_bf16", c_int),
                ("O", c_int), ("I_g", c_int), ("taps", c_int), ("normalize", c_int), ("perm", c_int), ("head_dim", c_int),
                ("row_stride", c_int), ("row_begin", c_int)]


class WtransDesc(C.Structure):
    """struct dd_wtrans_desc"""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("cout_g", c_int), ("cin_g", c_int), ("taps", c_int),
                ("groups", c_int), ("tile_begin", c_int)]


class AffineBwdDesc(C.Structure):
    """struct dd_affine_bwd_desc"""
    _fields_ = [("w", c_void_p), ("gain", c_void_p), ("dout", c_void_p), ("dweff", c_void_p), ("rowscale", c_void_p),
                ("O", c_int), ("I", c_int), ("groups", c_int), ("normalize", c_int)]


OPTIM_MAX_EMA = 4
GNORM_CHUNK = 8192


class GnormDesc(C.Structure):
    """struct dd_gnorm_desc"""
    _fields_ = [("g", c_void_p), ("numel", C.c_longlong), ("chunk_begin", c_int), ("_pad", c_int)]


class OptimDesc(C.Structure):
    """struct dd_optim_desc"""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("ema", c_void_p * OPTIM_MAX_EMA),
                ("numel", C.c_longlong), ("rows", c_int), ("row_len", c_int), ("normalize", c_int), ("row_begin", c_int)]


class OptimHyper(C.Structure):
    """struct dd_optim_hyper"""
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("bias_correction1", C.c_double), ("bias_correction2", C.c_double),
                ("ema_beta", C.c_double * OPTIM_MAX_EMA), ("feedback_beta", C.c_double * OPTIM_MAX_EMA),
                ("ema_is_f64", c_int * OPTIM_MAX_EMA), ("n_ema", c_int)]


EPI_NONE, EPI_SCALE_SILU, EPI_RESIDUAL = 0, 1, 2
EPI2_NONE, EPI2_SILU, EPI2_SCALE, EPI2_RAW = 0, 1, 2, 3
WFMT_BF16_OTI, WFMT_F32_OIT = 0, 1
WPERM_NONE, WPERM_QK, WPERM_QKV = 0, 1, 2

_SIGNATURES = {
    "dd_last_error": (C.c_char_p, []),
    "dd_abi_version": (c_int, []),
    "dd_device_info": (c_int, [C.POINTER(c_int)] * 3),
    "dd_weight_prep": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float, c_int, c_int,
                               c_int, c_int, c_void_p]),
    "dd_mpconv_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  C.POINTER(ConvEpilogue), c_void_p]),
    "dd_mpconv_forward_cat": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_mpconv_forward_naive": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p]),
    "dd_stem_patches": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_stem_patches_cols": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_conv_out": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                            c_int, c_int, c_void_p]),
    "dd_noise_embedding": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_float,
                                   c_void_p, c_int, c_int, c_void_p]),
    "dd_emb_affine": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "dd_label_embedding": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                   c_int, c_void_p]),
    "dd_sigma_logvar": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "dd_mp_fourier": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dd_axpby": (c_int, [c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_long, c_void_p]),
    "dd_pixnorm_silu": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_int, c_void_p]),
    "dd_cat_silu": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_int,
                            c_int, c_int, c_void_p]),
    "dd_avgpool2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_attention_qkv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_stft_mel": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_void_p, c_int, c_void_p]),
    "dd_fgla_istft": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int,
                              c_int, c_void_p, c_int, c_void_p]),
    "dd_fgla_stft_update": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_void_p, c_float, c_int, c_void_p]),
    "dd_ola_finalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dd_attention_axis": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_attention_train": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_mpconv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                c_int, c_void_p]),
    "dd_weight_transpose": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_weight_prep_bwd": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dd_weight_prep_batched": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dd_weight_normalize_batched": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dd_weight_transpose_batched": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dd_silu_scale_bwd": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_long, c_int,
                                  c_void_p]),
    "dd_pixnorm_silu_bwd": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_long, c_int, c_void_p]),
    "dd_cat_silu_bwd": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_int,
                                c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_enc_grad_combine": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    "dd_attn_in_bwd": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_long, c_int, c_void_p]),
    "dd_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "dd_emb_affine_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dd_noise_embedding_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_float,
                                       c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dd_label_embedding_bwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p]),
    "dd_sigma_logvar_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "dd_head_grad": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             c_void_p]),
    "dd_weight_prep_z2": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float, c_int, c_void_p]),
    "dd_reflect_fill_w": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_dae_stem": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_up2_silu_pad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_conv5x5_out": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_ddec_stem": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_void_p]),
    "dd_avgpool2_pad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_ddec_head": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_q4_stem": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_void_p]),
    "dd_gemm_f32": (c_int, [c_void_p, c_long, c_long, c_long, c_void_p, c_long, c_long, c_long, c_void_p, c_long, c_long, c_long,
                            c_int, c_int, c_int, c_int, c_int, c_int, c_long, c_float, c_float, c_float, c_int, c_void_p]),
    "dd_frame_reflect": (c_int, [c_void_p, c_void_p, c_int, c_long, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_complex_abs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "dd_mdct_ola": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dd_mel_linearize": (c_int, [c_void_p, c_void_p, c_long, c_float, c_float, c_void_p]),
    "dd_dae_enc_patches": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_dae_latents_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_pack_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_unpack_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_patches5x5": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_conv5x5_dense": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_mel_blend": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    "dd_mdct_phase_psd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_void_p,
                                  c_void_p, c_void_p]),
    "dd_sampler_cfg_lerp": (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_int, c_long, c_void_p]),
    "dd_conv_trace_read": (c_int, [c_void_p, c_int, c_void_p]),
    "dd_roll_pad_w": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_int, c_void_p]),
    "dd_crop_unroll_w": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_void_p]),
    "dd_stereo_fix_noise": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_long, c_void_p]),
    "dd_grad_norm_clip": (c_int, [c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p]),
    "dd_optim_step_batched": (c_int, [c_void_p, c_int, c_int, C.POINTER(OptimHyper), c_void_p, c_void_p]),
    "dd_sampler_update": (c_int, [c_void_p, c_void_p, c_float, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                  c_int, c_long, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building first if needed) the shared library; raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    # build() is a no-op when the stamp matches the sources (a stale library after a csrc/ or header edit would
    # otherwise load silently); without nvcc a library that is already there is used as it is
    try:
        path = _build.build()
    except RuntimeError:
        if not os.path.exists(_build.LIB_PATH):
            raise
        path = _build.LIB_PATH
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.dd_abi_version() != 1:
        raise RuntimeError(f"dualdiffusion_b200: ABI mismatch ({lib.dd_abi_version()} != 1); rebuild the library")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().dd_last_error().decode(errors="replace")
        raise RuntimeError(f"dualdiffusion_b200 C-ABI call failed ({status}): {msg}")


_last_device: Optional[torch.device] = None      # device of the tensor most recently passed through ptr()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    global _last_device
    if t is None:
        return None
    if t.is_cuda:
        _last_device = t.device
    return t.data_ptr()


def stream_ptr() -> int:
    """Current stream of the current device.  Every wrapper passes its tensors through ptr() before it asks for the
    stream (argument order), and the C ABI never switches devices: tensors on another device than the current one would
    be launched into the wrong context, so that case fails here, loudly."""
    if _last_device is not None and _last_device.index is not None and _last_device.index != torch.cuda.current_device():
        raise RuntimeError(f"dualdiffusion_b200: tensors live on {_last_device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device(...) / set_device")
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dualdiffusion_b200 has no CPU path: tensors must live on a CUDA device")
