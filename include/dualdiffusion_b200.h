/*
 * dualdiffusion_b200 — C ABI of the B200-native (sm_100a) denoising hot path.
 *
 * The reference (parlance-zz/dualdiffusion) is pure Python/PyTorch: it has no FFI of its own, and
 * every entry point below replaces a *PyTorch library call site* of the reference, cited per
 * function as <file>:<line> relative to /root/reference/src.  The reference-side binding is a
 * ctypes stub (see INTEGRATION.md and dualdiffusion_b200/_lib.py); the functions are called from
 * the drop-in module classes in dualdiffusion_b200/modules, never by reference code directly.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; dd_last_error() then returns a
 *     thread-local, NUL-terminated description (Python shim raises RuntimeError with it);
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call only
 *     enqueues work on that stream and never synchronises, so calls are CUDA-graph capturable;
 *   - activations are NHWC ("channels_last") bf16: [B][H][W][C], C contiguous;
 *   - model input/output latents keep the reference's NCHW fp32 layout.
 */
#ifndef DUALDIFFUSION_B200_H
#define DUALDIFFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_ABI_VERSION 1

#if defined(__GNUC__)
#define DD_API __attribute__((visibility("default")))
#else
#define DD_API
#endif

/* ---- library ------------------------------------------------------------------------------ */
DD_API const char* dd_last_error(void);
DD_API int dd_abi_version(void);
/* Fills SM count and compute capability of the current device. */
DD_API int dd_device_info(int* num_sms, int* cc_major, int* cc_minor);

/* ---- weight preparation: MPConv.forward weight path, modules/mp_tools.py:359-364 ----------- */
/* w      : [O][I_g][taps] weights as stored by the reference (OIHW, fp32 or bf16)
 * out    : DD_WFMT_BF16_OTI : bf16 [O'][taps][I_g]  (operand layout of dd_mpconv_forward)
 *          DD_WFMT_F32_OIT  : fp32 [O][I_g][taps]   (same order as the input; normalize_weights)
 * math   : w_o <- w_o / (1e-4 + ||w_o||_2 / sqrt(fan_in))   if normalize != 0  (training mode)
 *          w_o <- w_o * gain_host * (*gain_dev) / sqrt(fan_in),  fan_in = I_g * taps
 * gain_dev may be NULL (== 1).  perm = DD_WPERM_QK de-interleaves the attn_qk output channels:
 * reference row (head*2*d + 2*c + j) -> row (j*O/2 + head*d + c)  (unet_edm2_b4.py:137-138),
 * so q and k come out of the GEMM as two contiguous [heads][d] halves.                          */
#define DD_WFMT_BF16_OTI 0
#define DD_WFMT_F32_OIT 1
#define DD_WPERM_NONE 0
#define DD_WPERM_QK 1
#define DD_WPERM_QKV 2   /* row (head*3*d + 3*c + j) -> row (j*O/3 + head*d + c)  (old/unet_edm2_ddec_mdct_b3.py:149-150) */
/* out_row_stride (elements, 0 = dense I_g*taps) lets rows be written into a wider, zero-initialised buffer
 * (K padded to 64 for the stem GEMM; Cout padded to 16 rows for the head GEMM).                    */
DD_API int dd_weight_prep(const void* w, int w_is_bf16, void* out, int out_format, int O, int I_g, int taps,
                          const float* gain_dev, float gain_host, int normalize, int perm, int head_dim,
                          int out_row_stride, void* stream);

/* ---- MPConv forward: F.conv2d in MPConv.forward, modules/mp_tools.py:369 ------------------- */
/* Stride-1, zero "same" padding, ksize in {1,3}, Cin/groups and Cout/groups multiples of 32.
 * Fused epilogue (Block.forward, modules/unets/unet_edm2_b4.py:119-131,150-157):
 *   DD_EPI_NONE       : out = acc
 *   DD_EPI_SCALE_SILU : out = mp_silu(acc * scale[b][c])                  (:121-122, :150-151)
 *   DD_EPI_RESIDUAL   : out = clip(alpha*acc + beta*residual[b][h][w][c]) (mp_sum :131,:154; clip :157)
 * optional second output from the same (pre-rounding) value:
 *   DD_EPI2_SILU      : out2 = mp_silu(out)                               (next conv_res0 input, :119)
 *   DD_EPI2_SCALE     : out2 = out * scale2[b][c]                         (attn_qk input, :135-136)
 *   DD_EPI2_RAW       : out2 = acc                                        (train mode: saved for backward) */
#define DD_EPI_NONE 0
#define DD_EPI_SCALE_SILU 1
#define DD_EPI_RESIDUAL 2
#define DD_EPI2_NONE 0
#define DD_EPI2_SILU 1
#define DD_EPI2_SCALE 2
#define DD_EPI2_RAW 3    /* out2 = acc (pre-activation, saved for the backward pass of DD_EPI_SCALE_SILU) */
typedef struct dd_conv_epilogue {
    int mode;              /* DD_EPI_*  */
    int mode2;             /* DD_EPI2_* */
    float alpha, beta;     /* DD_EPI_RESIDUAL */
    float clip;            /* <= 0: no clip */
    const void* scale;     /* fp32 [B][Cout] */
    const void* scale2;    /* fp32 [B][Cout] */
    const void* residual;  /* bf16 [B][H][W][Cout] */
    void* out2;            /* bf16 [B][H][W][Cout] */
} dd_conv_epilogue;
DD_API int dd_mpconv_forward(const void* x, const void* w_prepped, void* out, int B, int H, int W, int Cin, int Cout,
                      int ksize, int groups, const dd_conv_epilogue* epilogue_host, void* stream);

/* 1x1 MPConv over the channel concatenation [x1 | x2] (bf16 NHWC, C1 and C2 channels) without writing the concatenation:
 * conv_skip of a decoder block after mp_cat (unet_edm2_b4.py:295-296, :129).  w_prepped: bf16 [Cout][C1 + C2] with the mp_cat
 * weights folded into its columns; C1 a multiple of 64 (32 when C1 + C2 is not a multiple of 64).  No epilogue.        */
DD_API int dd_mpconv_forward_cat(const void* x1, int C1, const void* x2, int C2, const void* w_prepped, void* out, int B,
                                 int H, int W, int Cout, void* stream);

/* Scalar CUDA-core convolution with the same contract as dd_mpconv_forward (no epilogue).
 * Debug/triangulation aid for the tests only; never called by the product path.                 */
DD_API int dd_mpconv_forward_naive(const void* x, const void* w_prepped, void* out, int B, int H, int W, int Cin, int Cout,
                            int ksize, int groups, void* stream);

/* ---- UNet stem and head: modules/unets/unet_edm2_b4.py:258-271 and :290-294 ---------------- */
/* Stem input cat(c_in(sigma)*x_in, ones, ln_freqs) (:262-271) written as zero-padded 3x3 patches
 * out[b][h][w][tap*(Cin+2)+c] (bf16, 64 columns) so that conv_in is dd_mpconv_forward(ksize=1, Cin=64)
 * with weights from dd_weight_prep(DD_WFMT_BF16_OTI, out_row_stride=64).                            */
DD_API int dd_stem_patches(const float* x_in_nchw, const float* sigma, float sigma_data, const float* ln_freqs_h,
                           void* out_patches, int B, int Cin, int H, int W, void* stream);
/* Same with `cols` (64 or 128) patch columns per pixel: 8-channel latents (unet_edm2_b4_2.py: 9 * (8 + 2) = 90 columns). */
DD_API int dd_stem_patches_cols(const float* x_in_nchw, const float* sigma, float sigma_data, const float* ln_freqs_h,
                                void* out_patches, int B, int Cin, int H, int W, int cols, void* stream);
/* D = c_skip(sigma)*x_in + c_out(sigma)*conv_out(x) on the tensor cores; w_prepped16: bf16 [16][9][C]
 * (rows >= Cout zero) from dd_weight_prep(gain = out_gain).  x_ref (optional, [B][Cout+1][H][W] fp32):
 * D = mp_sum(x_ref[:, :-1], D, t = x_ref[:, -1:])  (:293-294).  Result fp32 NCHW.                     */
DD_API int dd_conv_out(const void* x_nhwc, const void* w_prepped16, const float* x_in_nchw, const float* sigma,
                       float sigma_data, const float* x_ref_nchw, float* d_out_nchw, int B, int C, int H, int W,
                       int Cout, void* stream);

/* ---- embeddings: unet_edm2_b4.py:232-238, :273-276; MPFourier mp_tools.py:324-330 ---------- */
/* emb[b] = mp_silu(mp_sum(W_noise @ fourier(ln(sigma_b)/4) / sqrt(cnoise), label_emb[b], t))      */
DD_API int dd_noise_embedding(const float* sigma, const float* freqs, const float* phases, int cnoise, const void* w_noise,
                       int w_is_bf16, int normalize, const float* label_emb, float label_balance, float* emb_out,
                       int B, int cemb, void* stream);
/* Batched per-block embedding projections (emb_linear / emb_linear_qk / emb_linear_v, :121,:135,:150):
 *   out_j[b][o] = bias_j + gain_j * sum_i W_j[o][i] * emb[b][(o / (O_j/groups_j)) * I_j + i] / sqrt(I_j)
 * (+ per-row weight normalisation when normalize != 0).  descs_dev is a DEVICE array.             */
typedef struct dd_affine_desc {
    const void* w;       /* [O][I] fp32 or bf16 */
    const float* gain;   /* device scalar, may be NULL (== 1) */
    float* out;          /* [B][O] fp32 */
    int O, I, groups, w_is_bf16;
    float bias;
    int normalize;
} dd_affine_desc;
DD_API int dd_emb_affine(const dd_affine_desc* descs_dev, int n_descs, int max_O, const float* emb, int B, int cemb,
                  void* stream);

/* UNet.get_embeddings (:232-235): out[b] = mp_sum(W_u*1, W_l @ normalize(emb_in[b or 0]) / sqrt(I), t = mask[b]);
 * emb_in fp32 [Bc][I] with Bc in {1, Bm}; mask fp32 [Bm] (1 = conditioned, 0 = unconditional).     */
DD_API int dd_label_embedding(const float* emb_in, int Bc, int I, const void* w_label, const void* w_uncond,
                              int w_is_bf16, const float* mask, int Bm, int normalize, float* out, int cemb,
                              void* stream);
/* MPFourier.forward, 1-D input (mp_tools.py:324-330): out[i][c] = cos(x_i*freqs_c + phases_c)*sqrt(2)   */
DD_API int dd_mp_fourier(const float* x, int count, const float* freqs, const float* phases, int n, float* out,
                         void* stream);
/* UNet.get_sigma_loss_logvar (:237-238): out[i] = w . fourier(ln(sigma_i)/4) / sqrt(n)             */
DD_API int dd_sigma_logvar(const float* sigma, int count, const float* freqs, const float* phases, int n,
                           const void* w, int w_is_bf16, float* out, void* stream);

/* ---- elementwise block glue ---------------------------------------------------------------- */
/* pixel norm (mp_tools.py:42-49 with dim=1) + mp_silu: x = t/(1e-4+rms_c(t)); s = mp_silu(x)      */
DD_API int dd_pixnorm_silu(const void* t, void* x_out, void* s_out, long npix, int C, void* stream);
/* decoder input: xcat = [wa * up(a), wb * b] (mp_cat mp_tools.py:294-301; nearest x2 :79), s = mp_silu(xcat).
 * b may be NULL (Cb = 0); xcat_out may be NULL when only s is needed; H,W are OUTPUT sizes.        */
DD_API int dd_cat_silu(const void* a, int Ca, const void* b, int Cb, float wa, float wb, int upsample, void* xcat_out,
                void* s_out, int B, int H, int W, void* stream);
/* 2x2 mean pooling (mp_tools.py:77), H,W are INPUT sizes (even).                                  */
DD_API int dd_avgpool2(const void* x, void* out, int B, int H, int W, int C, void* stream);

/* out = clip(alpha*a + beta*b) on bf16 activations (mp_sum mp_tools.py:274-279 with float t).        */
DD_API int dd_axpby(const void* a, const void* b, float alpha, float beta, float clip, void* out, long n, void* stream);

/* ---- attention: unet_edm2_b4.py:137-151 ---------------------------------------------------- */
/* q|k halves [B][N][2C] (after DD_WPERM_QK), v [B][N][C]; per-head cosine normalisation of q,k,v over
 * head_dim (eps 1e-4), softmax(q k^T / sqrt(head_dim)) v, then out = mp_silu(y * scale_v[b][c]).     */
DD_API int dd_attention(const void* qk, const void* v, const float* scale_v, void* out, int B, int N, int heads,
                 int head_dim, void* stream);

/* Fused-projection variant of the newer lineage (unet_edm2_b4_2.py:146-156): qkv [B][N][3C] holds the q | k | v thirds
 * (after DD_WPERM_QKV); cosine normalisation of q, k, v and softmax(q k^T / sqrt(head_dim)) v as above, no gain and no
 * activation on the result: out_raw [B][N][C].                                                                        */
DD_API int dd_attention_qkv(const void* qkv, void* out_raw, int B, int N, int heads, int head_dim, void* stream);

/* Train-mode variant: additionally writes the pre-activation attention output a = softmax(qk^T/8) v
 * (bf16 [B][N][C]) that dd_attention_bwd and dd_silu_scale_bwd need.                                     */
DD_API int dd_attention_train(const void* qk, const void* v, const float* scale_v, void* out, void* raw_out, int B, int N,
                              int heads, int head_dim, void* stream);

/* Axis ("separable") attention of the legacy ddec UNets, modules/unets/old/unet_edm2_ddec_mdct_b3.py:144-163:
 * qkv [B][Z][H][W][3C] (channels_last_3d, q|k|v thirds after DD_WPERM_QKV), attention over H (axis 0, sequences
 * (b,z,w)) or W (axis 1, sequences (b,z,h)); out [B][Z][H][W][C] = mp_silu(attention).  The reference's
 * permute -> reshape(b*z*w, heads, d, 3, h) -> SDPA -> reshape -> permute round trip is folded into the kernel's
 * token / sequence strides: no tensor is physically permuted.                                               */
DD_API int dd_attention_axis(const void* qkv, void* out, int B, int Z, int H, int W, int heads, int head_dim, int axis,
                             void* stream);

/* ---- mel-STFT encode / FGLA decode: modules/formats/old/spectrogram.py:176-238 ------------- */
/* Shared conventions: n_fft in {6400, 4096} (compile-time mixed-radix plans), win_length == n_fft, center=True with
 * reflect padding, one-sided spectra of n_fft/2+1 bins.  `window` fp32 [n_fft]; `twiddles` = exp(-2 pi i m/(n_fft/2)),
 * m < n_fft/2, and `twiddles_half` = exp(-2 pi i k/n_fft), k <= n_fft/2, as interleaved (re,im) fp32 pairs.
 *
 * dd_stft_mel: SpectrogramConverter.audio_to_spectrogram (:176-179) + FrequencyScale.scale
 * (frequency_scale.py:127-128) + the affine of raw_to_sample (:223-226):
 *   out[s][f][t] = ((sum_k |STFT(raw_s)[k][t]| * fb[k][f]) ** exponent - mean) * scale
 * with the triangular filterbank given per filter f as a run of fb_count[f] weights starting at bin fb_start[f]
 * (weights at fb_weight[fb_offset[f] ...]).  raw [S][len] fp32, out [S][n_filters][n_frames], n_frames = 1+len/hop.
 * Live format MS_MDCT_DualFormat.raw_to_mel_spec (formats/ms_mdct_dual.py:230-257): pass the second (narrower)
 * window as `window2` and per-bin coefficients; the magnitude fed to the filterbank is then
 * |STFT_w1|[k]*coef1[k] + |STFT_w2|[k]*coef2[k] (window normalisation, blend weight and 1/mel-density folded into
 * coef1/coef2).  window2, coef1, coef2 may be NULL (single window, coefficient 1).                              */
DD_API int dd_stft_mel(const float* raw, int n_signals, int len, const float* window, const float* window2,
                       const float* coef1, const float* coef2, const float* twiddles,
                       const float* twiddles_half, int n_fft, int hop, const int* fb_start, const int* fb_count,
                       const int* fb_offset, const float* fb_weight, int n_filters, float exponent, float mean,
                       float scale, float* out, int n_frames, void* stream);
/* griffinlim (old/phase_recovery.py:40-129), one iteration = dd_fgla_istft + dd_fgla_stft_update.
 * State T [S][n_frames][bins] complex64 (frame-major, private layout); magnitudes mag_tk [S][n_frames][bins] fp32.
 * dd_fgla_istft (:84-95, :121-124): X = T/(|T|+1e-16) * M with M = mag (stereo == 0) or the stereo-coherence blend
 *   merged + max(interp_t,0)*(mag - merged), merged = (mag_s + mag_{s^1})/2; state == NULL means angles == 1 (:73).
 *   Writes the windowed overlap-add of all inverse frames into ola [S][ola_len] (zeroed by the call),
 *   ola_len = n_fft + hop*(n_frames-1); dividing by the window envelope and trimming n_fft/2 is left to the reader.
 * dd_fgla_stft_update (:97-117): rebuilt = STFT(ola/env trimmed); T <- rebuilt - momentum*T (first != 0: T <- rebuilt).
 * dd_ola_finalize: out[s][j] = ola[s][n_fft/2 + j] / env[n_fft/2 + j], j < len = hop*(n_frames-1).          */
DD_API int dd_fgla_istft(const float* state, const float* mag_tk, int n_signals, int n_frames, int stereo,
                         float interp_t, const float* window, const float* twiddles, const float* twiddles_half,
                         int n_fft, int hop, float* ola, int ola_len, void* stream);
DD_API int dd_fgla_stft_update(const float* ola, const float* env, int n_signals, int n_frames, int len,
                               const float* window, const float* twiddles, const float* twiddles_half, int n_fft,
                               int hop, float* state, float momentum, int first, void* stream);
DD_API int dd_ola_finalize(const float* ola, const float* env, int n_signals, int ola_len, int n_fft, int len,
                           float* out, void* stream);

/* ---- EDM sampler step glue: pipelines/dual_diffusion_pipeline.py:699-737 -------------------- */
/* n = element count of ONE latent batch (B*C*H*W); d_2b holds [cond ; uncond] = 2n elements.
 * cfg = lerp(D[B:], D[:B], cfg_scale); x_hat = lerp(cfg, sample, t_hat)   (:701, :712).
 * dup bit 0 writes x_hat twice ([x_hat ; x_hat], the `.repeat(2,1,1,1)` of :712) into a 2n buffer;
 * dup bit 1 = unconditional module (unet_class_embeddings is None, :703-704, :719-720): d holds n elements, cfg = D.   */
DD_API int dd_sampler_cfg_lerp(const float* d_2b, const float* sample, float cfg_scale, float t_hat, float* cfg_out,
                               float* x_hat_out, int dup, long n, void* stream);
/* cfg2 = lerp(D2[B:], D2[:B], cfg_scale); cfg = use_heun ? lerp(cfg1, cfg2, .5) : cfg1;
 * sample = lerp(cfg, sample, t) + p * noise   (:717-724, :734-737); noise may be NULL (p = 0).
 * dup != 0: sample_inout is a 2n buffer, read from its first half and written to both halves (:661).   */
DD_API int dd_sampler_update(const float* cfg1, const float* d2_2b, float cfg_scale, int use_heun, float t, float p,
                             const float* noise, float* sample_inout, float* cfg_out, int dup, long n, void* stream);

/* ==== backward pass of the UNet train step ======================================================
 * Reference: loss.backward() (training/trainer.py:1022-1044) through UNet.forward
 * (modules/unets/unet_edm2_b4.py:250-296), Block.forward (:110-158) and MPConv.forward
 * (modules/mp_tools.py:357-373).  The reference relies on PyTorch autograd (cuDNN dgrad/wgrad, SDPA backward,
 * eager elementwise backward kernels); each entry point below replaces one of those autograd nodes.
 * dgrad of an MPConv is dd_mpconv_forward itself on weights re-laid-out by dd_weight_transpose.            */

/* conv2d weight gradient (autograd of F.conv2d, mp_tools.py:369) on the tensor cores:
 *   dw[co][tap][ci] (fp32, the layout of dd_weight_prep's DD_WFMT_BF16_OTI output) = scale * sum_pix dy[pix][co] * x[pix+tap][ci]
 * x bf16 [B][H][W][Cin], dy bf16 [B][H][W][Cout]; accumulate != 0 adds into dw.                             */
DD_API int dd_mpconv_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                           int groups, float scale, int accumulate, void* stream);
/* bf16 [Cout][taps][cin_g] (dd_weight_prep output) -> bf16 [Cin][taps][cout_g] with the taps reversed: the operand
 * that turns dd_mpconv_forward(dy, ., Cin<->Cout swapped) into the data gradient of the convolution.      */
DD_API int dd_weight_transpose(const void* w_prepped, void* out, int Cout, int cin_g, int taps, int groups, void* stream);

/* dd_weight_prep (DD_WFMT_BF16_OTI output) for a whole parameter set in one launch: one CTA per weight row of every
 * descriptor.  Used by the train step, where every MPConv weight changes every optimizer step (mp_tools.py:359-364 runs
 * inside each forward there).                                                                                         */
typedef struct dd_wprep_desc {
    const void* w;       /* parameter [O][I_g][taps], fp32 or bf16 */
    void* out;           /* bf16 [O (or more, zero-filled by the caller)][row_stride] */
    const float* gain;   /* device scalar or NULL */
    float gain_host;
    int w_is_bf16, O, I_g, taps, normalize, perm, head_dim, row_stride;
    int row_begin;       /* exclusive prefix sum of O over the descriptor array */
} dd_wprep_desc;
DD_API int dd_weight_prep_batched(const dd_wprep_desc* descs_dev, int n_descs, int total_rows, void* stream);
/* normalize_weights() (mp_tools.py:375-378; trainer.py:1107-1108 runs it after every optimizer step) for a whole parameter
 * set in one launch, IN PLACE on fp32 parameters (descs[].w; .out / gain / perm fields unused).                          */
DD_API int dd_weight_normalize_batched(const dd_wprep_desc* descs_dev, int n_descs, int total_rows, void* stream);
/* dd_weight_transpose for a whole parameter set in one launch (32x32 tiles; tile_begin = exclusive prefix sum of
 * ceil(cin_g/32)*ceil(cout_g/32)*groups*taps).                                                                       */
typedef struct dd_wtrans_desc {
    const void* src;     /* bf16 [groups*cout_g][taps][cin_g] */
    void* dst;           /* bf16 [groups*cin_g][taps][cout_g], taps reversed */
    int cout_g, cin_g, taps, groups;
    int tile_begin;
} dd_wtrans_desc;
DD_API int dd_weight_transpose_batched(const dd_wtrans_desc* descs_dev, int n_descs, int total_tiles, void* stream);

/* Backward of dd_weight_prep (mp_tools.py:359-364), batched over parameters: one CTA per weight row.
 *   dw[o][i][tap] (=|+=) d(w_eff)/d(w) applied to dweff;  *dgain += <dweff, w_hat>/sqrt(fan_in) * gain_host       */
typedef struct dd_wbwd_desc {
    const float* w;      /* parameter, fp32 [O][I_g][taps] */
    const float* dweff;  /* fp32 gradient of the effective weight, element (row, tap*I_g + i), rows permuted like dd_weight_prep */
    float* dw;           /* fp32 [O][I_g][taps] */
    const float* gain;   /* device scalar or NULL */
    float* dgain;        /* device scalar accumulator or NULL */
    float gain_host;
    int O, I_g, taps, normalize, perm, head_dim, row_stride, accumulate;
    int row_begin;       /* exclusive prefix sum of O over the descriptor array */
    int t_cout_g;        /* 0: dweff as above.  > 0 (= O / groups): dweff holds the transposed, tap-reversed gradient
                          * [groups*I_g][taps][t_cout_g] that dd_mpconv_wgrad produces when called with x and dy exchanged
                          * (M = input channels): used for grouped layers with I_g > O/groups, where the 128-row MMA tile
                          * then straddles half as many groups of the block diagonal */
} dd_wbwd_desc;
DD_API int dd_weight_prep_bwd(const dd_wbwd_desc* descs_dev, int n_descs, int total_rows, void* stream);

/* y = mp_silu(pre*scale[b][c]) (conv_res0 epilogue :121-122, attention tail :150-151):
 *   dpre = coef*dy*silu'(pre*scale)*scale;  dscale[b][c] += sum_pix coef*dy*silu'(pre*scale)*pre.  bf16 [B][npix][C]. */
DD_API int dd_silu_scale_bwd(const void* dy, float coef, const void* pre, const float* scale, void* dpre, float* dscale,
                             int B, long npix, int C, void* stream);
/* encoder pixel-norm + mp_silu (:114-119): dt0 from g (block-output gradient, enters as ca*g through mp_sum) and ds. */
DD_API int dd_pixnorm_silu_bwd(const void* g, float ca, const void* ds, const void* t0, void* dt0, long npix, int C,
                               void* stream);
/* backward of dd_cat_silu: dxc = c1*d_xc + d_s*silu'(xc); da = mask(|a_prev|<clip)*wa*sum_2x2(dxc[:Ca]); db = wb*dxc[Ca:].
 * H,W are the sizes of xc; a_prev (the tensor that was passed as `a`) may be NULL (no clip mask).              */
DD_API int dd_cat_silu_bwd(const void* d_xc, float c1, const void* d_s, const void* xc, const void* a_prev, float clip,
                           float wa, float wb, int upsample, void* da, void* db, int B, int H, int W, int Ca, int Cb,
                           void* stream);
/* encoder chain: out = mask(|x_prev|<clip) * ((down ? avg_pool2d backward of dx0 : dx0) + dskip); H,W,C of x_prev.      */
DD_API int dd_enc_grad_combine(const void* dx0, int down, const void* dskip, const void* x_prev, float clip, void* out,
                               int B, int H, int W, int C, void* stream);
/* attention input (:133-136): dx2 = ca*g3 + dxv + dxs*c_qk[b][c]; dc_qk[b][c] += sum_pix dxs*x2.               */
DD_API int dd_attn_in_bwd(const void* g3, float ca, const void* dxv, const void* dxs, const void* x2, const float* c_qk,
                          void* dx2, float* dc_qk, int B, long npix, int C, void* stream);
/* SDPA + cosine-norm backward (:137-148): inputs as dd_attention_train, d_a = gradient of raw_out;
 * dqk [B][N][2C], dv [B][N][C] bf16; stats_ws fp32 [B][heads][N][2] scratch.                                */
DD_API int dd_attention_bwd(const void* qk, const void* v, const void* a_raw, const void* d_a, void* dqk, void* dv,
                            float* stats_ws, int B, int N, int heads, int head_dim, void* stream);

/* emb_linear* backward (batched like dd_emb_affine), two stages over a descriptor array:
 *   stage 1 (emb != NULL):  dweff_j[o][i] = sum_b dout_j[b][o]*emb[b][g*I+i]  and the row scales (rowscale: fp32 [O] scratch);
 *   stage 2 (demb != NULL): demb[b][g*I+i] += sum_o dout_j[b][o]*w_eff_j[o][i]   (needs stage 1 of the same descriptors).
 * Either pointer may be NULL to run one stage only: the train step runs stage 1 per gradient bucket (a block's embedding
 * weights travel with the block's bucket) and stage 2 once at the end.                                                  */
typedef struct dd_affine_bwd_desc {
    const float* w;      /* [O][I] fp32 */
    const float* gain;   /* device scalar or NULL */
    const float* dout;   /* [B][O] fp32 */
    float* dweff;        /* [O][I] fp32 out */
    float* rowscale;     /* [O] fp32 scratch */
    int O, I, groups, normalize;
} dd_affine_bwd_desc;
DD_API int dd_emb_affine_bwd(const dd_affine_bwd_desc* descs_dev, int n_descs, int max_O, int max_cols, const float* emb,
                             float* demb, int B, int cemb, void* stream);
/* backward of dd_noise_embedding: dweff [cemb][cnoise] (emb_noise), dlabel [B][cemb] (gradient of `embeddings`).  */
DD_API int dd_noise_embedding_bwd(const float* sigma, const float* freqs, const float* phases, int cnoise,
                                  const float* w_noise, int normalize, const float* label_emb, float label_balance,
                                  const float* demb, float* dweff, float* dlabel, int B, int cemb, void* stream);
/* backward of dd_label_embedding: dweff_label [cemb][I], dweff_uncond [cemb].                                  */
DD_API int dd_label_embedding_bwd(const float* emb_in, int Bc, int I, const float* mask, int Bm, const float* dout,
                                  float* dweff_label, float* dweff_uncond, int cemb, void* stream);
/* backward of dd_sigma_logvar: dw[n] (=|+=) sum_i dout[i]*fourier(ln(sigma_i)/4)/sqrt(n).                       */
DD_API int dd_sigma_logvar_bwd(const float* sigma, int count, const float* freqs, const float* phases, int n,
                               const float* dout, float* dw, int accumulate, void* stream);
/* UNet head (:290-294): dF = dD*c_out(sigma) (times the x_ref blend weight), NCHW fp32 -> NHWC bf16 padded to Cpad channels. */
DD_API int dd_head_grad(const float* dD_nchw, const float* sigma, float sigma_data, const float* x_ref_nchw, void* dF_nhwc,
                        int B, int Cout, int H, int W, int Cpad, void* stream);

/* ==== DAE_D3 diffusion-autoencoder decoder (modules/daes/dae_edm2_d3.py) ==========
 * Stereo depth Z = 2 of the reference's channels_last_3d tensors is folded into the channel dimension
 * ([B][H][Wp][2C] bf16, channel = z*C + c) so that every MPConv3D (:43-93) runs on dd_mpconv_forward; reflection
 * padding along W (:63-66) is physical: tensors carry pw halo columns per side (Wp = W + 2*pw).               */

/* MPConv3D weight path, eval mode (:70-81), into the folded layout.  w: [O][I][kz][taps] fp32 or bf16.
 *   kz == 2 -> bf16 [2*O][taps][i_stride], element (z'*O+o, tap, z*I+i) = scale*w[o][i][z xor z'][tap]  (dense conv)
 *   kz == 1 -> bf16 [2*O][taps][i_stride], element (z'*O+o, tap, i)     = scale*w[o][i][0][tap]         (groups = 2)
 * scale = gain_host * (*gain_dev) / sqrt(I*kz*taps); columns >= the input channel count are zero.               */
DD_API int dd_weight_prep_z2(const void* w, int w_is_bf16, void* out, int O, int I, int kz, int taps, const float* gain_dev,
                             float gain_host, int i_stride, void* stream);
/* ReflectionPad3d along W (:64): x[.., pw-k, :] = x[.., pw+k, :], x[.., pw+W-1+k, :] = x[.., pw+W-1-k, :], k = 1..pw. */
DD_API int dd_reflect_fill_w(void* x, int B, int H, int Wp, int C, int pw, void* stream);
/* DAE_D3.decode input (:358-359): latents fp32 NCHW (B, 2L, H, W) -> [B][H][W+2pw][Cpad] bf16, channel z*(L+1)+c with
 * the constant-one channel at c == L, zero above 2(L+1), halo columns mirrored.                                 */
DD_API int dd_dae_stem(const float* latents, void* out, int B, int L, int H, int W, int pw, int Cpad, void* stream);
/* Block.forward head of an "up" block (:188-189,196): xc = resample_3d(a, "up") (nearest x2 in H, W), s = mp_silu(xc);
 * a [B][Ha][Wa+2pw][C] -> xc, s [B][2Ha][2Wa+2pw][C], halo columns mirrored.                                    */
DD_API int dd_up2_silu_pad(const void* a, void* xc, void* s, int B, int Ha, int Wa, int C, int pw, void* stream);
/* conv_out (:368): MPConv3D (1,5,5), C -> 1 per stereo side, times *gain_dev.  x [B][H][W+2pw][2C] bf16 (pw >= 2),
 * w25 fp32 [C][25] pre-scaled by 1/sqrt(25 C); out fp32 NCHW (B, 2, H, W) = tensor_5d_to_4d of the reference.   */
DD_API int dd_conv5x5_out(const void* x, const float* w25, const float* gain_dev, float* out, int B, int H, int W, int C,
                          int pw, void* stream);

/* ==== diffusion-decoder UNet DDec_MCLT_UNet_B1 (modules/unets/unet_edm2_ddec_mclt_b1.py:278-326),
 * same folded-stereo / halo-column layout as the DAE decoder above ============================================== */

/* Network input (:294-309): per stereo side [c_in(sigma)*x_in, the k PSD bins of the mel row (x_ref view + permute of
 * :294-295), 1] -> [B][F][W+2pw][Cpad] bf16, channel z*(k+2)+c.  x_in fp32 (B,2,F,W), x_ref fp32 (B,2,F*k,W).      */
DD_API int dd_ddec_stem(const float* x_in, const float* x_ref, const float* sigma, float sigma_data, void* out, int B, int F,
                        int W, int k, int pw, int Cpad, void* stream);
/* resample_3d "down" (mp_tools.py:85-90) on a W-padded tensor: [B][H][W+2pw][C] -> [B][H/2][W/2+2pw][C].          */
DD_API int dd_avgpool2_pad(const void* x, void* out, int B, int H, int W, int C, int pw, void* stream);
/* Output head (:323-326): D = c_skip*x_in + c_out*F with F = channel z of the Cst-wide conv_out result
 * [B][H][W+2pw][Cst] bf16; x_in, out fp32 (B,2,H,W).                                                              */
DD_API int dd_ddec_head(const void* f, const float* x_in, const float* sigma, float sigma_data, float* out, int B, int H,
                        int W, int pw, int Cst, void* stream);

/* unet_edm2_q4_ddec.UNet input (modules/unets/unet_edm2_q4_ddec.py:268-277): mp_cat(c_in*x_in, permuted PSD x_ref) as
 * NHWC bf16 [B][F][W][Cpad]: [wa*c_in*x (C) | wb*x_ref[b][c][h*k+j] at channel C + j*C + c | 1 | 0...]; the constant-one
 * channel carries conv_in's bias through a centre-tap weight column.  x_in fp32 (B,C,F,W), x_ref fp32 (B,C,F*k,W).   */
DD_API int dd_q4_stem(const float* x_in, const float* x_ref, const float* sigma, float sigma_data, float wa, float wb,
                      void* out, int B, int C, int F, int W, int k, int Cpad, void* stream);

/* ==== MDCT side of the live format (utils/mclt.py:87-130, modules/formats/ms_mdct_dual.py:259-318).
 * The MCLT / inverse MCLT of all frames is one fp32 library GEMM against a host-built matrix (window, phase shifts,
 * 1/mel-density and scales folded in); these kernels do the framing, |.|, overlap-add and mel linearisation around it. */

/* Batched fp32 GEMM of the audio-format transforms, C[m][n] = sum_k A[m][k] * B[k][n] (fp32 FMA, CUDA cores; the
 * reference computes these with torch.matmul / torch.fft in fp32: utils/mclt.py:87-130, ms_mdct_dual.py:259-318,
 * frequency_scale.py:130-142).  Operands are addressed by element strides (A: a_m, a_k; B: b_k, b_n; C: c_m, c_n; plus a
 * batch stride each).  gather_hop > 0: B is not a matrix but the reflect-padded frames of a raw signal of gather_len
 * samples per batch item, B[k][n] = raw[reflect(n * gather_hop + k - gather_pad)] (mclt.py:90-95).  a_pow > 0: every A
 * element is read as clip(a * a_scale + a_offset, 0) ** a_pow (the mel "unscale").  relu != 0: C = max(C, 0).          */
DD_API int dd_gemm_f32(const float* a, long a_m, long a_k, long a_batch, const float* b, long b_k, long b_n, long b_batch,
                       float* c, long c_m, long c_n, long c_batch, int M, int N, int K, int batch, int gather_hop,
                       int gather_pad, long gather_len, float a_scale, float a_offset, float a_pow, int relu, void* stream);

/* mclt.py:90-95: out[s][t][n] = reflect-padded raw[s][t*hop + n - pad_left]; raw [S][L] fp32, out [S][T][block_width]. */
DD_API int dd_frame_reflect(const float* raw, float* out, int S, long L, int T, int block_width, int hop, int pad_left,
                            void* stream);
/* ms_mdct_dual.py:300-306: |re + i im| * scale; y [S][2N][T] (real rows, then imaginary rows) -> out [S][N][T].       */
DD_API int dd_complex_abs(const float* y, float* out, int S, int N, int T, float scale, void* stream);
/* mclt.py:123-130: 50 %-overlap add of the inverse frames y [S][T][2N] with the first / last half block cropped ->
 * out [S][(T-1)*N].                                                                                                   */
DD_API int dd_mdct_ola(const float* y, float* out, int S, int T, int N, void* stream);
/* ms_mdct_dual.py:261-265: out = clip(mel - offset, 0) ** inv_exponent.                                               */
DD_API int dd_mel_linearize(const float* mel, float* out, long n, float offset, float inv_exponent, void* stream);

/* MS_MDCT_DualFormat, second lineage (modules/formats/ms_mdct_dual_2.py).  raw_to_mel_spec (:198-216): the three per-window
 * mel spectrograms (dd_stft_mel, one window each) are blended per mel filter and compressed:
 *   out[s][f][t] = ((sum_i mels[i][s][f][t] * ww[f][i]) ** exponent + offset) * inv_scale,  mels [n_win][S][F][T].       */
DD_API int dd_mel_blend(const float* mels, const float* ww, int n_win, int S, int F, int T, float exponent, float offset,
                        float inv_scale, float* out, void* stream);
/* raw_to_mdct_phase_psd (:275-289) on the unscaled MCLT rows y [S][2N][T] (real rows, then imaginary rows):
 *   phase = clip(re / max(|z|, 1e-20), -1, 1) * phase_mul,  psd = ((|z| * inv_density[k]) ** exponent + offset) * inv_scale. */
DD_API int dd_mdct_phase_psd(const float* y, const float* inv_density, int S, int N, int T, float exponent, float offset,
                             float inv_scale, float phase_mul, float* phase, float* psd, void* stream);

/* DAE_D3.encode (modules/daes/dae_edm2_d3.py:342-354).  Input assembly for conv_in (1,5,5): per stereo side the 5x5 patch
 * of [mel, 1] (reflection along W, zeros along H), 50 values padded to 64 -> [B][H][W+2pw][128] bf16; mel fp32 (B,2,H,W). */
DD_API int dd_dae_enc_patches(const float* mel, void* out, int B, int H, int W, int pw, void* stream);
/* Tail: conv_latents_out result [B][H][W+2pw][Cst] bf16 (channel z*L+c) -> tensor_5d_to_4d + avg_pool2d(ratio) -> fp32 NCHW
 * (B, 2L, H/ratio, W/ratio), channel c*2+z.                                                                              */
DD_API int dd_dae_latents_pool(const void* f, float* out, int B, int L, int H, int W, int pw, int Cst, int ratio, void* stream);

/* dae_edm2_q4.DAE (modules/daes/dae_edm2_q4.py:170-300): 2-D zero-padded MPConv network; the blocks run on
 * dd_mpconv_forward / dd_pixnorm_silu / dd_avgpool2 / dd_cat_silu, these four entry points are its two ends.
 * conv_latents_in input (:292): fp32 NCHW (B, C, H, W) -> bf16 NHWC [B][H][W][Cpad]; channel `ones_channel` (>= C, or < 0
 * for none) is the constant 1 that multiplies the bias column of the prepared weight (mp_tools.py:370-371).          */
DD_API int dd_pack_nhwc(const float* x, void* out, int B, int C, int H, int W, int Cpad, int ones_channel, void* stream);
/* conv_latents_out result (:281): bf16 NHWC [B][H][W][Cpad] -> fp32 NCHW (B, C, H, W), the first C channels.        */
DD_API int dd_unpack_nchw(const void* x, float* out, int B, int C, int H, int W, int Cpad, void* stream);
/* conv_in (5,5) with bias (:232) as a K = cols GEMM: x fp32 (B, C, H, W) -> bf16 [B][H][W][cols], column tap*C + c the
 * zero-padded 5x5 patch (tap = ky*5 + kx), column 25*C the constant 1, zero above; cols 64 or 128.                   */
DD_API int dd_patches5x5(const float* x, void* out, int B, int C, int H, int W, int cols, void* stream);
/* conv_out (5,5), C -> Cout <= 4 dense, zero padding, times *gain_dev (:299): x bf16 [B][H][W][C], w fp32 [Cout][25][C]
 * pre-scaled by 1/sqrt(25 C), out fp32 (B, Cout, H, W).                                                              */
DD_API int dd_conv5x5_dense(const void* x, const float* w, const float* gain_dev, float* out, int B, int H, int W, int C,
                            int Cout, void* stream);

/* Diagnostic (tuning only, not on the product path): with DD_CONV_TRACE=1 in the environment when the library is loaded,
 * 3x3 layers of >= 8 rows run a traced instantiation of the halo kernel that records, for the first 4 CTAs and their
 * first 64 tiles, clock64 stamps of the three warp roles (8 slots: producer stage-free / issued, MMA accumulator-free /
 * operands-landed / issued, epilogue accumulator-complete / done / started waiting).  Copies the last traced launch's
 * stamps [4][8][64] and meta {num_tiles, grid, n_tile, a_stages, nbuf, epilogue warps, kchunks, staged} to the host.   */
DD_API int dd_conv_trace_read(unsigned long long* stamps_host, int n_stamps, int* meta_host);

/* ---- sampler-loop options (pipelines/dual_diffusion_pipeline.py) ------------------------------------------------ */
/* seamless_loop, :651-656: out[c][r][j] = x[r][(j - pad - shift) mod W] for j in [0, W + 2 pad) and c in [0, copies):
 * torch.roll(x, shift, -1) -> cat(x[..., -pad:], x, x[..., :pad]) (-> .repeat(copies, 1, 1, 1)).  x fp32 [rows][W].   */
DD_API int dd_roll_pad_w(const float* x, float* out, long rows, int W, int shift, int pad, int copies, void* stream);
/* seamless_loop, :729-732: out[r][i] = xp[r][pad + (i + shift) mod W] = torch.roll(xp[..., pad:-pad], -shift, -1).     */
DD_API int dd_crop_unroll_w(const float* xp, float* out, long rows, int W, int shift, int pad, void* stream);
/* stereo_fix, :638-640: noise[:, ::2] = noise[:, 1::2]; out = mp_sum(fresh, noise, t) (mp_tools.py:273-279).
 * noise, fresh, out fp32 (B, C, hw), C even; out may alias fresh (not noise).                                          */
DD_API int dd_stereo_fix_noise(const float* noise, const float* fresh, float t, float* out, int B, int C, long hw,
                               void* stream);

/* ==== Optimizer-side sweep of the train step: clip_grad_norm_ (training/trainer.py:1044), torch AdamW
 * step (trainer.py:461-473,1062), EMA / feedback-EMA lerps (training/ema.py:284-313) and normalize_weights
 * (trainer.py:1107-1108, modules/mp_tools.py:375-378) in two launches over a parameter set.                          */

/* Global L2 norm of a list of fp32 gradient tensors and the clip coefficient of torch.nn.utils.clip_grad_norm_:
 * out[0] = ||g||_2, out[1] = min(1, max_norm / (||g|| + 1e-6)) (1 if max_norm <= 0).  Deterministic: one partial per
 * 8192-element chunk (partials_dev holds total_chunks floats), summed in fixed order in fp64.                        */
#define DD_GNORM_CHUNK 8192
typedef struct dd_gnorm_desc {
    const float* g;
    long long numel;
    int chunk_begin;     /* exclusive prefix sum of ceil(numel / DD_GNORM_CHUNK) over the descriptor array */
    int _pad;
} dd_gnorm_desc;
DD_API int dd_grad_norm_clip(const dd_gnorm_desc* descs_dev, int n_descs, int total_chunks, float* partials_dev,
                             float max_norm, float* out_norm_coef_dev, void* stream);

/* One launch, one CTA per row: g' = g * coef (coef = norm_coef_dev[1], or 1 if NULL; g itself is not rewritten);
 * AdamW (decoupled decay) on p, m, v; for k < n_ema in order: ema_k = lerp(ema_k, p, 1 - ema_beta[k]) and, if
 * feedback_beta[k] >= 0, p = lerp(p, ema_k, 1 - feedback_beta[k]); finally, if normalize, every row of p viewed as
 * [rows][row_len] is divided by (1e-4 + ||row|| / sqrt(row_len)).  The EMA copies receive the un-normalised post-step
 * weights, as in the reference's order of calls.  Tensors without weight-norm are cut into rows of any convenient
 * length with normalize = 0 (the last row may be short: numel bounds it).                                            */
#define DD_OPTIM_MAX_EMA 4
typedef struct dd_optim_desc {
    float* p;            /* parameter, fp32, updated in place */
    const float* g;      /* gradient, fp32; NULL = no gradient this step: AdamW is skipped (m, v unused), the EMA
                          * lerps and the re-normalisation still run, as in the reference's separate sweeps */
    float* m;            /* exp_avg */
    float* v;            /* exp_avg_sq */
    void* ema[DD_OPTIM_MAX_EMA];   /* EMA copies of p (fp32, or fp64 where dd_optim_hyper.ema_is_f64[k]); NULL = skip */
    long long numel;
    int rows, row_len, normalize;
    int row_begin;       /* exclusive prefix sum of rows over the descriptor array */
} dd_optim_desc;
typedef struct dd_optim_hyper {
    double lr, beta1, beta2, eps, weight_decay;
    double bias_correction1, bias_correction2;     /* 1 - beta^step, computed by the caller for this step */
    double ema_beta[DD_OPTIM_MAX_EMA];
    double feedback_beta[DD_OPTIM_MAX_EMA];        /* < 0: no feedback from this EMA */
    int ema_is_f64[DD_OPTIM_MAX_EMA];
    int n_ema;
} dd_optim_hyper;
DD_API int dd_optim_step_batched(const dd_optim_desc* descs_dev, int n_descs, int total_rows,
                                 const dd_optim_hyper* hyper_host, const float* norm_coef_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUALDIFFUSION_B200_H */
