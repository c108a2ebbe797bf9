// HBM-bound glue kernels of the EDM2 UNet hot path: weight preparation, UNet stem/head convolutions
// (tiny channel counts, CUDA cores), embeddings, pixel-norm / concat / resample fusions and the EDM
// sampler step.  All of them are coalesced 128-bit-vector streaming kernels; reference call sites are
// cited in include/dualdiffusion_b200.h.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>

namespace {

constexpr float kNormEps = 1e-4f;   // modules/mp_tools.py:43

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

template <bool kBf16>
__device__ __forceinline__ float load_w(const void* w, size_t i) {
    if constexpr (kBf16) return __bfloat162float(static_cast<const __nv_bfloat16*>(w)[i]);
    else return static_cast<const float*>(w)[i];
}

// ------------------------------------------------------------------------------------------
// weight prep
// ------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void weight_prep_kernel(const void* __restrict__ w, void* __restrict__ out, int out_format, int O, int I_g,
                                   int taps, const float* __restrict__ gain_dev, float gain_host, int normalize,
                                   int perm, int head_dim, int row_stride) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ float red[32];
    const int o = blockIdx.x;
    const int fan_in = I_g * taps;
    const size_t base = (size_t)o * fan_in;
    float scale = gain_host * (gain_dev ? *gain_dev : 1.f) * rsqrtf((float)fan_in);
    if (normalize) {
        float ss = 0.f;
        for (int i = threadIdx.x; i < fan_in; i += blockDim.x) {
            const float v = load_w<kBf16>(w, base + i);
            ss += v * v;
        }
        ss = block_sum(ss, red);
        scale /= (kNormEps + sqrtf(ss) * rsqrtf((float)fan_in));
    }
    int o_dst = o;
    if (perm == DD_WPERM_QK) {
        const int head = o / (2 * head_dim), rem = o % (2 * head_dim);
        o_dst = (rem & 1) * (O / 2) + head * head_dim + (rem >> 1);
    } else if (perm == DD_WPERM_QKV) {
        const int head = o / (3 * head_dim), rem = o % (3 * head_dim);
        o_dst = (rem % 3) * (O / 3) + head * head_dim + rem / 3;
    }
    if (out_format == DD_WFMT_BF16_OTI) {
        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(out) + (size_t)o_dst * row_stride;
        for (int j = threadIdx.x; j < fan_in; j += blockDim.x) {      // j = tap * I_g + i  (coalesced writes)
            const int tap = j / I_g, i = j - tap * I_g;
            dst[j] = __float2bfloat16_rn(load_w<kBf16>(w, base + (size_t)i * taps + tap) * scale);
        }
    } else {
        float* dst = static_cast<float*>(out) + (size_t)o_dst * row_stride;
        for (int j = threadIdx.x; j < fan_in; j += blockDim.x) dst[j] = load_w<kBf16>(w, base + j) * scale;
    }
}

// ------------------------------------------------------------------------------------------
// pixel norm + mp_silu  (one warp per pixel, channels in registers)
// ------------------------------------------------------------------------------------------
constexpr int kMaxVecPerLane = 10;    // C <= 10*32*8 = 2560

// NV = 16-byte vectors per lane (C <= NV * 256) as a template parameter: compiled for the widest layer only (10), a
// 256-channel launch executed ten predicated copies of every loop body.
template <int NV>
__global__ void pixnorm_silu_kernel(const uint4* __restrict__ t, uint4* __restrict__ x_out, uint4* __restrict__ s_out,
                                    long npix, int C) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const int lane = threadIdx.x & 31;
    const long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= npix) return;
    const int nvec = C >> 3;
    const uint4* src = t + pix * nvec;
    uint4 reg[NV];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + k * 32;
        if (v < nvec) {
            reg[k] = __ldg(src + v);
            const uint32_t u[4] = {reg[k].x, reg[k].y, reg[k].z, reg[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(u[j]); ss += f.x * f.x + f.y * f.y; }
        }
    }
    ss = warp_sum(ss);
    const float inv = 1.f / (kNormEps + sqrtf(ss) * rsqrtf((float)C));
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + k * 32;
        if (v < nvec) {
            const uint32_t u[4] = {reg[k].x, reg[k].y, reg[k].z, reg[k].w};
            uint32_t xo[4], so[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = unpack_bf16x2(u[j]);
                f.x *= inv; f.y *= inv;
                xo[j] = pack_bf16x2(f.x, f.y);
                so[j] = pack_bf16x2(mp_silu_fast(f.x), mp_silu_fast(f.y));
            }
            x_out[pix * nvec + v] = make_uint4(xo[0], xo[1], xo[2], xo[3]);
            s_out[pix * nvec + v] = make_uint4(so[0], so[1], so[2], so[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// decoder input: (optional nearest x2 upsample of a) ++ (optional skip b), scaled, plus mp_silu
// ------------------------------------------------------------------------------------------
// I = index type: 32-bit whenever the tensor has fewer than 2^31 vectors (64-bit div / mod by a run-time value is a ~100
// instruction sequence, and this kernel does up to five per 16-byte vector: with `long` it was bound by them, not by HBM)
template <typename I>
__global__ void cat_silu_kernel(const uint4* __restrict__ a, int va, const uint4* __restrict__ b, int vb, float wa,
                                float wb, int up, uint4* __restrict__ xcat, uint4* __restrict__ s, int B, int H, int W) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const I vt = (I)(va + vb);
    const I total = (I)B * (I)H * (I)W * vt;
    for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
        const I pix = idx / vt;
        const int v = (int)(idx - pix * vt);
        uint4 q;
        float sc;
        if (v < va) {
            I apix = pix;
            if (up) {
                const I r = pix / (I)W;
                const int w = (int)(pix - r * (I)W);
                const I bb = r / (I)H;
                const int h = (int)(r - bb * (I)H);
                apix = (bb * (I)(H >> 1) + (I)(h >> 1)) * (I)(W >> 1) + (I)(w >> 1);
            }
            q = __ldg(a + (size_t)apix * va + v);
            sc = wa;
        } else {
            q = __ldg(b + (size_t)pix * vb + (v - va));
            sc = wb;
        }
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
        uint32_t xo[4], so[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_bf16x2(u[j]);
            f.x *= sc; f.y *= sc;
            xo[j] = pack_bf16x2(f.x, f.y);
            so[j] = pack_bf16x2(mp_silu_fast(f.x), mp_silu_fast(f.y));
        }
        if (xcat) xcat[idx] = make_uint4(xo[0], xo[1], xo[2], xo[3]);
        s[idx] = make_uint4(so[0], so[1], so[2], so[3]);
    }
}

template <typename I>
__global__ void avgpool2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int nvec) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const int Ho = H >> 1, Wo = W >> 1;
    const I total = (I)B * (I)Ho * (I)Wo * (I)nvec;
    for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
        const I pix = idx / (I)nvec;
        const int v = (int)(idx - pix * (I)nvec);
        const I row = pix / (I)Wo;
        const int w = (int)(pix - row * (I)Wo);
        const I b = row / (I)Ho;
        const int h = (int)(row - b * (I)Ho);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const uint4 q = __ldg(x + (((size_t)b * H + 2 * h + dy) * W + 2 * w + dx) * nvec + v);
                const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(u[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
            }
        out[idx] = make_uint4(pack_bf16x2(acc[0] * .25f, acc[1] * .25f), pack_bf16x2(acc[2] * .25f, acc[3] * .25f),
                              pack_bf16x2(acc[4] * .25f, acc[5] * .25f), pack_bf16x2(acc[6] * .25f, acc[7] * .25f));
    }
}

// ------------------------------------------------------------------------------------------
// UNet stem: preconditioning + constant/positional channels, written as 3x3 patches ("im2col") so that
// conv_in runs as a K=64 GEMM on the tensor cores.  patch[pix][tap*CT + c], zero padded to 64 columns.
// ------------------------------------------------------------------------------------------
template <typename I>
__global__ void stem_patches_kernel(const float* __restrict__ x_in, const float* __restrict__ sigma, float sigma_data,
                                    const float* __restrict__ ln_freqs, uint4* __restrict__ out, int B, int Cin, int H,
                                    int W, int vecs) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const int CT = Cin + 2;
    const I total = (I)B * (I)H * (I)W * (I)vecs;    // vecs x (8 bf16) per pixel: 64 or 128 patch columns
    for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
        const I pix = idx / (I)vecs;
        const int v = (int)(idx - pix * (I)vecs);
        const I row = pix / (I)W;
        const int w = (int)(pix - row * (I)W);
        const int b = (int)(row / (I)H);
        const int h = (int)(row - (I)b * (I)H);
        const float sg = __ldg(sigma + b);
        const float c_in = rsqrtf(sigma_data * sigma_data + sg * sg);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = v * 8 + j;
            float val = 0.f;
            if (k < 9 * CT) {
                const int tap = k / CT, c = k - tap * CT;
                const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                    if (c < Cin) val = c_in * __ldg(x_in + (((size_t)b * Cin + c) * H + hh) * W + ww);
                    else if (c == Cin) val = 1.f;
                    else val = __ldg(ln_freqs + hh);
                }
            }
            f[j] = val;
        }
        out[idx] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7]));
    }
}

// ------------------------------------------------------------------------------------------
// embeddings
// ------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void noise_embedding_kernel(const float* __restrict__ sigma, const float* __restrict__ freqs,
                                       const float* __restrict__ phases, int cnoise, const void* __restrict__ w,
                                       int normalize, const float* __restrict__ label, float t, float* __restrict__ out,
                                       int cemb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    extern __shared__ float four[];
    const int b = blockIdx.y;
    const float c_noise = logf(sigma[b]) * 0.25f;
    for (int i = threadIdx.x; i < cnoise; i += blockDim.x)
        four[i] = cosf(c_noise * freqs[i] + phases[i]) * 1.41421356237f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= cemb) return;
    float acc = 0.f, ss = 0.f;
    for (int i = lane; i < cnoise; i += 32) {
        const float wv = load_w<kBf16>(w, (size_t)o * cnoise + i);
        acc += wv * four[i];
        ss += wv * wv;
    }
    acc = warp_sum(acc);
    ss = warp_sum(ss);
    if (lane == 0) {
        float scale = rsqrtf((float)cnoise);
        if (normalize) scale /= (kNormEps + sqrtf(ss) * rsqrtf((float)cnoise));
        const float e = acc * scale;
        const float l = label[(size_t)b * cemb + o];
        const float m = (e + t * (l - e)) * rsqrtf((1.f - t) * (1.f - t) + t * t);
        out[(size_t)b * cemb + o] = mp_silu_f(m);
    }
}

// One warp per output row.  The launch streams every emb_linear* weight of the UNet (a quarter of its parameters) for a few
// rows of emb, so it is bound by reading them: 16-byte loads (8 bf16 / 2 x 4 fp32 per lane and trip) whenever the rows allow.
template <bool kBf16>
__device__ __forceinline__ void load_w8(const void* w, size_t i, float (&f)[8]) {      // i % 8 == 0, 16-byte aligned
    if constexpr (kBf16) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(w) + i));
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16x2(u[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(w) + i));
        const float4 b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(w) + i) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
}

template <bool kBf16>
__device__ __forceinline__ void emb_affine_row(const dd_affine_desc& d, const float* __restrict__ e0, int o, int lane, int B,
                                               int cemb, float gain) {
    const bool vec = d.I % 8 == 0 && (reinterpret_cast<uintptr_t>(d.w) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(e0) & 15u) == 0 && cemb % 4 == 0;
    for (int b0 = 0; b0 < B; b0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float ss = 0.f;
        if (vec) {
            for (int i = lane * 8; i < d.I; i += 256) {
                float wv[8];
                load_w8<kBf16>(d.w, (size_t)o * d.I + i, wv);
#pragma unroll
                for (int k = 0; k < 8; ++k) ss += wv[k] * wv[k];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (b0 + j < B) {
                        const float4* e = reinterpret_cast<const float4*>(e0 + (size_t)(b0 + j) * cemb + i);
                        const float4 ea = __ldg(e), eb = __ldg(e + 1);
                        acc[j] += wv[0] * ea.x + wv[1] * ea.y + wv[2] * ea.z + wv[3] * ea.w + wv[4] * eb.x + wv[5] * eb.y +
                                  wv[6] * eb.z + wv[7] * eb.w;
                    }
            }
        } else {
            for (int i = lane; i < d.I; i += 32) {
                const float wv = load_w<kBf16>(d.w, (size_t)o * d.I + i);
                ss += wv * wv;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (b0 + j < B) acc[j] += wv * e0[(size_t)(b0 + j) * cemb + i];
            }
        }
        ss = warp_sum(ss);
        float scale = gain * rsqrtf((float)d.I);
        if (d.normalize) scale /= (kNormEps + sqrtf(ss) * rsqrtf((float)d.I));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = warp_sum(acc[j]);
            if (lane == 0 && b0 + j < B) d.out[(size_t)(b0 + j) * d.O + o] = d.bias + a * scale;
        }
    }
}

__global__ void emb_affine_kernel(const dd_affine_desc* __restrict__ descs, const float* __restrict__ emb, int B,
                                  int cemb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const dd_affine_desc d = descs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= d.O) return;
    const int g = o / (d.O / d.groups);
    const float* e0 = emb + (size_t)g * d.I;
    const float gain = d.gain ? *d.gain : 1.f;
    if (d.w_is_bf16) emb_affine_row<true>(d, e0, o, lane, B, cemb, gain);
    else emb_affine_row<false>(d, e0, o, lane, B, cemb, gain);
}


// ------------------------------------------------------------------------------------------
// class-label embedding (UNet.get_embeddings) and loss log-variance head (get_sigma_loss_logvar)
// ------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void label_embedding_kernel(const float* __restrict__ emb_in, int Bc, int I, const void* __restrict__ w_label,
                                       const void* __restrict__ w_uncond, const float* __restrict__ mask, int normalize,
                                       float* __restrict__ out, int cemb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int bm = blockIdx.y;
    if (o >= cemb) return;
    const float* e = emb_in + (size_t)(Bc == 1 ? 0 : bm) * I;
    float acc = 0.f, ee = 0.f, ww = 0.f;
    for (int i = lane; i < I; i += 32) {
        const float wv = load_w<kBf16>(w_label, (size_t)o * I + i);
        const float ev = e[i];
        acc += wv * ev; ee += ev * ev; ww += wv * wv;
    }
    acc = warp_sum(acc); ee = warp_sum(ee); ww = warp_sum(ww);
    if (lane == 0) {
        const float rs = rsqrtf((float)I);
        float c = acc * rs / (kNormEps + sqrtf(ee) * rs);           // normalize(emb_in) then W/sqrt(I)
        float u = load_w<kBf16>(w_uncond, o);                          // fan_in 1: W * 1
        if (normalize) {
            c /= (kNormEps + sqrtf(ww) * rs);
            u /= (kNormEps + fabsf(u));
        }
        const float t = mask[bm];
        out[(size_t)bm * cemb + o] = (u + t * (c - u)) * rsqrtf((1.f - t) * (1.f - t) + t * t);
    }
}

template <bool kBf16>
__global__ void logvar_kernel(const float* __restrict__ sigma, const float* __restrict__ freqs,
                              const float* __restrict__ phases, int n, const void* __restrict__ w,
                              float* __restrict__ out, int count) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= count) return;
    const float c = logf(sigma[idx]) * 0.25f;
    float acc = 0.f;
    for (int i = lane; i < n; i += 32)
        acc += load_w<kBf16>(w, i) * cosf(c * freqs[i] + phases[i]) * 1.41421356237f;
    acc = warp_sum(acc);
    if (lane == 0) out[idx] = acc * rsqrtf((float)n);
}


__global__ void mp_fourier_kernel(const float* __restrict__ x, const float* __restrict__ freqs,
                                  const float* __restrict__ phases, int n, float* __restrict__ out, long total) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % n);
        out[i] = cosf(x[i / n] * freqs[c] + phases[c]) * 1.41421356237f;
    }
}

// out = clip(alpha*a + beta*b)   (mp_sum / lerp on bf16 activations)
__global__ void axpby_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, float alpha, float beta,
                             float clip, uint4* __restrict__ out, long nvec) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
        const uint4 qa = __ldg(a + i), qb = __ldg(b + i);
        const uint32_t ua[4] = {qa.x, qa.y, qa.z, qa.w}, ub[4] = {qb.x, qb.y, qb.z, qb.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack_bf16x2(ua[j]), fb = unpack_bf16x2(ub[j]);
            const float r0 = fminf(fmaxf(alpha * fa.x + beta * fb.x, -clip), clip);
            const float r1 = fminf(fmaxf(alpha * fa.y + beta * fb.y, -clip), clip);
            o[j] = pack_bf16x2(r0, r1);
        }
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------------
// EDM sampler step (fp32 state)
// ------------------------------------------------------------------------------------------
__global__ void sampler_cfg_lerp_kernel(const float4* __restrict__ d, const float4* __restrict__ sample, float cfg,
                                        float t_hat, float4* __restrict__ cfg_out, float4* __restrict__ xhat, int dup,
                                        long n4) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 c = d[i], u = (dup & 2) ? c : d[i + n4], s = sample[i];      // bit 1: unconditional (no CFG pair)
        float4 o, x;
        o.x = u.x + cfg * (c.x - u.x); o.y = u.y + cfg * (c.y - u.y);
        o.z = u.z + cfg * (c.z - u.z); o.w = u.w + cfg * (c.w - u.w);
        x.x = o.x + t_hat * (s.x - o.x); x.y = o.y + t_hat * (s.y - o.y);
        x.z = o.z + t_hat * (s.z - o.z); x.w = o.w + t_hat * (s.w - o.w);
        cfg_out[i] = o;
        if (xhat) {
            xhat[i] = x;
            if (dup & 1) xhat[i + n4] = x;
        }
    }
}

__global__ void sampler_update_kernel(const float4* __restrict__ cfg1, const float4* __restrict__ d2, float cfg,
                                      int use_heun, float t, float p, const float4* __restrict__ noise,
                                      float4* __restrict__ sample, float4* __restrict__ cfg_out, int dup, long n4) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 o = cfg1[i];
        if (use_heun) {
            const float4 c = d2[i], u = (dup & 2) ? c : d2[i + n4];
            o.x = o.x + 0.5f * ((u.x + cfg * (c.x - u.x)) - o.x);
            o.y = o.y + 0.5f * ((u.y + cfg * (c.y - u.y)) - o.y);
            o.z = o.z + 0.5f * ((u.z + cfg * (c.z - u.z)) - o.z);
            o.w = o.w + 0.5f * ((u.w + cfg * (c.w - u.w)) - o.w);
        }
        float4 s = sample[i];
        s.x = o.x + t * (s.x - o.x); s.y = o.y + t * (s.y - o.y);
        s.z = o.z + t * (s.z - o.z); s.w = o.w + t * (s.w - o.w);
        if (noise) {
            const float4 nz = noise[i];
            s.x += p * nz.x; s.y += p * nz.y; s.z += p * nz.z; s.w += p * nz.w;
        }
        sample[i] = s;
        if (dup & 1) sample[i + n4] = s;
        if (cfg_out) cfg_out[i] = o;
    }
}

// ------------------------------------------------------------------------------------------
// naive reference convolution (tests only)
// ------------------------------------------------------------------------------------------
__global__ void conv_naive_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                  __nv_bfloat16* __restrict__ out, int B, int H, int W, int Cin, int Cout, int ks,
                                  int groups) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    const long total = (long)B * H * W * Cout;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int co = (int)(idx % Cout);
        const long pix = idx / Cout;
        const int wq = (int)(pix % W), h = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
        const int g = co / cout_g;
        float acc = 0.f;
        for (int tap = 0; tap < ks * ks; ++tap) {
            const int hh = h + tap / ks - ks / 2, ww = wq + tap % ks - ks / 2;
            if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const __nv_bfloat16* xp = x + (((long)b * H + hh) * W + ww) * Cin + g * cin_g;
            const __nv_bfloat16* wp = w + ((size_t)co * ks * ks + tap) * cin_g;
            for (int i = 0; i < cin_g; ++i) acc += __bfloat162float(xp[i]) * __bfloat162float(wp[i]);
        }
        out[idx] = __float2bfloat16_rn(acc);
    }
}

inline int grid_for(long total, int block, int cap_mult = 8) {
    const long blocks = (total + block - 1) / block;
    return (int)std::max<long>(1, std::min<long>(blocks, (long)dd_num_sms() * cap_mult));
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int dd_weight_prep(const void* w, int w_is_bf16, void* out, int out_format, int O, int I_g, int taps,
                              const float* gain_dev, float gain_host, int normalize, int perm, int head_dim,
                              int out_row_stride, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(w && out && O > 0 && I_g > 0 && taps > 0, "dd_weight_prep: bad arguments");
    DD_REQUIRE(out_format == DD_WFMT_BF16_OTI || out_format == DD_WFMT_F32_OIT, "dd_weight_prep: bad out_format");
    DD_REQUIRE(perm == DD_WPERM_NONE || (perm == DD_WPERM_QK && head_dim > 0 && O % (2 * head_dim) == 0) ||
                   (perm == DD_WPERM_QKV && head_dim > 0 && O % (3 * head_dim) == 0),
               "dd_weight_prep: bad permutation arguments");
    const int row_stride = out_row_stride > 0 ? out_row_stride : I_g * taps;
    DD_REQUIRE(row_stride >= I_g * taps, "dd_weight_prep: out_row_stride smaller than a row");
    if (w_is_bf16)
        DD_CHECK_CUDA(dd_launch_pdl(weight_prep_kernel<true>, dim3(O), dim3(128), 0, stream, w, out, out_format, O, I_g, taps, gain_dev, gain_host, normalize,
                                                        perm, head_dim, row_stride));
    else
        DD_CHECK_CUDA(dd_launch_pdl(weight_prep_kernel<false>, dim3(O), dim3(128), 0, stream, w, out, out_format, O, I_g, taps, gain_dev, gain_host,
                                                         normalize, perm, head_dim, row_stride));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_pixnorm_silu(const void* t, void* x_out, void* s_out, long npix, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(t && x_out && s_out, "dd_pixnorm_silu: null pointer");
    DD_REQUIRE(C % 8 == 0 && C <= kMaxVecPerLane * 256, "dd_pixnorm_silu: C=%d unsupported", C);
    if (npix == 0) return 0;
    const int warps = 8;
    const int nv = ceil_div(C / 8, 32);
    auto kernel = nv <= 1 ? pixnorm_silu_kernel<1> : nv <= 2 ? pixnorm_silu_kernel<2> : nv <= 3 ? pixnorm_silu_kernel<3>
                : nv <= 4 ? pixnorm_silu_kernel<4> : nv <= 5 ? pixnorm_silu_kernel<5> : nv <= 6 ? pixnorm_silu_kernel<6>
                : nv <= 8 ? pixnorm_silu_kernel<8> : pixnorm_silu_kernel<kMaxVecPerLane>;
    DD_CHECK_CUDA(dd_launch_pdl(kernel, dim3((unsigned)((npix + warps - 1) / warps)), dim3(warps * 32), 0, stream,
                                static_cast<const uint4*>(t), static_cast<uint4*>(x_out), static_cast<uint4*>(s_out), npix,
                                C));
    return 0;
}

extern "C" int dd_cat_silu(const void* a, int Ca, const void* b, int Cb, float wa, float wb, int upsample, void* xcat_out,
                           void* s_out, int B, int H, int W, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(a && s_out && Ca > 0 && Ca % 8 == 0 && Cb % 8 == 0, "dd_cat_silu: bad arguments");
    DD_REQUIRE(Cb == 0 || b, "dd_cat_silu: skip pointer missing");
    DD_REQUIRE(!upsample || (H % 2 == 0 && W % 2 == 0), "dd_cat_silu: upsample needs even output size");
    const long total = (long)B * H * W * ((Ca + Cb) / 8);
    if (total == 0) return 0;
    if (total < (1l << 31) - (long)grid_for(total, 256) * 256)
        DD_CHECK_CUDA(dd_launch_pdl(cat_silu_kernel<unsigned>, dim3(grid_for(total, 256)), dim3(256), 0, stream,
                                    static_cast<const uint4*>(a), Ca / 8, static_cast<const uint4*>(b), Cb / 8, wa, wb,
                                    upsample, static_cast<uint4*>(xcat_out), static_cast<uint4*>(s_out), B, H, W));
    else
        DD_CHECK_CUDA(dd_launch_pdl(cat_silu_kernel<long>, dim3(grid_for(total, 256)), dim3(256), 0, stream,
                                    static_cast<const uint4*>(a), Ca / 8, static_cast<const uint4*>(b), Cb / 8, wa, wb,
                                    upsample, static_cast<uint4*>(xcat_out), static_cast<uint4*>(s_out), B, H, W));
    return 0;
}

extern "C" int dd_avgpool2(const void* x, void* out, int B, int H, int W, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "dd_avgpool2: bad arguments");
    const long total = (long)B * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return 0;
    const bool narrow = total < (1l << 31) - (long)grid_for(total, 256) * 256;      // idx + stride must not wrap
    DD_CHECK_CUDA(dd_launch_pdl(narrow ? avgpool2_kernel<unsigned> : avgpool2_kernel<long>, dim3(grid_for(total, 256)), dim3(256), 0, stream,
                                static_cast<const uint4*>(x), static_cast<uint4*>(out), B, H, W, C / 8));
    return 0;
}

extern "C" int dd_stem_patches_cols(const float* x_in, const float* sigma, float sigma_data, const float* ln_freqs,
                                    void* out, int B, int Cin, int H, int W, int cols, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x_in && sigma && ln_freqs && out, "dd_stem_patches: null pointer");
    DD_REQUIRE(cols == 64 || cols == 128, "dd_stem_patches: patch width %d unsupported (64 or 128)", cols);
    DD_REQUIRE(9 * (Cin + 2) <= cols, "dd_stem_patches: in_channels=%d unsupported (9*(Cin+2) must be <= %d)", Cin, cols);
    const long total = (long)B * H * W * (cols / 8);
    if (total == 0) return 0;
    const bool narrow = total < (1l << 31) - (long)grid_for(total, 256) * 256;
    DD_CHECK_CUDA(dd_launch_pdl(narrow ? stem_patches_kernel<unsigned> : stem_patches_kernel<long>, dim3(grid_for(total, 256)), dim3(256), 0, stream, x_in, sigma,
                                sigma_data, ln_freqs, static_cast<uint4*>(out), B, Cin, H, W, cols / 8));
    return 0;
}

extern "C" int dd_stem_patches(const float* x_in, const float* sigma, float sigma_data, const float* ln_freqs,
                               void* out, int B, int Cin, int H, int W, void* stream_) {
    return dd_stem_patches_cols(x_in, sigma, sigma_data, ln_freqs, out, B, Cin, H, W, 64, stream_);
}

extern "C" int dd_noise_embedding(const float* sigma, const float* freqs, const float* phases, int cnoise,
                                  const void* w_noise, int w_is_bf16, int normalize, const float* label_emb,
                                  float label_balance, float* emb_out, int B, int cemb, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(sigma && freqs && phases && w_noise && label_emb && emb_out, "dd_noise_embedding: null pointer");
    const dim3 grid(ceil_div(cemb, 8), B);
    const size_t smem = (size_t)cnoise * sizeof(float);
    if (w_is_bf16)
        DD_CHECK_CUDA(dd_launch_pdl(noise_embedding_kernel<true>, dim3(grid), dim3(256), smem, stream, sigma, freqs, phases, cnoise, w_noise, normalize,
                                                                  label_emb, label_balance, emb_out, cemb));
    else
        DD_CHECK_CUDA(dd_launch_pdl(noise_embedding_kernel<false>, dim3(grid), dim3(256), smem, stream, sigma, freqs, phases, cnoise, w_noise, normalize,
                                                                   label_emb, label_balance, emb_out, cemb));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_emb_affine(const dd_affine_desc* descs_dev, int n_descs, int max_O, const float* emb, int B, int cemb,
                             void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && emb && n_descs > 0 && max_O > 0, "dd_emb_affine: bad arguments");
    const dim3 grid(ceil_div(max_O, 8), n_descs);
    DD_CHECK_CUDA(dd_launch_pdl(emb_affine_kernel, dim3(grid), dim3(256), 0, stream, descs_dev, emb, B, cemb));
    DD_CHECK_LAUNCH();
    return 0;
}


extern "C" int dd_label_embedding(const float* emb_in, int Bc, int I, const void* w_label, const void* w_uncond,
                                  int w_is_bf16, const float* mask, int Bm, int normalize, float* out, int cemb,
                                  void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(emb_in && w_label && w_uncond && mask && out, "dd_label_embedding: null pointer");
    DD_REQUIRE(Bc == 1 || Bc == Bm, "dd_label_embedding: embedding batch %d does not broadcast to mask batch %d", Bc, Bm);
    const dim3 grid(ceil_div(cemb, 8), Bm);
    if (w_is_bf16)
        DD_CHECK_CUDA(dd_launch_pdl(label_embedding_kernel<true>, dim3(grid), dim3(256), 0, stream, emb_in, Bc, I, w_label, w_uncond, mask, normalize, out, cemb));
    else
        DD_CHECK_CUDA(dd_launch_pdl(label_embedding_kernel<false>, dim3(grid), dim3(256), 0, stream, emb_in, Bc, I, w_label, w_uncond, mask, normalize, out, cemb));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_sigma_logvar(const float* sigma, int count, const float* freqs, const float* phases, int n,
                               const void* w, int w_is_bf16, float* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(sigma && freqs && phases && w && out, "dd_sigma_logvar: null pointer");
    if (count == 0) return 0;
    if (w_is_bf16) logvar_kernel<true><<<ceil_div(count, 8), 256, 0, stream>>>(sigma, freqs, phases, n, w, out, count);
    else logvar_kernel<false><<<ceil_div(count, 8), 256, 0, stream>>>(sigma, freqs, phases, n, w, out, count);
    DD_CHECK_LAUNCH();
    return 0;
}


extern "C" int dd_mp_fourier(const float* x, int count, const float* freqs, const float* phases, int n, float* out,
                             void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && freqs && phases && out, "dd_mp_fourier: null pointer");
    const long total = (long)count * n;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(mp_fourier_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, x, freqs, phases, n, out, total));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_axpby(const void* a, const void* b, float alpha, float beta, float clip, void* out, long n,
                        void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(a && b && out, "dd_axpby: null pointer");
    DD_REQUIRE(n % 8 == 0, "dd_axpby: element count must be a multiple of 8");
    if (n == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(axpby_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, stream, static_cast<const uint4*>(a), static_cast<const uint4*>(b),
                                                           alpha, beta, clip > 0.f ? clip : INFINITY,
                                                           static_cast<uint4*>(out), n / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_sampler_cfg_lerp(const float* d_2b, const float* sample, float cfg_scale, float t_hat, float* cfg_out,
                                   float* x_hat_out, int dup, long n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(d_2b && sample && cfg_out, "dd_sampler_cfg_lerp: null pointer");
    DD_REQUIRE(n % 4 == 0, "dd_sampler_cfg_lerp: element count must be a multiple of 4");
    if (n == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(sampler_cfg_lerp_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, stream, 
        reinterpret_cast<const float4*>(d_2b), reinterpret_cast<const float4*>(sample), cfg_scale, t_hat,
        reinterpret_cast<float4*>(cfg_out), reinterpret_cast<float4*>(x_hat_out), dup, n / 4));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_sampler_update(const float* cfg1, const float* d2_2b, float cfg_scale, int use_heun, float t, float p,
                                 const float* noise, float* sample, float* cfg_out, int dup, long n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(cfg1 && sample && (!use_heun || d2_2b), "dd_sampler_update: null pointer");
    DD_REQUIRE(n % 4 == 0, "dd_sampler_update: element count must be a multiple of 4");
    if (n == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(sampler_update_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, stream, 
        reinterpret_cast<const float4*>(cfg1), reinterpret_cast<const float4*>(d2_2b), cfg_scale, use_heun, t, p,
        reinterpret_cast<const float4*>(noise), reinterpret_cast<float4*>(sample), reinterpret_cast<float4*>(cfg_out),
        dup, n / 4));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_mpconv_forward_naive(const void* x, const void* w, void* out, int B, int H, int W, int Cin, int Cout,
                                       int ksize, int groups, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && w && out, "dd_mpconv_forward_naive: null pointer");
    const long total = (long)B * H * W * Cout;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(conv_naive_kernel, dim3(grid_for(total, 256, 32)), dim3(256), 0, stream, 
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(out), B,
        H, W, Cin, Cout, ksize, groups));
    DD_CHECK_LAUNCH();
    return 0;
}
