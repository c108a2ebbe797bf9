// Glue kernels of the DAE_D3 diffusion-autoencoder decoder (reference: /root/reference/src/modules/daes/dae_edm2_d3.py,
// MPConv3D :43-93, Block.forward :186-238, DAE_D3.decode :356-369).
//
// Layout.  The reference keeps stereo as a depth axis Z = 2 of channels_last_3d tensors (B, C, 2, H, W) and convolves
// with (kz, 3, 3) kernels: reflection padding along W, one-sided reflection along Z, zero padding along H.  Here Z is
// folded into the channel dimension -- activations are NHWC bf16 [B][H][Wp][2C], channel = z*C + c -- which turns
//   * a (2,3,3) MPConv3D into a dense 3x3 convolution 2Cin -> 2Cout whose weight is block-circulant in z
//     (out[z'] = sum_z W[kz = z xor z'] * x[z]: exactly the reference's [x0, x1, x0] one-sided reflection), and
//   * a (1,k,k) MPConv3D into a 2-group convolution with the weight duplicated,
// so every MPConv3D runs on the tcgen05 implicit-GEMM kernels of conv_igemm.cu with identical FLOPs.  Reflection
// along W is physical: tensors carry `pw` halo columns on each side (Wp = W + 2*pw) that hold the mirrored columns; the
// convolution computes all Wp columns with TMA zero fill beyond them and dd_reflect_fill_w re-mirrors the halo columns
// of its output.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>

namespace {

inline int grid_for_d(long total, int block, int cap_mult = 16) {
    const long blocks = (total + block - 1) / block;
    return (int)std::max<long>(1, std::min<long>(blocks, (long)dd_num_sms() * cap_mult));
}

__device__ __forceinline__ void unpack8d(const uint4& q, float (&f)[8]) {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16x2(u[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8d(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// reflect a logical column index into [0, W)
__device__ __forceinline__ int reflect_idx(int w, int W) {
    if (w < 0) w = -w;
    if (w >= W) w = 2 * (W - 1) - w;
    return w;
}

// ------------------------------------------------------------------------------------------
// MPConv3D weight prep with Z folded into channels (eval mode: scale = gain / sqrt(fan_in), dae_edm2_d3.py:80-81)
//   kz == 2: out[(z', o)][tap][z*I + i] = scale * w[o][i][z xor z'][tap]      (dense over 2I input channels)
//   kz == 1: out[(z', o)][tap][i]       = scale * w[o][i][0][tap]             (2 groups, duplicated)
// ------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void weight_prep_z2_kernel(const void* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int I, int kz,
                                      int taps, const float* __restrict__ gain_dev, float gain_host, int i_stride) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int r = blockIdx.x;
    const int zp = r / O, o = r - zp * O;
    const float scale = gain_host * (gain_dev ? *gain_dev : 1.f) * rsqrtf((float)(I * kz * taps));
    const int row_len = taps * i_stride;
    __nv_bfloat16* dst = out + (size_t)r * row_len;
    const int n_in = kz == 2 ? 2 * I : I;
    for (int j = threadIdx.x; j < row_len; j += blockDim.x) {
        const int tap = j / i_stride, c = j - tap * i_stride;
        float v = 0.f;
        if (c < n_in) {
            const int z = c / I, i = c - z * I;
            const int kzi = kz == 2 ? (z ^ zp) : 0;
            const size_t src = (((size_t)o * I + i) * kz + kzi) * taps + tap;
            v = (kBf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(w)[src]) : static_cast<const float*>(w)[src]) * scale;
        }
        dst[j] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------
// halo columns <- mirrored logical columns:  x[.., pw-k] = x[.., pw+k],  x[.., pw+W-1+k] = x[.., pw+W-1-k]
// ------------------------------------------------------------------------------------------
__global__ void reflect_fill_w_kernel(uint4* __restrict__ x, long rows, int Wp, int nvec, int pw) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const long total = rows * 2 * pw * nvec;
    const int W = Wp - 2 * pw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int v = (int)(idx % nvec);
        long r = idx / nvec;
        const int k = (int)(r % (2 * pw));
        r /= (2 * pw);
        int dst_col, src_col;
        if (k < pw) { dst_col = pw - 1 - k; src_col = pw + 1 + k; }
        else        { const int kk = k - pw; dst_col = pw + W + kk; src_col = pw + W - 2 - kk; }
        uint4* row = x + r * (long)Wp * nvec;
        row[(long)dst_col * nvec + v] = row[(long)src_col * nvec + v];
    }
}

// ------------------------------------------------------------------------------------------
// decoder stem input (DAE_D3.decode :358-359): latents (B, 2L, H, W) fp32 NCHW, channel c8 = c*2 + z
//   -> [B][H][Wp][Cpad] bf16 with channel z*(L+1) + c, the constant-one channel at c == L, zeros above 2(L+1)
// ------------------------------------------------------------------------------------------
__global__ void dae_stem_kernel(const float* __restrict__ lat, __nv_bfloat16* __restrict__ out, int B, int L, int H, int W,
                                int pw, int Cpad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Wp = W + 2 * pw;
    const long total = (long)B * H * Wp * Cpad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(idx % Cpad);
        long pix = idx / Cpad;
        const int p = (int)(pix % Wp);
        pix /= Wp;
        const int h = (int)(pix % H), b = (int)(pix / H);
        const int w = reflect_idx(p - pw, W);
        float v = 0.f;
        if (ch < 2 * (L + 1)) {
            const int z = ch / (L + 1), c = ch - z * (L + 1);
            v = c == L ? 1.f : lat[(((size_t)b * 2 * L + c * 2 + z) * H + h) * W + w];
        }
        out[idx] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------
// resample_3d "up" (mp_tools.py:92-93: nearest x2 in H and W) on a W-padded tensor + mp_silu
// ------------------------------------------------------------------------------------------
// I = index type (32-bit whenever the output has fewer than 2^31 vectors: 64-bit div / mod by run-time values costs more
// instructions than the rest of the body)
template <typename I>
__global__ void up2_silu_pad_kernel(const uint4* __restrict__ a, uint4* __restrict__ xc, uint4* __restrict__ s, int B, int Ha,
                                    int Wa, int pw, int nvec) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int H = 2 * Ha, W = 2 * Wa, Wp = W + 2 * pw, Wpa = Wa + 2 * pw;
    const I total = (I)B * (I)H * (I)Wp * (I)nvec;
    for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
        I pix = idx / (I)nvec;
        const int v = (int)(idx - pix * (I)nvec);
        const I r = pix / (I)Wp;
        const int p = (int)(pix - r * (I)Wp);
        const I b = r / (I)H;
        const int h = (int)(r - b * (I)H);
        const int w = reflect_idx(p - pw, W);
        const uint4 q = __ldg(a + (((size_t)b * Ha + (h >> 1)) * Wpa + (w >> 1) + pw) * nvec + v);
        float f[8], o[8];
        unpack8d(q, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = mp_silu_fast(f[j]);
        xc[idx] = q;
        s[idx] = pack8d(o);
    }
}

// ------------------------------------------------------------------------------------------
// conv_out: MPConv3D (1,5,5), C channels per stereo side -> 1 (DAE_D3.decode :368), direct convolution on CUDA cores
// (9 GFLOP per 45 s item; the kernel is bound by reading the level-0 activation once).
//   x [B][H][Wp][2C] bf16 (pw >= 2), w fp32 [C][25] pre-scaled by 1/sqrt(25 C) (dd_weight_prep, DD_WFMT_F32_OIT),
//   out fp32 [B][2][H][W]
// ------------------------------------------------------------------------------------------
// Tile of 32 x 8 output pixels per CTA: the (32+4) x (8+4) input pixels x 2C channels are staged once in shared
// memory (16-byte channel vectors XOR-swizzled by the pixel column so that the 32 lanes of a warp, which read the same
// channel vector of 32 neighbouring pixels, hit 32 different banks), then every thread accumulates its pixel's 25 taps.
constexpr int kC5W = 32, kC5H = 8;

template <int C>
__global__ void __launch_bounds__(kC5W * kC5H) conv5x5_out_kernel(const uint4* __restrict__ x, const float* __restrict__ wq,
                                                                 const float* __restrict__ gain, float* __restrict__ out,
                                                                 int B, int H, int W, int pw) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    constexpr int NV = C / 8;                    // uint4 vectors per stereo side
    constexpr int PV = 2 * NV;                   // vectors per pixel
    constexpr int TW = kC5W + 4, TH = kC5H + 4;
    extern __shared__ __align__(16) uint8_t smem_c5[];
    uint4* tile = reinterpret_cast<uint4*>(smem_c5);                       // [TH][TW][PV], vector index ^ (col & 7)
    float* ws = reinterpret_cast<float*>(tile + TH * TW * PV);             // [C][25]
    const float g = gain ? *gain : 1.f;
    const int tid = threadIdx.y * kC5W + threadIdx.x;
    for (int i = tid; i < 25 * C; i += kC5W * kC5H) ws[i] = wq[i] * g;
    const int Wp = W + 2 * pw;
    const int b = blockIdx.z, h0 = blockIdx.y * kC5H, w0 = blockIdx.x * kC5W;
    for (int i = tid; i < TH * TW * PV; i += kC5W * kC5H) {
        const int v = i % PV, col = (i / PV) % TW, row = i / (PV * TW);
        const int hh = h0 + row - 2, pc = w0 + col + pw - 2;               // physical column (halo columns hold the reflection)
        uint4 q = make_uint4(0, 0, 0, 0);
        if (hh >= 0 && hh < H && pc < Wp) q = __ldg(x + (((long)b * H + hh) * Wp + pc) * PV + v);
        tile[(row * TW + col) * PV + (v ^ (col & 7))] = q;
    }
    __syncthreads();
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < 5; ++dy) {
#pragma unroll 1
        for (int dx = 0; dx < 5; ++dx) {
            const int col = threadIdx.x + dx;
            const uint4* px = tile + ((threadIdx.y + dy) * TW + col) * PV;
            const float* wt = ws + dy * 5 + dx;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float f0[8], f1[8];
                unpack8d(px[v ^ (col & 7)], f0);
                unpack8d(px[(NV + v) ^ (col & 7)], f1);
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc0 += f0[j] * wt[(v * 8 + j) * 25]; acc1 += f1[j] * wt[(v * 8 + j) * 25]; }
            }
        }
    }
    if (w < W && h < H) {
        out[(((size_t)b * 2 + 0) * H + h) * W + w] = acc0;
        out[(((size_t)b * 2 + 1) * H + h) * W + w] = acc1;
    }
}

// ------------------------------------------------------------------------------------------
// ddec UNet input (unet_edm2_ddec_mclt_b1.py:294-309): per stereo side z the channels are
//   [ c_in(sigma) * x_in[b][z] ,  x_ref[b][z][h*k + j] for j < k  (the view/permute(0,3,1,2,4) of :294-295) ,  1 ]
// -> folded [B][F][Wp][Cpad] bf16, channel z*(k+2) + c, zeros above 2(k+2), halo columns mirrored
// ------------------------------------------------------------------------------------------
__global__ void ddec_stem_kernel(const float* __restrict__ x_in, const float* __restrict__ x_ref,
                                 const float* __restrict__ sigma, float sigma_data, __nv_bfloat16* __restrict__ out, int B,
                                 int Fq, int W, int k, int pw, int Cpad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Wp = W + 2 * pw, ct = k + 2;
    const long total = (long)B * Fq * Wp * Cpad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(idx % Cpad);
        long pix = idx / Cpad;
        const int p = (int)(pix % Wp);
        pix /= Wp;
        const int h = (int)(pix % Fq), b = (int)(pix / Fq);
        const int w = reflect_idx(p - pw, W);
        float v = 0.f;
        if (ch < 2 * ct) {
            const int z = ch / ct, c = ch - z * ct;
            if (c == 0) {
                const float sg = sigma[b];
                v = x_in[(((size_t)b * 2 + z) * Fq + h) * W + w] * rsqrtf(sigma_data * sigma_data + sg * sg);
            } else if (c <= k) {
                v = x_ref[(((size_t)b * 2 + z) * Fq * k + (size_t)h * k + (c - 1)) * W + w];
            } else {
                v = 1.f;
            }
        }
        out[idx] = __float2bfloat16_rn(v);
    }
}

// resample_3d "down" (mp_tools.py:85-90: 2x2 mean over H, W) on a W-padded tensor, halo columns of the result mirrored
__global__ void avgpool2_pad_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int pw,
                                    int nvec) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Ho = H >> 1, Wo = W >> 1, Wp = W + 2 * pw, Wpo = Wo + 2 * pw;
    const long total = (long)B * Ho * Wpo * nvec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int v = (int)(idx % nvec);
        long pix = idx / nvec;
        const int p = (int)(pix % Wpo);
        pix /= Wpo;
        const int h = (int)(pix % Ho), b = (int)(pix / Ho);
        const int w = reflect_idx(p - pw, Wo);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float f[8];
                unpack8d(__ldg(x + (((long)b * H + 2 * h + dy) * Wp + 2 * w + dx + pw) * nvec + v), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
        out[idx] = pack8d(acc);
    }
}

// ddec output head (:323-326): D = c_skip*x_in + c_out*F, F = conv_out result (folded: channel z of a Cst-wide pixel)
__global__ void ddec_head_kernel(const __nv_bfloat16* __restrict__ f, const float* __restrict__ x_in,
                                 const float* __restrict__ sigma, float sigma_data, float* __restrict__ out, int B, int H,
                                 int W, int pw, int Cst) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Wp = W + 2 * pw;
    const long total = (long)B * 2 * H * W;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int w = (int)(idx % W);
        long r = idx / W;
        const int h = (int)(r % H);
        r /= H;
        const int z = (int)(r % 2), b = (int)(r / 2);
        const float sg = sigma[b], sd2 = sigma_data * sigma_data;
        const float c_skip = sd2 / (sg * sg + sd2), c_out = sg * sigma_data * rsqrtf(sg * sg + sd2);
        const float fv = __bfloat162float(f[(((size_t)b * H + h) * Wp + w + pw) * Cst + z]);
        out[idx] = c_skip * x_in[idx] + c_out * fv;
    }
}

// ------------------------------------------------------------------------------------------
// unet_edm2_q4_ddec input (unet_edm2_q4_ddec.py:268-277): mp_cat(c_in*x_in, x_ref') with
//   x_ref' = x_ref.view(B, C, F, k, W).permute(0, 3, 1, 2, 4).reshape(B, k*C, F, W)   (channel j*C + c)
// -> NHWC bf16 [B][F][W][Cpad]: channels [wa*c_in*x (C) | wb*x_ref' (k*C) | 1 (feeds the conv_in bias column) | 0...]
// ------------------------------------------------------------------------------------------
__global__ void q4_stem_kernel(const float* __restrict__ x_in, const float* __restrict__ x_ref, const float* __restrict__ sigma,
                               float sigma_data, float wa, float wb, __nv_bfloat16* __restrict__ out, int B, int C, int Fq,
                               int W, int k, int Cpad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const long total = (long)B * Fq * W * Cpad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(idx % Cpad);
        long pix = idx / Cpad;
        const int w = (int)(pix % W);
        pix /= W;
        const int h = (int)(pix % Fq), b = (int)(pix / Fq);
        float v = 0.f;
        if (ch < C) {
            const float sg = sigma[b];
            // the reference rounds c_in*x to bf16 before the concat scale is applied (:274-277)
            const float xs = __bfloat162float(__float2bfloat16_rn(x_in[(((size_t)b * C + ch) * Fq + h) * W + w] *
                                                                  rsqrtf(sigma_data * sigma_data + sg * sg)));
            v = wa * xs;
        } else if (ch < C + k * C) {
            const int j = (ch - C) / C, c = (ch - C) - j * C;
            const float xr = __bfloat162float(__float2bfloat16_rn(x_ref[(((size_t)b * C + c) * Fq * k + (size_t)h * k + j) * W + w]));
            v = wb * xr;
        } else if (ch == C + k * C) {
            v = 1.f;
        }
        out[idx] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------
// DAE_D3.encode input (dae_edm2_d3.py:344-346 + conv_in (1,5,5), :283): per stereo side the 5x5 patch of [mel, 1]
// (reflection along W, zeros along H) as 50 values padded to 64 -> [B][H][Wp][128] bf16 (channel z*64 + tap*2 + c);
// conv_in then runs as a 2-group K = 64 tensor-core GEMM.
// ------------------------------------------------------------------------------------------
__global__ void dae_enc_patches_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                       int pw) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Wp = W + 2 * pw;
    const long total = (long)B * H * Wp * 128;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(idx & 127);
        long pix = idx >> 7;
        const int p = (int)(pix % Wp);
        pix /= Wp;
        const int h = (int)(pix % H), b = (int)(pix / H);
        const int z = ch >> 6, k = ch & 63;
        float v = 0.f;
        if (k < 50) {
            const int tap = k >> 1, c = k & 1;
            const int hh = h + tap / 5 - 2;
            if (hh >= 0 && hh < H) {
                const int w = reflect_idx(p - pw, W);
                const int ww = reflect_idx(w + tap % 5 - 2, W);
                v = c == 0 ? mel[(((size_t)b * 2 + z) * H + hh) * W + ww] : 1.f;
            }
        }
        out[idx] = __float2bfloat16_rn(v);
    }
}

// DAE_D3.encode tail (:352-354): conv_latents_out result [B][H][Wp][Cst] bf16 (channel z*L + c) -> tensor_5d_to_4d
// (channel c*2 + z) -> avg_pool2d(ratio) -> fp32 NCHW (B, 2L, H/ratio, W/ratio)
__global__ void dae_latents_pool_kernel(const __nv_bfloat16* __restrict__ f, float* __restrict__ out, int B, int L, int H,
                                        int W, int pw, int Cst, int ratio) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Ho = H / ratio, Wo = W / ratio, Wp = W + 2 * pw;
    const long total = (long)B * 2 * L * Ho * Wo;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int wo = (int)(idx % Wo);
        long r = idx / Wo;
        const int ho = (int)(r % Ho);
        r /= Ho;
        const int c8 = (int)(r % (2 * L)), b = (int)(r / (2 * L));
        const int c = c8 >> 1, z = c8 & 1;
        float acc = 0.f;
        for (int dy = 0; dy < ratio; ++dy)
            for (int dx = 0; dx < ratio; ++dx)
                acc += __bfloat162float(f[(((size_t)b * H + ho * ratio + dy) * Wp + wo * ratio + dx + pw) * Cst + z * L + c]);
        out[idx] = acc / (float)(ratio * ratio);
    }
}

// ------------------------------------------------------------------------------------------
// dae_edm2_q4.DAE (modules/daes/dae_edm2_q4.py): the plain 2-D sibling of DAE_D3 -- stereo is a channel pair, every
// convolution a zero-padded MPConv (mp_tools.py:357-373).  Only the ends of the network need kernels of their own.
// ------------------------------------------------------------------------------------------
// fp32 NCHW (B, C, H, W) -> bf16 NHWC [B][H][W][Cpad]; channel `ones_channel` (>= C, or < 0 for none) is the constant 1
// that carries a convolution bias (centre-tap weight column), the other padding channels are zero.
__global__ void pack_nhwc_kernel(const float* __restrict__ x, uint4* __restrict__ out, int B, int C, int H, int W, int nvec,
                                 int ones_channel) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const unsigned total = (unsigned)B * H * W * nvec;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned pix = idx / (unsigned)nvec;
        const int v = (int)(idx - pix * (unsigned)nvec);
        const unsigned b = pix / (unsigned)(H * W);
        const unsigned hw = pix - b * (unsigned)(H * W);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = v * 8 + j;
            f[j] = c < C ? __ldg(x + ((size_t)b * C + c) * H * W + hw) : (c == ones_channel ? 1.f : 0.f);
        }
        out[idx] = pack8d(f);
    }
}

// bf16 NHWC [B][H][W][Cpad] -> fp32 NCHW (B, C, H, W), the first C channels.
__global__ void unpack_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W,
                                   int Cpad) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const unsigned total = (unsigned)B * C * H * W;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned hw = idx % (unsigned)(H * W);
        const unsigned bc = idx / (unsigned)(H * W);
        const unsigned b = bc / (unsigned)C, c = bc - b * (unsigned)C;
        out[idx] = __bfloat162float(x[((size_t)b * H * W + hw) * Cpad + c]);
    }
}

// conv_in (5,5) of a C-channel fp32 image (C = 2) as a K = cols GEMM: patch[pix][tap*C + c], tap = ky*5 + kx, zeros outside
// the image, column 25*C the constant 1 (the bias column), zero above.
__global__ void patches5x5_kernel(const float* __restrict__ x, uint4* __restrict__ out, int B, int C, int H, int W, int vecs) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const unsigned total = (unsigned)B * H * W * vecs;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned pix = idx / (unsigned)vecs;
        const int v = (int)(idx - pix * (unsigned)vecs);
        const unsigned row = pix / (unsigned)W;
        const int w = (int)(pix - row * (unsigned)W);
        const int b = (int)(row / (unsigned)H);
        const int h = (int)(row - (unsigned)b * (unsigned)H);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = v * 8 + j;
            float val = 0.f;
            if (k < 25 * C) {
                const int tap = k / C, c = k - tap * C;
                const int hh = h + tap / 5 - 2, ww = w + tap % 5 - 2;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(x + (((size_t)b * C + c) * H + hh) * W + ww);
            } else if (k == 25 * C) {
                val = 1.f;
            }
            f[j] = val;
        }
        out[idx] = pack8d(f);
    }
}

// conv_out (5,5), C -> Cout <= 4 dense, zero padding, times *gain_dev: direct convolution on CUDA cores (the layer is
// bound by reading the level-0 activation once; 25 C Cout MACs per pixel).  A warp owns 32 consecutive pixels of a row:
// lanes split the channels (two per lane and trip), every pixel's sum is reduced across the warp and kept by the lane of
// that pixel, so that the fp32 NCHW result is written with coalesced stores.
//   x [B][H][W][C] bf16, w fp32 [Cout][25][C] pre-scaled by 1/sqrt(25 C), out fp32 (B, Cout, H, W)
constexpr int kC5dWarps = 8;
constexpr int kC5dMaxOut = 4;
template <int Cout>
__global__ void __launch_bounds__(kC5dWarps * 32)
conv5x5_dense_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wq, const float* __restrict__ gain_dev,
                     float* __restrict__ out, int B, int H, int W, int C) {
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    extern __shared__ float ws[];      // [Cout][25][C]
    for (int i = threadIdx.x; i < Cout * 25 * C; i += blockDim.x) ws[i] = wq[i];
    __syncthreads();
    const float gain = gain_dev ? *gain_dev : 1.f;
    const int lane = threadIdx.x & 31;
    const int strips_w = (W + 31) / 32;
    const unsigned total = (unsigned)B * H * strips_w;
    for (unsigned sidx = blockIdx.x * kC5dWarps + (threadIdx.x >> 5); sidx < total; sidx += gridDim.x * kC5dWarps) {
        const unsigned row = sidx / (unsigned)strips_w;
        const int w0 = (int)(sidx - row * (unsigned)strips_w) * 32;
        const int b = (int)(row / (unsigned)H);
        const int h = (int)(row - (unsigned)b * (unsigned)H);
        float res[Cout];
#pragma unroll
        for (int o = 0; o < Cout; ++o) res[o] = 0.f;
        const int npix = min(32, W - w0);
        for (int p = 0; p < npix; ++p) {
            float acc[Cout];
#pragma unroll
            for (int o = 0; o < Cout; ++o) acc[o] = 0.f;
            for (int ky = 0; ky < 5; ++ky) {
                const int hh = h + ky - 2;
                if (hh < 0 || hh >= H) continue;
                for (int kx = 0; kx < 5; ++kx) {
                    const int ww = w0 + p + kx - 2;
                    if (ww < 0 || ww >= W) continue;
                    const __nv_bfloat16* xp = x + (((size_t)b * H + hh) * W + ww) * C;
                    const float* wt = ws + (ky * 5 + kx) * C;
                    for (int c = lane * 2; c < C; c += 64) {
                        const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xp + c)));
#pragma unroll
                        for (int o = 0; o < Cout; ++o) {
                            const float2 wv = *reinterpret_cast<const float2*>(wt + (size_t)o * 25 * C + c);
                            acc[o] = fmaf(f.x, wv.x, fmaf(f.y, wv.y, acc[o]));
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < Cout; ++o) {
                const float t = warp_sum(acc[o]);
                if (lane == p) res[o] = t;
            }
        }
        if (lane < npix) {
#pragma unroll
            for (int o = 0; o < Cout; ++o) out[(((size_t)b * Cout + o) * H + h) * W + w0 + lane] = res[o] * gain;
        }
    }
}

}  // namespace

extern "C" int dd_weight_prep_z2(const void* w, int w_is_bf16, void* out, int O, int I, int kz, int taps,
                                 const float* gain_dev, float gain_host, int i_stride, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(w && out && O > 0 && I > 0 && taps > 0, "dd_weight_prep_z2: bad arguments");
    DD_REQUIRE(kz == 1 || kz == 2, "dd_weight_prep_z2: kz=%d unsupported (stereo depth 2)", kz);
    const int n_in = kz == 2 ? 2 * I : I;
    if (i_stride <= 0) i_stride = n_in;
    DD_REQUIRE(i_stride >= n_in, "dd_weight_prep_z2: i_stride smaller than the input channel count");
    if (w_is_bf16)
        DD_CHECK_CUDA(dd_launch_pdl(weight_prep_z2_kernel<true>, dim3(2 * O), dim3(128), 0, stream, w, static_cast<__nv_bfloat16*>(out), O, I, kz, taps, gain_dev,
                                                               gain_host, i_stride));
    else
        DD_CHECK_CUDA(dd_launch_pdl(weight_prep_z2_kernel<false>, dim3(2 * O), dim3(128), 0, stream, w, static_cast<__nv_bfloat16*>(out), O, I, kz, taps, gain_dev,
                                                                gain_host, i_stride));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_reflect_fill_w(void* x, int B, int H, int Wp, int C, int pw, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && C % 8 == 0 && pw >= 1 && Wp - 2 * pw >= pw + 1, "dd_reflect_fill_w: bad arguments");
    const long rows = (long)B * H;
    const long total = rows * 2 * pw * (C / 8);
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(reflect_fill_w_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, static_cast<uint4*>(x), rows, Wp, C / 8, pw));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_dae_stem(const float* latents, void* out, int B, int L, int H, int W, int pw, int Cpad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(latents && out && L > 0 && 2 * (L + 1) <= Cpad && pw >= 1 && W > pw, "dd_dae_stem: bad arguments");
    const long total = (long)B * H * (W + 2 * pw) * Cpad;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(dae_stem_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, latents, static_cast<__nv_bfloat16*>(out), B, L, H, W, pw,
                                                                Cpad));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_up2_silu_pad(const void* a, void* xc, void* s, int B, int Ha, int Wa, int C, int pw, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(a && xc && s && C % 8 == 0 && pw >= 1 && Wa >= 1, "dd_up2_silu_pad: bad arguments");
    const long total = (long)B * 2 * Ha * (2 * Wa + 2 * pw) * (C / 8);
    if (total == 0) return 0;
    const int grid = grid_for_d(total, 256);
    const bool narrow = total < (1L << 31) - (long)grid * 256;      // idx + stride must not wrap
    DD_CHECK_CUDA(dd_launch_pdl(narrow ? up2_silu_pad_kernel<unsigned> : up2_silu_pad_kernel<long>, dim3(grid), dim3(256), 0, stream,
                                static_cast<const uint4*>(a), static_cast<uint4*>(xc), static_cast<uint4*>(s), B, Ha, Wa, pw, C / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_conv5x5_out(const void* x, const float* w25, const float* gain_dev, float* out, int B, int H, int W, int C,
                              int pw, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && w25 && out && pw >= 2, "dd_conv5x5_out: bad arguments (needs two halo columns)");
    if ((long)B * H * W == 0) return 0;
    DD_REQUIRE(C == 32 || C == 64, "dd_conv5x5_out: C=%d unsupported (32 or 64 channels per stereo side)", C);
    DD_REQUIRE(B <= 65535 && ceil_div(H, kC5H) <= 65535, "dd_conv5x5_out: grid too large");
    const dim3 grid(ceil_div(W, kC5W), ceil_div(H, kC5H), B), block(kC5W, kC5H);
    const size_t smem = (size_t)(kC5W + 4) * (kC5H + 4) * (2 * C / 8) * 16 + (size_t)25 * C * sizeof(float);
    if (C == 32) {
        static bool done = false;
        if (!done) { DD_CHECK_CUDA(cudaFuncSetAttribute(conv5x5_out_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
        DD_CHECK_CUDA(dd_launch_pdl(conv5x5_out_kernel<32>, dim3(grid), dim3(block), smem, stream, static_cast<const uint4*>(x), w25, gain_dev, out, B, H, W, pw));
    } else {
        static bool done = false;
        if (!done) { DD_CHECK_CUDA(cudaFuncSetAttribute(conv5x5_out_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
        DD_CHECK_CUDA(dd_launch_pdl(conv5x5_out_kernel<64>, dim3(grid), dim3(block), smem, stream, static_cast<const uint4*>(x), w25, gain_dev, out, B, H, W, pw));
    }
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_ddec_stem(const float* x_in, const float* x_ref, const float* sigma, float sigma_data, void* out, int B,
                            int F, int W, int k, int pw, int Cpad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x_in && x_ref && sigma && out && k >= 1 && 2 * (k + 2) <= Cpad && pw >= 1 && W > pw, "dd_ddec_stem: bad arguments");
    const long total = (long)B * F * (W + 2 * pw) * Cpad;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(ddec_stem_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, x_in, x_ref, sigma, sigma_data,
                                                                 static_cast<__nv_bfloat16*>(out), B, F, W, k, pw, Cpad));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_avgpool2_pad(const void* x, void* out, int B, int H, int W, int C, int pw, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && pw >= 1 && W / 2 > pw, "dd_avgpool2_pad: bad arguments");
    const long total = (long)B * (H / 2) * (W / 2 + 2 * pw) * (C / 8);
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(avgpool2_pad_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, static_cast<const uint4*>(x), static_cast<uint4*>(out), B,
                                                                    H, W, pw, C / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_ddec_head(const void* f, const float* x_in, const float* sigma, float sigma_data, float* out, int B, int H,
                            int W, int pw, int Cst, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(f && x_in && sigma && out && Cst >= 2, "dd_ddec_head: bad arguments");
    const long total = (long)B * 2 * H * W;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(ddec_head_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(f), x_in, sigma, sigma_data,
                                                                 out, B, H, W, pw, Cst));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_q4_stem(const float* x_in, const float* x_ref, const float* sigma, float sigma_data, float wa, float wb,
                          void* out, int B, int C, int F, int W, int k, int Cpad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x_in && x_ref && sigma && out && k >= 1 && C * (k + 1) + 1 <= Cpad, "dd_q4_stem: bad arguments");
    const long total = (long)B * F * W * Cpad;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(q4_stem_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, x_in, x_ref, sigma, sigma_data, wa, wb,
                                                               static_cast<__nv_bfloat16*>(out), B, C, F, W, k, Cpad));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_dae_enc_patches(const float* mel, void* out, int B, int H, int W, int pw, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(mel && out && pw >= 1 && W > 2 && H > 0, "dd_dae_enc_patches: bad arguments");
    const long total = (long)B * H * (W + 2 * pw) * 128;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(dae_enc_patches_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, mel, static_cast<__nv_bfloat16*>(out), B, H, W, pw));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_dae_latents_pool(const void* f, float* out, int B, int L, int H, int W, int pw, int Cst, int ratio,
                                   void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(f && out && L > 0 && 2 * L <= Cst && ratio >= 1 && H % ratio == 0 && W % ratio == 0,
               "dd_dae_latents_pool: bad arguments");
    const long total = (long)B * 2 * L * (H / ratio) * (W / ratio);
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(dae_latents_pool_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(f), out, B, L, H, W,
                                                                        pw, Cst, ratio));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_pack_nhwc(const float* x, void* out, int B, int C, int H, int W, int Cpad, int ones_channel, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && out && C > 0 && Cpad >= C && Cpad % 8 == 0 && ones_channel < Cpad && (ones_channel < 0 || ones_channel >= C),
               "dd_pack_nhwc: bad arguments");
    const long total = (long)B * H * W * (Cpad / 8);
    if (total == 0) return 0;
    DD_REQUIRE(total < (1L << 31) - (1L << 24), "dd_pack_nhwc: tensor too large");
    DD_CHECK_CUDA(dd_launch_pdl(pack_nhwc_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, x, static_cast<uint4*>(out), B, C, H, W,
                                Cpad / 8, ones_channel));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_unpack_nchw(const void* x, float* out, int B, int C, int H, int W, int Cpad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && out && C > 0 && Cpad >= C, "dd_unpack_nchw: bad arguments");
    const long total = (long)B * C * H * W;
    if (total == 0) return 0;
    DD_REQUIRE(total < (1L << 31) - (1L << 24), "dd_unpack_nchw: tensor too large");
    DD_CHECK_CUDA(dd_launch_pdl(unpack_nchw_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(x), out, B, C,
                                H, W, Cpad));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_patches5x5(const float* x, void* out, int B, int C, int H, int W, int cols, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && out && C > 0, "dd_patches5x5: bad arguments");
    DD_REQUIRE(cols == 64 || cols == 128, "dd_patches5x5: patch width %d unsupported (64 or 128)", cols);
    DD_REQUIRE(25 * C + 1 <= cols, "dd_patches5x5: in_channels=%d unsupported (25*C + 1 must be <= %d)", C, cols);
    const long total = (long)B * H * W * (cols / 8);
    if (total == 0) return 0;
    DD_REQUIRE(total < (1L << 31) - (1L << 24), "dd_patches5x5: tensor too large");
    DD_CHECK_CUDA(dd_launch_pdl(patches5x5_kernel, dim3(grid_for_d(total, 256)), dim3(256), 0, stream, x, static_cast<uint4*>(out), B, C, H, W,
                                cols / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_conv5x5_dense(const void* x, const float* w, const float* gain_dev, float* out, int B, int H, int W, int C,
                                int Cout, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && w && out, "dd_conv5x5_dense: null pointer");
    DD_REQUIRE(C > 0 && C % 2 == 0 && Cout >= 1 && Cout <= kC5dMaxOut, "dd_conv5x5_dense: C=%d Cout=%d unsupported", C, Cout);
    const size_t smem = (size_t)Cout * 25 * C * sizeof(float);
    static_assert(kC5dMaxOut == 4, "dispatch below");
    DD_REQUIRE(smem <= 200 * 1024, "dd_conv5x5_dense: weights (%zu bytes) do not fit in shared memory", smem);
    const long strips = (long)B * H * ((W + 31) / 32);
    if (strips == 0) return 0;
    DD_REQUIRE(strips < (1L << 31) - (1L << 24), "dd_conv5x5_dense: tensor too large");
    auto kernel = Cout == 1 ? conv5x5_dense_kernel<1> : Cout == 2 ? conv5x5_dense_kernel<2> : Cout == 3 ? conv5x5_dense_kernel<3>
                                                                                                       : conv5x5_dense_kernel<4>;
    if (smem > 48 * 1024) DD_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const long want = (strips + kC5dWarps - 1) / kC5dWarps;
    const int grid = (int)(want < (long)dd_num_sms() * 8 ? want : (long)dd_num_sms() * 8);
    DD_CHECK_CUDA(dd_launch_pdl(kernel, dim3(grid), dim3(kC5dWarps * 32), smem, stream, static_cast<const __nv_bfloat16*>(x), w, gain_dev, out, B, H,
                                W, C));
    DD_CHECK_LAUNCH();
    return 0;
}
