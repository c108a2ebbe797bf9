// Backward-pass glue of the EDM2 UNet train step (reference: loss.backward() through
// modules/unets/unet_edm2_b4.py:110-158,250-296 and modules/mp_tools.py:42-49,268-301,357-373, driven by
// training/trainer.py:1022-1044).  The two GEMM-shaped halves of every MPConv backward run on the tensor
// cores (dgrad = dd_mpconv_forward on transposed/flipped weights, wgrad = dd_mpconv_wgrad); this file holds
// everything between them: the weight-normalisation backward, the derivatives of the fused elementwise block
// glue (pixel-norm, mp_silu, mp_sum, mp_cat, clip, resample), the attention backward and the embedding heads.
// Activations are NHWC bf16, all arithmetic is fp32.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>

namespace {

constexpr float kNormEps = 1e-4f;   // modules/mp_tools.py:43
constexpr float kInvSiluGain = 1.0f / 0.596f;

// d/dx mp_silu(x) = sigmoid(x) * (1 + x * (1 - sigmoid(x))) / 0.596
// sigmoid(x) = 0.5 + 0.5 tanh(x/2): one MUFU (tanh.approx, ~2^-11) instead of an exp and an IEEE division -- the result
// multiplies a bf16 gradient and is rounded to bf16, and the glue kernels that call this were issue-bound on the latter
__device__ __forceinline__ float mp_silu_grad(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    const float s = fmaf(0.5f, t, 0.5f);
    return s * fmaf(x, 1.0f - s, 1.0f) * kInvSiluGain;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16x2(u[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

__device__ __forceinline__ float block_sum_b(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

inline int grid_for_b(long total, int block, int cap_mult = 8) {
    const long blocks = (total + block - 1) / block;
    return (int)std::max<long>(1, std::min<long>(blocks, (long)dd_num_sms() * cap_mult));
}

// ------------------------------------------------------------------------------------------
// bf16 weight transpose for dgrad: Wp[g*cout_g + co][tap][ci] -> Wd[g*cin_g + ci][taps-1-tap][co]
// ------------------------------------------------------------------------------------------
__global__ void weight_transpose_kernel(const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ wd, int cout_g,
                                        int cin_g, int taps) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ __nv_bfloat16 tile[32][33];
    const int g = blockIdx.z / taps, tap = blockIdx.z - g * taps;
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int co = co0 + r, ci = ci0 + threadIdx.x;
        if (co < cout_g && ci < cin_g)
            tile[r][threadIdx.x] = wp[((size_t)(g * cout_g + co) * taps + tap) * cin_g + ci];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int ci = ci0 + r, co = co0 + threadIdx.x;
        if (co < cout_g && ci < cin_g)
            wd[((size_t)(g * cin_g + ci) * taps + (taps - 1 - tap)) * cout_g + co] = tile[threadIdx.x][r];
    }
}

__global__ void weight_transpose_batched_kernel(const dd_wtrans_desc* __restrict__ descs, int n_descs) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ __nv_bfloat16 tile[32][33];
    int lo = 0, hi = n_descs - 1;
    const int t = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
    }
    const dd_wtrans_desc d = descs[lo];
    int r = t - d.tile_begin;
    const int nx = (d.cin_g + 31) / 32, ny = (d.cout_g + 31) / 32;
    const int bx = r % nx; r /= nx;
    const int by = r % ny; r /= ny;
    const int g = r / d.taps, tap = r - g * d.taps;
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(d.src);
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(d.dst);
    const int ci0 = bx * 32, co0 = by * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int co = co0 + rr, ci = ci0 + threadIdx.x;
        if (co < d.cout_g && ci < d.cin_g)
            tile[rr][threadIdx.x] = wp[((size_t)(g * d.cout_g + co) * d.taps + tap) * d.cin_g + ci];
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int ci = ci0 + rr, co = co0 + threadIdx.x;
        if (co < d.cout_g && ci < d.cin_g)
            wd[((size_t)(g * d.cin_g + ci) * d.taps + (d.taps - 1 - tap)) * d.cout_g + co] = tile[threadIdx.x][rr];
    }
}

// normalize_weights() (mp_tools.py:375-378, run by the trainer after every optimizer step, trainer.py:1107-1108) for a
// whole parameter set in one launch, in place on the fp32 parameters: w <- w / (eps + ||w|| / sqrt(fan_in)) per row.
__global__ void __launch_bounds__(256) weight_normalize_batched_kernel(const dd_wprep_desc* __restrict__ descs, int n_descs,
                                                                       int total_rows) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    // one warp per weight row, two consecutive rows per descriptor search (as weight_prep_batched_kernel below)
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 2;
    if (row0 >= total_rows) return;
    int lo = 0, hi = n_descs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].row_begin <= row0) lo = mid; else hi = mid - 1;
    }
    dd_wprep_desc d = descs[lo];
    for (int r = 0; r < 2; ++r) {
        const int row = row0 + r;
        if (row >= total_rows) return;
        while (lo + 1 < n_descs && descs[lo + 1].row_begin <= row) d = descs[++lo];
        const int o = row - d.row_begin;
        if (o >= d.O) continue;
        const int f = d.I_g * d.taps;
        float* w = static_cast<float*>(const_cast<void*>(d.w)) + (size_t)o * f;
        float ss = 0.f;
#pragma unroll 4
        for (int i = lane; i < f; i += 32) ss += w[i] * w[i];
        ss = warp_sum(ss);
        const float inv = 1.f / (kNormEps + sqrtf(ss) * rsqrtf((float)f));
#pragma unroll 4
        for (int i = lane; i < f; i += 32) w[i] *= inv;
    }
}

// dd_weight_prep for every parameter of a model in one launch (same arithmetic as weight_prep_kernel, elementwise.cu).
// One warp per weight row, kWpWarpRows consecutive rows per descriptor search, the (tap, i) walk without divisions.
constexpr int kWpWarps = 8, kWpWarpRows = 2;
__global__ void __launch_bounds__(kWpWarps * 32) weight_prep_batched_kernel(const dd_wprep_desc* __restrict__ descs, int n_descs,
                                                                           int total_rows) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * kWpWarps + (threadIdx.x >> 5)) * kWpWarpRows;
    if (row0 >= total_rows) return;
    int lo = 0, hi = n_descs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].row_begin <= row0) lo = mid; else hi = mid - 1;
    }
    dd_wprep_desc d = descs[lo];
    for (int r = 0; r < kWpWarpRows; ++r) {
        const int row = row0 + r;
        if (row >= total_rows) return;
        while (lo + 1 < n_descs && descs[lo + 1].row_begin <= row) d = descs[++lo];
        const int o = row - d.row_begin;
        if (o >= d.O) continue;
        const int fan_in = d.I_g * d.taps;
        const size_t base = (size_t)o * fan_in;
        const float* wf = static_cast<const float*>(d.w);
        const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(d.w);
        float scale = d.gain_host * (d.gain ? *d.gain : 1.f) * rsqrtf((float)fan_in);
        if (d.normalize) {
            float ss = 0.f;
#pragma unroll 4
            for (int i = lane; i < fan_in; i += 32) {
                const float v = d.w_is_bf16 ? __bfloat162float(wb[base + i]) : wf[base + i];
                ss += v * v;
            }
            ss = warp_sum(ss);
            scale /= (kNormEps + sqrtf(ss) * rsqrtf((float)fan_in));
        }
        int o_dst = o;
        if (d.perm == DD_WPERM_QK) {
            const int head = o / (2 * d.head_dim), rem = o % (2 * d.head_dim);
            o_dst = (rem & 1) * (d.O / 2) + head * d.head_dim + (rem >> 1);
        } else if (d.perm == DD_WPERM_QKV) {
            const int head = o / (3 * d.head_dim), rem = o % (3 * d.head_dim);
            o_dst = (rem % 3) * (d.O / 3) + head * d.head_dim + rem / 3;
        }
        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(d.out) + (size_t)o_dst * d.row_stride;
        for (int tap = 0; tap < d.taps; ++tap) {                        // dst[tap * I_g + i]: coalesced writes
#pragma unroll 4
            for (int i = lane; i < d.I_g; i += 32) {
                const size_t src = base + (size_t)i * d.taps + tap;
                dst[tap * d.I_g + i] = __float2bfloat16_rn((d.w_is_bf16 ? __bfloat162float(wb[src]) : wf[src]) * scale);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// weight-prep backward (batched over parameters): dL/dW_eff -> dL/dW (+ dL/dgain)
//   forward (mp_tools.py:359-364): w_hat = w / (eps + ||w|| / sqrt(f))  [training], w_eff = w_hat * gain / sqrt(f)
// ------------------------------------------------------------------------------------------
// One warp per weight row (fan-in <= a few thousand elements: no block-wide barriers), kWpbRows consecutive rows per warp so
// that the descriptor search -- ~log2(n_descs) dependent loads -- is paid once per kWpbRows rows; the row index -> (i, tap)
// split avoids a division by a run-time value for the two kernel sizes that occur (1 and 9 taps).
constexpr int kWpbWarps = 8, kWpbRows = 2;

__device__ __forceinline__ void split_tap(int j, int taps, int& i, int& tap) {
    if (taps == 1) { i = j; tap = 0; }
    else if (taps == 9) { i = j / 9; tap = j - i * 9; }
    else { i = j / taps; tap = j - i * taps; }
}

__global__ void __launch_bounds__(kWpbWarps * 32) weight_prep_bwd_kernel(const dd_wbwd_desc* __restrict__ descs, int n_descs,
                                                                        int total_rows) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * kWpbWarps + (threadIdx.x >> 5)) * kWpbRows;
    if (row0 >= total_rows) return;
    // find the descriptor owning the first row (row_begin is an exclusive prefix sum)
    int lo = 0, hi = n_descs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].row_begin <= row0) lo = mid; else hi = mid - 1;
    }
    dd_wbwd_desc d = descs[lo];
    for (int r = 0; r < kWpbRows; ++r) {
        const int row = row0 + r;
        if (row >= total_rows) return;
        while (lo + 1 < n_descs && descs[lo + 1].row_begin <= row) d = descs[++lo];
        const int o = row - d.row_begin;
        if (o >= d.O) continue;
        const int f = d.I_g * d.taps;
        const float rs = rsqrtf((float)f);
        int o_src = o;
        if (d.perm == DD_WPERM_QK) {
            const int head = o / (2 * d.head_dim), rem = o % (2 * d.head_dim);
            o_src = (rem & 1) * (d.O / 2) + head * d.head_dim + (rem >> 1);
        }
        const float* w = d.w + (size_t)o * f;                       // [i][tap]
        float* dw = d.dw + (size_t)o * f;
        const float gain = d.gain_host * (d.gain ? *d.gain : 1.f);
        // dL/dW_eff of (row o, tap, i): [o][tap][i], or for an operand-swapped wgrad [group*I_g + i][taps-1-tap][o % cout_g]
        const float* g = d.dweff + (size_t)o_src * d.row_stride;
        size_t s_i = 1, s_tap = (size_t)d.I_g;
        if (d.t_cout_g > 0) {
            const int grp = o / d.t_cout_g;
            const size_t row_t = (size_t)d.taps * d.t_cout_g;
            g = d.dweff + (size_t)grp * d.I_g * row_t + (size_t)(d.taps - 1) * d.t_cout_g + (o - grp * d.t_cout_g);
            s_i = row_t;
            s_tap = (size_t)0 - (size_t)d.t_cout_g;                 // (unsigned wrap: taps run backwards)
        }
        float ss = 0.f, gw = 0.f;
#pragma unroll 4
        for (int j = lane; j < f; j += 32) {                        // j = i*taps + tap (parameter order)
            int i, tap;
            split_tap(j, d.taps, i, tap);
            const float wv = w[j], gv = g[tap * s_tap + i * s_i];
            ss += wv * wv;
            gw += gv * wv;
        }
        ss = warp_sum(ss);
        gw = warp_sum(gw);
        const float nrm = sqrtf(ss);
        float a, b;                  // dw = a * G - b * w
        float dgain;
        if (d.normalize) {
            const float n = kNormEps + nrm * rs;
            a = gain * rs / n;
            b = nrm > 0.f ? (gain * rs) * gw * rs / (n * n * nrm) : 0.f;
            dgain = gw * rs / n;
        } else {
            a = gain * rs;
            b = 0.f;
            dgain = gw * rs;
        }
#pragma unroll 4
        for (int j = lane; j < f; j += 32) {
            int i, tap;
            split_tap(j, d.taps, i, tap);
            const float v = a * g[tap * s_tap + i * s_i] - b * w[j];
            dw[j] = d.accumulate ? dw[j] + v : v;
        }
        if (d.dgain && lane == 0) atomicAdd(d.dgain, dgain * d.gain_host);
    }
}

// ------------------------------------------------------------------------------------------
// y = mp_silu(pre * scale[b][c]) backward:  dpre = coef*dy*silu'(u)*scale,  dscale[b][c] += sum_pix coef*dy*silu'(u)*pre
// ------------------------------------------------------------------------------------------
// A block covers a strip of pixels x 32 channel-vectors (256 channels): threadIdx.x walks the channel vectors
// (coalesced 512 B rows), threadIdx.y the pixels; every thread keeps kPixPerThread independent loads in flight and the
// per-channel sums meet in shared memory before one atomicAdd per channel and block.
constexpr int kRedRows = 8;

__device__ __forceinline__ void strip_reduce_atomic(float (&acc)[8], float* __restrict__ dst, bool vok,
                                                    float (*red)[32][8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && vok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < kRedRows; ++r) t += red[r][threadIdx.x][j];
            atomicAdd(dst + j, t);
        }
    }
}

constexpr int kSsbPixPerThread = 4;

__global__ void __launch_bounds__(256, 2)
silu_scale_bwd_kernel(const uint4* __restrict__ dy, float coef, const uint4* __restrict__ pre,
                      const float* __restrict__ scale, uint4* __restrict__ dpre, float* __restrict__ dscale, long npix,
                      int nvec) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ float red[kRedRows][32][8];
    const int b = blockIdx.z;
    const int v = blockIdx.y * 32 + threadIdx.x;
    const bool vok = v < nvec;
    float sc[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = vok ? __ldg(scale + (size_t)b * nvec * 8 + v * 8 + j) : 0.f; acc[j] = 0.f; }
    // A CTA walks several 64-pixel strips and keeps the dscale partial sums in registers across them: one shared-memory
    // reduction and one set of atomics per CTA (one per strip made 700 k atomics onto 2048 addresses at level 0, and the
    // short CTAs spent their time in launch / drain: 1.8 TB/s).
    constexpr long kStrip = kRedRows * kSsbPixPerThread;
#pragma unroll 1
    for (long s0 = (long)blockIdx.x * kStrip; s0 < npix; s0 += (long)gridDim.x * kStrip) {
        const long p0 = s0 + threadIdx.y;
        uint4 qg[kSsbPixPerThread], qx[kSsbPixPerThread];
#pragma unroll
        for (int i = 0; i < kSsbPixPerThread; ++i) {
            const long pix = p0 + (long)i * kRedRows;
            if (vok && pix < npix) {
                const size_t off = ((size_t)b * npix + pix) * nvec + v;
                qg[i] = __ldg(dy + off);
                qx[i] = __ldg(pre + off);
            }
        }
#pragma unroll
        for (int i = 0; i < kSsbPixPerThread; ++i) {
            const long pix = p0 + (long)i * kRedRows;
            if (vok && pix < npix) {
                float g[8], x[8], o[8];
                unpack8(qg[i], g);
                unpack8(qx[i], x);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = coef * g[j] * mp_silu_grad(x[j] * sc[j]);
                    o[j] = t * sc[j];
                    acc[j] += t * x[j];
                }
                dpre[((size_t)b * npix + pix) * nvec + v] = pack8(o);
            }
        }
    }
    strip_reduce_atomic(acc, dscale + (size_t)b * nvec * 8 + v * 8, vok, red);
}

// ------------------------------------------------------------------------------------------
// encoder: xn = t0/(eps + rms_c(t0)), s = mp_silu(xn).  Given g (grad of the block's mp_sum wrt xn is ca*g) and
// ds (grad wrt s): dxn = ca*g + ds*silu'(xn);  dt0 = dxn/n - t0 * <dxn,t0> / (n^2 * C * rms)
// ------------------------------------------------------------------------------------------
constexpr int kMaxVecPerLaneB = 10;    // C <= 2560

// NV = 16-byte vectors per lane (C <= NV * 256), a template parameter: with the bound of the widest layer (10) compiled in,
// a 256-channel launch executed ten predicated copies of every pass and held 166 registers (12 warps per SM).
template <int NV>
__global__ void pixnorm_silu_bwd_kernel(const uint4* __restrict__ g, float ca, const uint4* __restrict__ ds,
                                        const uint4* __restrict__ t0, uint4* __restrict__ dt0, long npix, int C) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int lane = threadIdx.x & 31;
    const long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= npix) return;
    const int nvec = C >> 3;
    uint4 rt[NV];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + k * 32;
        if (v < nvec) {
            rt[k] = __ldg(t0 + pix * nvec + v);
            float f[8];
            unpack8(rt[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    ss = warp_sum(ss);
    const float rms = sqrtf(ss) * rsqrtf((float)C);
    const float n = kNormEps + rms, inv = 1.f / n;
    // pass 1: dxn and <dxn, t0>
    float dot = 0.f;
    uint4 rd[NV];       // dxn kept as packed fp32 pairs would double registers: keep bf16-packed dxn
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + k * 32;
        if (v < nvec) {
            float ft[8], fg[8], fs[8], dx[8];
            unpack8(rt[k], ft);
            unpack8(__ldg(g + pix * nvec + v), fg);
            unpack8(__ldg(ds + pix * nvec + v), fs);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dx[j] = ca * fg[j] + fs[j] * mp_silu_grad(ft[j] * inv);
                dot += dx[j] * ft[j];
            }
            rd[k] = pack8(dx);
        }
    }
    dot = warp_sum(dot);
    const float kb = rms > 0.f ? dot / (n * n * (float)C * rms) : 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = lane + k * 32;
        if (v < nvec) {
            float ft[8], dx[8], o[8];
            unpack8(rt[k], ft);
            unpack8(rd[k], dx);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = dx[j] * inv - ft[j] * kb;
            dt0[pix * nvec + v] = pack8(o);
        }
    }
}

// ------------------------------------------------------------------------------------------
// decoder input backward: xc = [wa*up(a), wb*b], s = mp_silu(xc).
//   dxc = c1*d_xc + d_s*silu'(xc);  da = mask(a) * wa * (sum over the 2x2 replicas of dxc[..:Ca]);  db = wb * dxc[Ca:]
// ------------------------------------------------------------------------------------------
// I = index type (unsigned when the tensors have fewer than 2^31 vectors: 64-bit div / mod is a long instruction sequence)
template <typename I>
__global__ void cat_silu_bwd_kernel(const uint4* __restrict__ d_xc, float c1, const uint4* __restrict__ d_s,
                                    const uint4* __restrict__ xc, const uint4* __restrict__ a_prev, float clip, float wa,
                                    float wb, int up, uint4* __restrict__ da, uint4* __restrict__ db, int B, int H, int W,
                                    int va, int vb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int vt = va + vb;
    const int Ha = up ? H >> 1 : H, Wa = up ? W >> 1 : W;
    const I total_a = (I)B * (I)Ha * (I)Wa * (I)va, total_b = (I)B * (I)H * (I)W * (I)vb;
    for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total_a + total_b; idx += (I)gridDim.x * blockDim.x) {
        if (idx < total_a) {
            const I apix = idx / (I)va;
            const int v = (int)(idx - apix * (I)va);
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int reps = up ? 2 : 1;
            const I row = apix / (I)Wa;
            const int w = (int)(apix - row * (I)Wa);
            const I bb = row / (I)Ha;
            const int h = (int)(row - bb * (I)Ha);
            for (int dy = 0; dy < reps; ++dy)
                for (int dx = 0; dx < reps; ++dx) {
                    const size_t pix = ((size_t)bb * H + h * reps + dy) * W + w * reps + dx;
                    float g1[8], g2[8], x[8];
                    unpack8(__ldg(d_xc + pix * vt + v), g1);
                    unpack8(__ldg(d_s + pix * vt + v), g2);
                    unpack8(__ldg(xc + pix * vt + v), x);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += c1 * g1[j] + g2[j] * mp_silu_grad(x[j]);
                }
            float o[8];
            if (a_prev) {
                float ap[8];
                unpack8(__ldg(a_prev + idx), ap);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fabsf(ap[j]) < clip ? wa * acc[j] : 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = wa * acc[j];
            }
            da[idx] = pack8(o);
        } else {
            const I k = idx - total_a;
            const size_t pix = k / (I)vb;
            const int v = (int)(k - (I)pix * (I)vb);
            float g1[8], g2[8], x[8], o[8];
            unpack8(__ldg(d_xc + pix * vt + va + v), g1);
            unpack8(__ldg(d_s + pix * vt + va + v), g2);
            unpack8(__ldg(xc + pix * vt + va + v), x);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = wb * (c1 * g1[j] + g2[j] * mp_silu_grad(x[j]));
            db[k] = pack8(o);
        }
    }
}

// encoder chain: gradient of a block output = mask(x_prev) * (avgpool-backward(dx0) or dx0, plus the skip gradient)
template <typename I>
__global__ void enc_grad_combine_kernel(const uint4* __restrict__ dx0, int down, const uint4* __restrict__ dskip,
                                        const uint4* __restrict__ x_prev, float clip, uint4* __restrict__ out, int B, int H,
                                        int W, int nvec) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const I total = (I)B * (I)H * (I)W * (I)nvec;
    for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
        float a[8], o[8];
        if (down) {
            const I pix = idx / (I)nvec;
            const int v = (int)(idx - pix * (I)nvec);
            const I row = pix / (I)W;
            const int w = (int)(pix - row * (I)W);
            const I b = row / (I)H;
            const int h = (int)(row - b * (I)H);
            unpack8(__ldg(dx0 + (((size_t)b * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) * nvec + v), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] *= 0.25f;
        } else {
            unpack8(__ldg(dx0 + idx), a);
        }
        if (dskip) {
            float s[8];
            unpack8(__ldg(dskip + idx), s);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += s[j];
        }
        if (x_prev) {
            float xp[8];
            unpack8(__ldg(x_prev + idx), xp);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fabsf(xp[j]) < clip ? a[j] : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = a[j];
        }
        out[idx] = pack8(o);
    }
}

// attention input: x2 feeds mp_sum (ca*g3), attn_v (dxv) and attn_qk through xs = x2*c_qk (dxs):
//   dx2 = ca*g3 + dxv + dxs*c_qk[b][c];   dc_qk[b][c] += sum_pix dxs*x2
constexpr int kAibPixPerThread = 4;

__global__ void __launch_bounds__(256)
attn_in_bwd_kernel(const uint4* __restrict__ g3, float ca, const uint4* __restrict__ dxv, const uint4* __restrict__ dxs,
                   const uint4* __restrict__ x2, const float* __restrict__ c_qk, uint4* __restrict__ dx2,
                   float* __restrict__ dc_qk, long npix, int nvec) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ float red[kRedRows][32][8];
    const int b = blockIdx.z;
    const int v = blockIdx.y * 32 + threadIdx.x;
    const bool vok = v < nvec;
    const long p0 = (long)blockIdx.x * (kRedRows * kAibPixPerThread) + threadIdx.y;
    float sc[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = vok ? __ldg(c_qk + (size_t)b * nvec * 8 + v * 8 + j) : 0.f; acc[j] = 0.f; }
    uint4 q0[kAibPixPerThread], q1[kAibPixPerThread], q2[kAibPixPerThread], q3[kAibPixPerThread];
#pragma unroll
    for (int i = 0; i < kAibPixPerThread; ++i) {
        const long pix = p0 + (long)i * kRedRows;
        if (vok && pix < npix) {
            const size_t off = ((size_t)b * npix + pix) * nvec + v;
            q0[i] = __ldg(g3 + off); q1[i] = __ldg(dxv + off); q2[i] = __ldg(dxs + off); q3[i] = __ldg(x2 + off);
        }
    }
#pragma unroll
    for (int i = 0; i < kAibPixPerThread; ++i) {
        const long pix = p0 + (long)i * kRedRows;
        if (vok && pix < npix) {
            float g[8], dv[8], dsx[8], x[8], o[8];
            unpack8(q0[i], g); unpack8(q1[i], dv); unpack8(q2[i], dsx); unpack8(q3[i], x);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] = ca * g[j] + dv[j] + dsx[j] * sc[j];
                acc[j] += dsx[j] * x[j];
            }
            dx2[((size_t)b * npix + pix) * nvec + v] = pack8(o);
        }
    }
    strip_reduce_atomic(acc, dc_qk + (size_t)b * nvec * 8 + v * 8, vok, red);
}

// ------------------------------------------------------------------------------------------
// attention backward (unet_edm2_b4.py:137-148): q,k,v cosine-normalised over the 64 head channels, P = softmax(q k^T/8),
// a = P v.  FlashAttention-2 style on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate), two kernels:
//   kernel Q  (per 64-query tile; K^, V^ of the head resident in shared memory):
//       pass 1: L_i = logsumexp_j s_ij;  D_i = <da_i, a_i>
//       pass 2: P = exp(s - L), dP = dA V^^T, dS = P o (dP - D)/8, dQ^ += dS K^        -> dq (norm backward), stats (L, D)
//   kernel KV (per 64-key tile; Q^, dA of the head resident):
//       S^T = K^ Q^^T, P^T = exp(S^T - L), dP^T = V^ dA^T, dS^T = P^T o (dP^T - D)/8,
//       dV^ += P^T dA, dK^ += dS^T Q^                                                  -> dk, dv (norm backward)
// P and dS are rounded to bf16 before their second MMA, q^/k^/v^/dA are bf16 in shared memory (as in the forward).
// ------------------------------------------------------------------------------------------
constexpr int kAD = 64;
constexpr int kBT = 64;                 // token rows per CTA (4 warps x 16)
constexpr int kBwdThreads = 128;
constexpr int kRS = kAD + 8;            // bf16 elements per shared-memory row (conflict-free fragment loads)
constexpr float kSl2 = 0.125f * 1.44269504089f;   // 1/sqrt(64) * log2(e)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// Stage rows [first, first + n_rows) of one head (64 channels, token stride `tok_stride`) into dst[n_rows][kRS] as bf16,
// optionally cosine-normalised; rows >= limit are zero.  128 threads: 8 lanes per token, 16 tokens per pass.
__device__ __forceinline__ void stage_rows_bwd(const __nv_bfloat16* __restrict__ head_base, size_t tok_stride, int first,
                                               int n_rows, int limit, __nv_bfloat16* __restrict__ dst, bool normalize) {
    const int sub = threadIdx.x & 7, tok = threadIdx.x >> 3;
    constexpr int kBatch = 12;          // loads in flight per thread: staging is a chain of memory round trips otherwise
    for (int r0 = 0; r0 < n_rows; r0 += 16 * kBatch) {
        uint4 q[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int r = r0 + u * 16 + tok;
            q[u] = make_uint4(0, 0, 0, 0);
            if (r < n_rows && first + r < limit)
                q[u] = __ldg(reinterpret_cast<const uint4*>(head_base + (size_t)(first + r) * tok_stride + sub * 8));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int r = r0 + u * 16 + tok;
            if (r0 + u * 16 >= n_rows) break;                 // warp-uniform
            float f[8];
            unpack8(q[u], f);
            float inv = 1.f;
            if (normalize) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                ss += __shfl_xor_sync(0xffffffffu, ss, 4);
                inv = 1.f / (kNormEps + sqrtf(ss) * 0.125f);
            }
            if (r < n_rows) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] *= inv;
                *reinterpret_cast<uint4*>(dst + (size_t)r * kRS + sub * 8) = pack8(f);
            }
        }
    }
}

// A-operand fragments (16 rows x 64 channels = 4 k-steps) of rows [row0, row0+16) of a [.][kRS] tile
__device__ __forceinline__ void load_a_frags(const __nv_bfloat16* tile, int row0, int g, int t, uint32_t (&a)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const __nv_bfloat16* p0 = tile + (size_t)(row0 + g) * kRS + kk * 16 + 2 * t;
        const __nv_bfloat16* p1 = p0 + 8 * kRS;
        a[kk][0] = *reinterpret_cast<const uint32_t*>(p0);
        a[kk][1] = *reinterpret_cast<const uint32_t*>(p1);
        a[kk][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        a[kk][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
    }
}

// acc[16 x 64] = A[16 x 64ch] * M[rows r0..r0+63][64ch]^T   (M row-major, channels contiguous: the "n = row, k = channel" form)
__device__ __forceinline__ void mma_a_rowsT(float (&acc)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* m, int r0,
                                            int g, int t) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        const __nv_bfloat16* bp = m + (size_t)(r0 + n * 8 + g) * kRS + 2 * t;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            mma16816(acc[n], a[kk], *reinterpret_cast<const uint32_t*>(bp + kk * 16),
                     *reinterpret_cast<const uint32_t*>(bp + kk * 16 + 8));
    }
}

// out[16 x 64ch] += F[16 x 64 rows] * M[rows r0..r0+63][64ch]   (F given as 4 k-step A fragments; M^T fragments via ldmatrix.trans)
__device__ __forceinline__ void mma_frag_rows(float (&out)[8][4], const uint32_t (&f)[4][4], const __nv_bfloat16* m, int r0,
                                              int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const __nv_bfloat16* row = m + (size_t)(r0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kRS + (lane >> 4) * 8;
#pragma unroll
        for (int n = 0; n < 8; n += 2) {
            uint32_t vb[4];
            ldsm_x4_trans(vb, row + n * 8);
            mma16816(out[n], f[kk], vb[0], vb[1]);
            mma16816(out[n + 1], f[kk], vb[2], vb[3]);
        }
    }
}

// Backward of x^ = x / (eps + ||x||/8) for the two rows (g, g+8) a thread co-owns: x and dx^ in the accumulator layout
// (cols n*8 + 2t, +1), writes dx (bf16) to dst rows.  Row sums are reduced over the 4 lanes of a quad.
__device__ __forceinline__ void norm_bwd_store(const __nv_bfloat16* __restrict__ x_base, __nv_bfloat16* __restrict__ dst_base,
                                               size_t tok_stride_x, size_t tok_stride_dst, int row_a, int N, int t,
                                               const float (&dxh)[8][4]) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = row_a + r * 8;
        const bool ok = row < N;
        float x[16];
        float ss = 0.f, dot = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float2 v = make_float2(0.f, 0.f);
            if (ok) v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x_base + (size_t)row * tok_stride_x + n * 8 + 2 * t));
            x[2 * n] = v.x; x[2 * n + 1] = v.y;
            ss += v.x * v.x + v.y * v.y;
            dot += v.x * dxh[n][2 * r] + v.y * dxh[n][2 * r + 1];
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1); dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        const float nrm = sqrtf(ss);
        const float nn = kNormEps + nrm * 0.125f;
        const float a = 1.f / nn, b = nrm > 0.f ? dot * 0.125f / (nn * nn * nrm) : 0.f;
        if (ok) {
#pragma unroll
            for (int n = 0; n < 8; ++n)
                *reinterpret_cast<uint32_t*>(dst_base + (size_t)row * tok_stride_dst + n * 8 + 2 * t) =
                    pack_bf16x2(dxh[n][2 * r] * a - x[2 * n] * b, dxh[n][2 * r + 1] * a - x[2 * n + 1] * b);
        }
    }
}

__global__ void __launch_bounds__(kBwdThreads)
attention_bwd_q_kernel(const __nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ v,
                       const __nv_bfloat16* __restrict__ a_raw, const __nv_bfloat16* __restrict__ da,
                       __nv_bfloat16* __restrict__ dqk, float* __restrict__ stats, int N, int heads, int npad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_bwd);        // [npad][kRS]  k^
    __nv_bfloat16* Vs = Ks + (size_t)npad * kRS;                           // [npad][kRS]  v^
    __nv_bfloat16* Qs = Vs + (size_t)npad * kRS;                           // [kBT][kRS]   q^ tile
    __nv_bfloat16* Gs = Qs + (size_t)kBT * kRS;                            // [kBT][kRS]   dA tile
    const int C = heads * kAD;
    const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * kBT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* q_base = qk + (size_t)b * N * 2 * C + head * kAD;
    const __nv_bfloat16* a_base = a_raw + (size_t)b * N * C + head * kAD;
    const __nv_bfloat16* g_base = da + (size_t)b * N * C + head * kAD;
    stage_rows_bwd(q_base + C, 2 * (size_t)C, 0, npad, N, Ks, true);
    stage_rows_bwd(v + (size_t)b * N * C + head * kAD, (size_t)C, 0, npad, N, Vs, true);
    stage_rows_bwd(q_base, 2 * (size_t)C, q0, kBT, N, Qs, true);
    stage_rows_bwd(g_base, (size_t)C, q0, kBT, N, Gs, false);
    __syncthreads();

    const int row0 = warp * 16, row_a = q0 + row0 + g;
    uint32_t qa[4][4], ga[4][4];
    load_a_frags(Qs, row0, g, t, qa);
    load_a_frags(Gs, row0, g, t, ga);
    // D_i = <da_i, a_i>
    float D[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = row_a + r * 8;
        if (row < N) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float2 av = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(a_base + (size_t)row * C + n * 8 + 2 * t));
                const float2 gv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(g_base + (size_t)row * C + n * 8 + 2 * t));
                D[r] += av.x * gv.x + av.y * gv.y;
            }
        }
        D[r] += __shfl_xor_sync(0xffffffffu, D[r], 1);
        D[r] += __shfl_xor_sync(0xffffffffu, D[r], 2);
    }
    // pass 1: log-sum-exp of the score rows (log2 domain)
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int kb = 0; kb < npad; kb += 64) {
        float s[8][4];
        mma_a_rowsT(s, qa, Ks, kb, g, t);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int key = kb + n * 8 + 2 * t;
            if (key >= N) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
            if (key + 1 >= N) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);       // finite: every 64-key block holds >= 1 real key
            l_run[r] *= fast_exp2((m_run[r] - m_new) * kSl2);
            m_run[r] = m_new;
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            l_run[0] += fast_exp2((s[n][0] - m_run[0]) * kSl2) + fast_exp2((s[n][1] - m_run[0]) * kSl2);
            l_run[1] += fast_exp2((s[n][2] - m_run[1]) * kSl2) + fast_exp2((s[n][3] - m_run[1]) * kSl2);
        }
    }
    float L2[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
        L2[r] = m_run[r] * kSl2 + log2f(l_run[r]);
    }
    // pass 2: dQ^
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
    for (int kb = 0; kb < npad; kb += 64) {
        float s[8][4], dp[8][4];
        mma_a_rowsT(s, qa, Ks, kb, g, t);
        mma_a_rowsT(dp, ga, Vs, kb, g, t);
        uint32_t dsa[4][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int key = kb + n * 8 + 2 * t;
            const bool k0 = key < N, k1 = key + 1 < N;
            const float p0 = k0 ? fast_exp2(s[n][0] * kSl2 - L2[0]) : 0.f, p1 = k1 ? fast_exp2(s[n][1] * kSl2 - L2[0]) : 0.f;
            const float p2 = k0 ? fast_exp2(s[n][2] * kSl2 - L2[1]) : 0.f, p3 = k1 ? fast_exp2(s[n][3] * kSl2 - L2[1]) : 0.f;
            dsa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0 * (dp[n][0] - D[0]) * 0.125f, p1 * (dp[n][1] - D[0]) * 0.125f);
            dsa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2 * (dp[n][2] - D[1]) * 0.125f, p3 * (dp[n][3] - D[1]) * 0.125f);
        }
        mma_frag_rows(dq, dsa, Ks, kb, lane);
    }
    norm_bwd_store(q_base, dqk + (size_t)b * N * 2 * C + head * kAD, 2 * (size_t)C, 2 * (size_t)C, row_a, N, t, dq);
    if (t == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = row_a + r * 8;
            if (row < N) {
                float* st = stats + (((size_t)b * heads + head) * N + row) * 2;
                st[0] = L2[r];
                st[1] = D[r];
            }
        }
    }
}

__global__ void __launch_bounds__(kBwdThreads)
attention_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qk, const __nv_bfloat16* __restrict__ v,
                        const __nv_bfloat16* __restrict__ da, const float* __restrict__ stats,
                        __nv_bfloat16* __restrict__ dqk, __nv_bfloat16* __restrict__ dv_out, int N, int heads, int npad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_bwd);        // [npad][kRS]  q^ (all queries)
    __nv_bfloat16* Gs = Qs + (size_t)npad * kRS;                           // [npad][kRS]  dA
    __nv_bfloat16* Kt = Gs + (size_t)npad * kRS;                           // [kBT][kRS]   k^ tile
    __nv_bfloat16* Vt = Kt + (size_t)kBT * kRS;                            // [kBT][kRS]   v^ tile
    float* Ls = reinterpret_cast<float*>(Vt + (size_t)kBT * kRS);          // [npad]
    float* Ds = Ls + npad;                                                 // [npad]
    const int C = heads * kAD;
    const int head = blockIdx.y, b = blockIdx.z, k0 = blockIdx.x * kBT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* q_base = qk + (size_t)b * N * 2 * C + head * kAD;
    const __nv_bfloat16* v_base = v + (size_t)b * N * C + head * kAD;
    const float* st = stats + ((size_t)b * heads + head) * N * 2;
    stage_rows_bwd(q_base, 2 * (size_t)C, 0, npad, N, Qs, true);
    stage_rows_bwd(da + (size_t)b * N * C + head * kAD, (size_t)C, 0, npad, N, Gs, false);
    stage_rows_bwd(q_base + C, 2 * (size_t)C, k0, kBT, N, Kt, true);
    stage_rows_bwd(v_base, (size_t)C, k0, kBT, N, Vt, true);
    for (int i = threadIdx.x; i < npad; i += kBwdThreads) {
        Ls[i] = i < N ? st[2 * i] : INFINITY;            // padded queries: P = exp2(. - inf) = 0
        Ds[i] = i < N ? st[2 * i + 1] : 0.f;
    }
    __syncthreads();

    const int row0 = warp * 16, row_a = k0 + row0 + g;
    uint32_t ka[4][4], va[4][4];
    load_a_frags(Kt, row0, g, t, ka);
    load_a_frags(Vt, row0, g, t, va);
    float dk[8][4], dvv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dvv[n][0] = dvv[n][1] = dvv[n][2] = dvv[n][3] = 0.f; }
    for (int qb = 0; qb < npad; qb += 64) {
        float s[8][4], dp[8][4];
        mma_a_rowsT(s, ka, Qs, qb, g, t);             // S^T: rows = keys, cols = queries
        mma_a_rowsT(dp, va, Gs, qb, g, t);            // dP^T
        uint32_t pa[4][4], dsa[4][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int i = qb + n * 8 + 2 * t;
            const float l0 = Ls[i], l1 = Ls[i + 1], d0 = Ds[i], d1 = Ds[i + 1];
            const float p0 = fast_exp2(s[n][0] * kSl2 - l0), p1 = fast_exp2(s[n][1] * kSl2 - l1);
            const float p2 = fast_exp2(s[n][2] * kSl2 - l0), p3 = fast_exp2(s[n][3] * kSl2 - l1);
            pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
            dsa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0 * (dp[n][0] - d0) * 0.125f, p1 * (dp[n][1] - d1) * 0.125f);
            dsa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2 * (dp[n][2] - d0) * 0.125f, p3 * (dp[n][3] - d1) * 0.125f);
        }
        mma_frag_rows(dvv, pa, Gs, qb, lane);         // dV^ += P^T dA
        mma_frag_rows(dk, dsa, Qs, qb, lane);         // dK^ += dS^T Q^
    }
    norm_bwd_store(q_base + C, dqk + (size_t)b * N * 2 * C + C + head * kAD, 2 * (size_t)C, 2 * (size_t)C, row_a, N, t, dk);
    norm_bwd_store(v_base, dv_out + (size_t)b * N * C + head * kAD, (size_t)C, (size_t)C, row_a, N, t, dvv);
}

// ------------------------------------------------------------------------------------------
// embedding heads
// ------------------------------------------------------------------------------------------
// emb_linear*: c[b][o] = bias + gain * sum_i w_hat[o][i] emb[b][g*I+i] / sqrt(I).  Stage 1 (warp per row): row scale and
// dW_eff[o][i] = sum_b dc[b][o] emb[b][g*I+i].
__global__ void emb_affine_bwd_w_kernel(const dd_affine_bwd_desc* __restrict__ descs, const float* __restrict__ emb, int B,
                                        int cemb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const dd_affine_bwd_desc d = descs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= d.O) return;
    const int g = o / (d.O / d.groups);
    const float* e0 = emb + (size_t)g * d.I;
    const bool vec = d.I % 4 == 0 && cemb % 4 == 0 && ((reinterpret_cast<uintptr_t>(d.w) | reinterpret_cast<uintptr_t>(d.dweff) |
                                                       reinterpret_cast<uintptr_t>(e0)) & 15u) == 0;
    float ss = 0.f;
    if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(d.w + (size_t)o * d.I);
        for (int i = lane; i < d.I / 4; i += 32) { const float4 q = __ldg(w4 + i); ss += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
    } else {
        for (int i = lane; i < d.I; i += 32) { const float wv = d.w[(size_t)o * d.I + i]; ss += wv * wv; }
    }
    ss = warp_sum(ss);
    float scale = (d.gain ? *d.gain : 1.f) * rsqrtf((float)d.I);
    if (d.normalize) scale /= (kNormEps + sqrtf(ss) * rsqrtf((float)d.I));
    if (lane == 0) d.rowscale[o] = scale;
    if (vec) {
        float4* g4 = reinterpret_cast<float4*>(d.dweff + (size_t)o * d.I);
        for (int i = lane; i < d.I / 4; i += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int b = 0; b < B; ++b) {
                const float dc = d.dout[(size_t)b * d.O + o];
                const float4 e = __ldg(reinterpret_cast<const float4*>(e0 + (size_t)b * cemb) + i);
                acc.x += dc * e.x; acc.y += dc * e.y; acc.z += dc * e.z; acc.w += dc * e.w;
            }
            g4[i] = acc;
        }
    } else {
        for (int i = lane; i < d.I; i += 32) {
            float acc = 0.f;
            for (int b = 0; b < B; ++b) acc += d.dout[(size_t)b * d.O + o] * e0[(size_t)b * cemb + i];
            d.dweff[(size_t)o * d.I + i] = acc;
        }
    }
}

// Stage 2: demb[b][g*I+i] += sum_o dc[b][o] * rowscale[o] * w[o][i]   (thread per input column, rows chunked over blockIdx.z)
__global__ void emb_affine_bwd_x_kernel(const dd_affine_bwd_desc* __restrict__ descs, float* __restrict__ demb, int B,
                                        int cemb, int row_chunks) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const dd_affine_bwd_desc d = descs[blockIdx.y];
    const int col = blockIdx.x * blockDim.x + threadIdx.x;       // g*I + i
    if (col >= d.groups * d.I) return;
    const int g = col / d.I, i = col - g * d.I;
    const int rows_g = d.O / d.groups;
    const int r0 = (int)((long)blockIdx.z * rows_g / row_chunks), r1 = (int)((long)(blockIdx.z + 1) * rows_g / row_chunks);
    for (int b0 = 0; b0 < B; b0 += 8) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int r = r0;
        for (; r + 4 <= r1; r += 4) {                              // four weight rows in flight per thread
            const int o = g * rows_g + r;
            float wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = __ldg(d.w + (size_t)(o + u) * d.I + i);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float ws = wv[u] * d.rowscale[o + u];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (b0 + j < B) acc[j] += d.dout[(size_t)(b0 + j) * d.O + o + u] * ws;
            }
        }
        for (; r < r1; ++r) {
            const int o = g * rows_g + r;
            const float wv = d.w[(size_t)o * d.I + i] * d.rowscale[o];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (b0 + j < B) acc[j] += d.dout[(size_t)(b0 + j) * d.O + o] * wv;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (b0 + j < B) atomicAdd(demb + (size_t)(b0 + j) * cemb + col, acc[j]);
    }
}

// emb[b][o] = mp_silu(m), m = (e + t*(l - e))/nt, e = w_eff[o] . fourier(ln(sigma_b)/4)   (unet_edm2_b4.py:273-276)
__global__ void noise_embedding_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ freqs,
                                           const float* __restrict__ phases, int cnoise, const float* __restrict__ w,
                                           int normalize, const float* __restrict__ label, float t,
                                           const float* __restrict__ demb, float* __restrict__ dweff,
                                           float* __restrict__ dlabel, int B, int cemb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    extern __shared__ float four[];        // [B][cnoise]
    for (int k = threadIdx.x; k < B * cnoise; k += blockDim.x) {
        const int b = k / cnoise, i = k - b * cnoise;
        four[k] = cosf(logf(sigma[b]) * 0.25f * freqs[i] + phases[i]) * 1.41421356237f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= cemb) return;
    float ss = 0.f;
    for (int i = lane; i < cnoise; i += 32) { const float wv = w[(size_t)o * cnoise + i]; ss += wv * wv; }
    ss = warp_sum(ss);
    float scale = rsqrtf((float)cnoise);
    if (normalize) scale /= (kNormEps + sqrtf(ss) * rsqrtf((float)cnoise));
    const float nt = rsqrtf((1.f - t) * (1.f - t) + t * t);
    for (int i0 = 0; i0 < cnoise; i0 += 32) dweff[(size_t)o * cnoise + i0 + lane] = 0.f;   // cnoise % 32 == 0 (checked)
    for (int b = 0; b < B; ++b) {
        float acc = 0.f;
        for (int i = lane; i < cnoise; i += 32) acc += w[(size_t)o * cnoise + i] * four[b * cnoise + i];
        acc = warp_sum(acc);
        const float e = acc * scale;
        const float l = label[(size_t)b * cemb + o];
        const float m = (e + t * (l - e)) * nt;
        const float dm = demb[(size_t)b * cemb + o] * mp_silu_grad(m);
        const float de = dm * (1.f - t) * nt;
        if (lane == 0) dlabel[(size_t)b * cemb + o] = dm * t * nt;
        for (int i = lane; i < cnoise; i += 32) dweff[(size_t)o * cnoise + i] += de * four[b * cnoise + i];
    }
}

// get_embeddings backward (unet_edm2_b4.py:232-235): out[b][o] = (u + t_b (c - u))/n_b, c = w_l_eff[o] . normalize(emb_in[b]),
// u = w_u_eff[o].  Writes dL/dw_l_eff [cemb][I] and dL/dw_u_eff [cemb].
__global__ void label_embedding_bwd_kernel(const float* __restrict__ emb_in, int Bc, int I, const float* __restrict__ mask,
                                           int Bm, const float* __restrict__ dout, float* __restrict__ dweff_label,
                                           float* __restrict__ dweff_uncond, int cemb) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    extern __shared__ float ehat[];        // [Bc][I]
    __shared__ float red[32];
    for (int b = 0; b < Bc; ++b) {
        float ss = 0.f;
        for (int i = threadIdx.x; i < I; i += blockDim.x) { const float v = emb_in[(size_t)b * I + i]; ss += v * v; }
        ss = block_sum_b(ss, red);
        const float inv = 1.f / (kNormEps + sqrtf(ss) * rsqrtf((float)I));
        for (int i = threadIdx.x; i < I; i += blockDim.x) ehat[b * I + i] = emb_in[(size_t)b * I + i] * inv;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= cemb) return;
    float du = 0.f;
    for (int i = lane; i < I; i += 32) {
        float acc = 0.f;
        for (int b = 0; b < Bm; ++b) {
            const float t = mask[b];
            const float nt = rsqrtf((1.f - t) * (1.f - t) + t * t);
            acc += dout[(size_t)b * cemb + o] * t * nt * ehat[(Bc == 1 ? 0 : b) * I + i];
        }
        dweff_label[(size_t)o * I + i] = acc;
    }
    if (lane == 0) {
        for (int b = 0; b < Bm; ++b) {
            const float t = mask[b];
            du += dout[(size_t)b * cemb + o] * (1.f - t) * rsqrtf((1.f - t) * (1.f - t) + t * t);
        }
        dweff_uncond[o] = du;
    }
}

// get_sigma_loss_logvar backward (:237-238): out[i] = w . fourier(ln(sigma_i)/4) / sqrt(n)
__global__ void logvar_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ freqs,
                                  const float* __restrict__ phases, int n, const float* __restrict__ dout, int count,
                                  float* __restrict__ dw, int accumulate) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float acc = 0.f;
    for (int i = 0; i < count; ++i) acc += dout[i] * cosf(logf(sigma[i]) * 0.25f * freqs[j] + phases[j]);
    acc *= 1.41421356237f * rsqrtf((float)n);
    dw[j] = accumulate ? dw[j] + acc : acc;
}

// UNet head backward: D = c_skip*x_in + c_out*F (optionally D' = mp_sum(x_ref[:, :-1], D, t = x_ref[:, -1:])), fp32 NCHW.
// Writes dF as NHWC bf16 with the channel dimension zero-padded to `Cpad` (the dgrad / wgrad kernels' operand).
__global__ void head_grad_kernel(const float* __restrict__ dD, const float* __restrict__ sigma, float sigma_data,
                                 const float* __restrict__ x_ref, __nv_bfloat16* __restrict__ dF, int B, int Cout, int H,
                                 int W, int Cpad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const long total = (long)B * H * W * Cpad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cpad);
        const long pix = idx / Cpad;
        float val = 0.f;
        if (c < Cout) {
            const long hw = pix % ((long)H * W);
            const int b = (int)(pix / ((long)H * W));
            const float sg = sigma[b];
            const float c_out = sg * sigma_data * rsqrtf(sg * sg + sigma_data * sigma_data);
            val = dD[((size_t)b * Cout + c) * H * W + hw] * c_out;
            if (x_ref) {
                const float tt = x_ref[((size_t)b * (Cout + 1) + Cout) * H * W + hw];
                val *= tt * rsqrtf((1.f - tt) * (1.f - tt) + tt * tt);
            }
        }
        dF[idx] = __float2bfloat16_rn(val);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int dd_weight_transpose(const void* w_prepped, void* out, int Cout, int cin_g, int taps, int groups,
                                   void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(w_prepped && out && Cout > 0 && cin_g > 0 && taps > 0 && groups > 0 && Cout % groups == 0,
               "dd_weight_transpose: bad arguments");
    const int cout_g = Cout / groups;
    DD_REQUIRE((long)groups * taps <= 65535, "dd_weight_transpose: groups*taps exceeds the grid limit");
    const dim3 grid(ceil_div(cin_g, 32), ceil_div(cout_g, 32), groups * taps);
    DD_CHECK_CUDA(dd_launch_pdl(weight_transpose_kernel, dim3(grid), dim3(dim3(32, 8)), 0, stream, static_cast<const __nv_bfloat16*>(w_prepped),
                                                             static_cast<__nv_bfloat16*>(out), cout_g, cin_g, taps));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_weight_normalize_batched(const dd_wprep_desc* descs_dev, int n_descs, int total_rows, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && n_descs > 0 && total_rows > 0, "dd_weight_normalize_batched: bad arguments");
    DD_CHECK_CUDA(dd_launch_pdl(weight_normalize_batched_kernel, dim3(ceil_div(total_rows, 16)), dim3(256), 0, stream, descs_dev, n_descs, total_rows));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_weight_prep_batched(const dd_wprep_desc* descs_dev, int n_descs, int total_rows, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && n_descs > 0 && total_rows > 0, "dd_weight_prep_batched: bad arguments");
    DD_CHECK_CUDA(dd_launch_pdl(weight_prep_batched_kernel, dim3(ceil_div(total_rows, kWpWarps * kWpWarpRows)), dim3(kWpWarps * 32), 0, stream, descs_dev,
                                n_descs, total_rows));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_weight_transpose_batched(const dd_wtrans_desc* descs_dev, int n_descs, int total_tiles, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && n_descs > 0 && total_tiles > 0, "dd_weight_transpose_batched: bad arguments");
    DD_CHECK_CUDA(dd_launch_pdl(weight_transpose_batched_kernel, dim3(total_tiles), dim3(dim3(32, 8)), 0, stream, descs_dev, n_descs));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_weight_prep_bwd(const dd_wbwd_desc* descs_dev, int n_descs, int total_rows, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && n_descs > 0 && total_rows > 0, "dd_weight_prep_bwd: bad arguments");
    DD_CHECK_CUDA(dd_launch_pdl(weight_prep_bwd_kernel, dim3(ceil_div(total_rows, kWpbWarps * kWpbRows)), dim3(kWpbWarps * 32), 0, stream, descs_dev,
                                n_descs, total_rows));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_silu_scale_bwd(const void* dy, float coef, const void* pre, const float* scale, void* dpre,
                                 float* dscale, int B, long npix, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(dy && pre && scale && dpre && dscale, "dd_silu_scale_bwd: null pointer");
    DD_REQUIRE(C % 8 == 0 && B > 0 && B <= 65535, "dd_silu_scale_bwd: bad shape");
    if (npix == 0) return 0;
    const int strip = kRedRows * kSsbPixPerThread;
    const int ny = ceil_div(C / 8, 32);
    const long strips = (npix + strip - 1) / strip;
    const long want = std::max<long>(1, (long)dd_num_sms() * 4 / ((long)ny * B));      // ~4 CTAs per SM over the whole grid
    const dim3 grid((unsigned)std::min<long>(strips, want), ny, B);
    DD_CHECK_CUDA(dd_launch_pdl(silu_scale_bwd_kernel, dim3(grid), dim3(dim3(32, kRedRows)), 0, stream, static_cast<const uint4*>(dy), coef,
                                                                   static_cast<const uint4*>(pre), scale,
                                                                   static_cast<uint4*>(dpre), dscale, npix, C / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_pixnorm_silu_bwd(const void* g, float ca, const void* ds, const void* t0, void* dt0, long npix, int C,
                                   void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(g && ds && t0 && dt0, "dd_pixnorm_silu_bwd: null pointer");
    DD_REQUIRE(C % 8 == 0 && C <= kMaxVecPerLaneB * 256, "dd_pixnorm_silu_bwd: C=%d unsupported", C);
    if (npix == 0) return 0;
    const int warps = 4;
    const int nv = ceil_div(C / 8, 32);
    auto kernel = nv <= 1 ? pixnorm_silu_bwd_kernel<1> : nv <= 2 ? pixnorm_silu_bwd_kernel<2> : nv <= 3 ? pixnorm_silu_bwd_kernel<3>
                : nv <= 4 ? pixnorm_silu_bwd_kernel<4> : nv <= 5 ? pixnorm_silu_bwd_kernel<5> : nv <= 6 ? pixnorm_silu_bwd_kernel<6>
                : nv <= 8 ? pixnorm_silu_bwd_kernel<8> : pixnorm_silu_bwd_kernel<kMaxVecPerLaneB>;
    DD_CHECK_CUDA(dd_launch_pdl(kernel, dim3((unsigned)((npix + warps - 1) / warps)), dim3(warps * 32), 0, stream,
        static_cast<const uint4*>(g), ca, static_cast<const uint4*>(ds), static_cast<const uint4*>(t0),
        static_cast<uint4*>(dt0), npix, C));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_cat_silu_bwd(const void* d_xc, float c1, const void* d_s, const void* xc, const void* a_prev, float clip,
                               float wa, float wb, int upsample, void* da, void* db, int B, int H, int W, int Ca, int Cb,
                               void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(d_xc && d_s && xc && da, "dd_cat_silu_bwd: null pointer");
    DD_REQUIRE(Ca > 0 && Ca % 8 == 0 && Cb % 8 == 0 && (Cb == 0 || db), "dd_cat_silu_bwd: bad arguments");
    DD_REQUIRE(!upsample || (H % 2 == 0 && W % 2 == 0), "dd_cat_silu_bwd: upsample needs even output size");
    const long total = (long)B * (upsample ? H / 2 : H) * (upsample ? W / 2 : W) * (Ca / 8) + (long)B * H * W * (Cb / 8);
    if (total == 0) return 0;
    const int grid = grid_for_b(total, 256);
    const bool narrow = total < (1L << 31) - (long)grid * 256;      // idx + stride must not wrap
    DD_CHECK_CUDA(dd_launch_pdl(narrow ? cat_silu_bwd_kernel<unsigned> : cat_silu_bwd_kernel<long>, dim3(grid), dim3(256), 0, stream,
        static_cast<const uint4*>(d_xc), c1, static_cast<const uint4*>(d_s), static_cast<const uint4*>(xc),
        static_cast<const uint4*>(a_prev), clip > 0.f ? clip : INFINITY, wa, wb, upsample, static_cast<uint4*>(da),
        static_cast<uint4*>(db), B, H, W, Ca / 8, Cb / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_enc_grad_combine(const void* dx0, int down, const void* dskip, const void* x_prev, float clip, void* out,
                                   int B, int H, int W, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(dx0 && out && C % 8 == 0, "dd_enc_grad_combine: bad arguments");
    DD_REQUIRE(!down || (H % 2 == 0 && W % 2 == 0), "dd_enc_grad_combine: downsampled block needs even size");
    const long total = (long)B * H * W * (C / 8);
    if (total == 0) return 0;
    const int grid = grid_for_b(total, 256);
    const bool narrow = total < (1L << 31) - (long)grid * 256;      // idx + stride must not wrap
    DD_CHECK_CUDA(dd_launch_pdl(narrow ? enc_grad_combine_kernel<unsigned> : enc_grad_combine_kernel<long>, dim3(grid), dim3(256), 0, stream,
        static_cast<const uint4*>(dx0), down, static_cast<const uint4*>(dskip), static_cast<const uint4*>(x_prev),
        clip > 0.f ? clip : INFINITY, static_cast<uint4*>(out), B, H, W, C / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_attn_in_bwd(const void* g3, float ca, const void* dxv, const void* dxs, const void* x2,
                              const float* c_qk, void* dx2, float* dc_qk, int B, long npix, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(g3 && dxv && dxs && x2 && c_qk && dx2 && dc_qk, "dd_attn_in_bwd: null pointer");
    DD_REQUIRE(C % 8 == 0 && B > 0 && B <= 65535, "dd_attn_in_bwd: bad shape");
    if (npix == 0) return 0;
    const int strip = kRedRows * kAibPixPerThread;
    const dim3 grid((unsigned)((npix + strip - 1) / strip), ceil_div(C / 8, 32), B);
    DD_CHECK_CUDA(dd_launch_pdl(attn_in_bwd_kernel, dim3(grid), dim3(dim3(32, kRedRows)), 0, stream, static_cast<const uint4*>(g3), ca,
                                                                static_cast<const uint4*>(dxv),
                                                                static_cast<const uint4*>(dxs),
                                                                static_cast<const uint4*>(x2), c_qk,
                                                                static_cast<uint4*>(dx2), dc_qk, npix, C / 8));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_attention_bwd(const void* qk, const void* v, const void* a_raw, const void* d_a, void* dqk, void* dv,
                                float* stats_ws, int B, int N, int heads, int head_dim, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(qk && v && a_raw && d_a && dqk && dv && stats_ws, "dd_attention_bwd: null pointer");
    DD_REQUIRE(head_dim == kAD, "dd_attention_bwd: head_dim=%d unsupported (64)", head_dim);
    DD_REQUIRE(N > 0 && B > 0 && B <= 65535 && heads <= 65535, "dd_attention_bwd: bad shape");
    const int npad = ceil_div(N, 64) * 64;
    const size_t smem = ((size_t)2 * npad * kRS + (size_t)2 * kBT * kRS) * 2 + (size_t)2 * npad * sizeof(float);
    DD_REQUIRE(smem <= 227 * 1024, "dd_attention_bwd: sequence length %d exceeds the shared-memory resident design", N);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DD_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const dim3 grid(ceil_div(N, kBT), heads, B);
    DD_CHECK_CUDA(dd_launch_pdl(attention_bwd_q_kernel, dim3(grid), dim3(kBwdThreads), smem, stream, static_cast<const __nv_bfloat16*>(qk),
                                                                static_cast<const __nv_bfloat16*>(v),
                                                                static_cast<const __nv_bfloat16*>(a_raw),
                                                                static_cast<const __nv_bfloat16*>(d_a),
                                                                static_cast<__nv_bfloat16*>(dqk), stats_ws, N, heads, npad));
    DD_CHECK_LAUNCH();
    DD_CHECK_CUDA(dd_launch_pdl(attention_bwd_kv_kernel, dim3(grid), dim3(kBwdThreads), smem, stream, static_cast<const __nv_bfloat16*>(qk),
                                                                 static_cast<const __nv_bfloat16*>(v),
                                                                 static_cast<const __nv_bfloat16*>(d_a), stats_ws,
                                                                 static_cast<__nv_bfloat16*>(dqk),
                                                                 static_cast<__nv_bfloat16*>(dv), N, heads, npad));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_emb_affine_bwd(const dd_affine_bwd_desc* descs_dev, int n_descs, int max_O, int max_cols,
                                 const float* emb, float* demb, int B, int cemb, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && (emb || demb) && n_descs > 0 && max_O > 0 && max_cols > 0, "dd_emb_affine_bwd: bad arguments");
    if (emb) {          // stage 1: weight gradients + row scales of these descriptors
        DD_CHECK_CUDA(dd_launch_pdl(emb_affine_bwd_w_kernel, dim3(ceil_div(max_O, 8), n_descs), dim3(256), 0, stream, descs_dev, emb, B,
                                    cemb));
        DD_CHECK_LAUNCH();
    }
    if (demb) {         // stage 2: gradient of the embedding vector (needs stage 1 of every descriptor it is given)
        const int row_chunks = 4;
        DD_CHECK_CUDA(dd_launch_pdl(emb_affine_bwd_x_kernel, dim3(ceil_div(max_cols, 128), n_descs, row_chunks), dim3(128), 0, stream,
                                    descs_dev, demb, B, cemb, row_chunks));
        DD_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int dd_noise_embedding_bwd(const float* sigma, const float* freqs, const float* phases, int cnoise,
                                      const float* w_noise, int normalize, const float* label_emb, float label_balance,
                                      const float* demb, float* dweff, float* dlabel, int B, int cemb, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(sigma && freqs && phases && w_noise && label_emb && demb && dweff && dlabel,
               "dd_noise_embedding_bwd: null pointer");
    DD_REQUIRE(cnoise % 32 == 0, "dd_noise_embedding_bwd: cnoise=%d must be a multiple of 32", cnoise);
    const size_t smem = (size_t)B * cnoise * sizeof(float);
    DD_REQUIRE(smem <= 48 * 1024, "dd_noise_embedding_bwd: batch %d x cnoise %d exceeds shared memory", B, cnoise);
    DD_CHECK_CUDA(dd_launch_pdl(noise_embedding_bwd_kernel, dim3(ceil_div(cemb, 8)), dim3(256), smem, stream, sigma, freqs, phases, cnoise, w_noise, normalize,
                                                                        label_emb, label_balance, demb, dweff, dlabel, B,
                                                                        cemb));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_label_embedding_bwd(const float* emb_in, int Bc, int I, const float* mask, int Bm, const float* dout,
                                      float* dweff_label, float* dweff_uncond, int cemb, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(emb_in && mask && dout && dweff_label && dweff_uncond, "dd_label_embedding_bwd: null pointer");
    DD_REQUIRE(Bc == 1 || Bc == Bm, "dd_label_embedding_bwd: embedding batch %d does not broadcast to mask batch %d", Bc, Bm);
    const size_t smem = (size_t)Bc * I * sizeof(float);
    DD_REQUIRE(smem <= 48 * 1024, "dd_label_embedding_bwd: batch %d x dim %d exceeds shared memory", Bc, I);
    DD_CHECK_CUDA(dd_launch_pdl(label_embedding_bwd_kernel, dim3(ceil_div(cemb, 8)), dim3(256), smem, stream, emb_in, Bc, I, mask, Bm, dout, dweff_label,
                                                                        dweff_uncond, cemb));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_sigma_logvar_bwd(const float* sigma, int count, const float* freqs, const float* phases, int n,
                                   const float* dout, float* dw, int accumulate, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(sigma && freqs && phases && dout && dw, "dd_sigma_logvar_bwd: null pointer");
    if (n == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(logvar_bwd_kernel, dim3(ceil_div(n, 128)), dim3(128), 0, stream, sigma, freqs, phases, n, dout, count, dw, accumulate));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_head_grad(const float* dD, const float* sigma, float sigma_data, const float* x_ref, void* dF, int B,
                            int Cout, int H, int W, int Cpad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(dD && sigma && dF && Cpad >= Cout && Cpad % 8 == 0, "dd_head_grad: bad arguments");
    const long total = (long)B * H * W * Cpad;
    if (total == 0) return 0;
    DD_CHECK_CUDA(dd_launch_pdl(head_grad_kernel, dim3(grid_for_b(total, 256)), dim3(256), 0, stream, dD, sigma, sigma_data, x_ref,
                                                                 static_cast<__nv_bfloat16*>(dF), B, Cout, H, W, Cpad));
    DD_CHECK_LAUNCH();
    return 0;
}
