// Optimizer-side sweep of the train step: what the reference trainer does after backward() as
// 3-4 full passes over the 293 M parameters --
//   accelerator.clip_grad_norm_        training/trainer.py:1044   (norm of all gradients, then g *= coef)
//   optimizer.step()  (torch AdamW)    training/trainer.py:461-473,1062
//   ema_manager.update()               training/ema.py:284-313    (_foreach_lerp_ per EMA, optional feedback lerp)
//   module.normalize_weights()         training/trainer.py:1107-1108 -> modules/mp_tools.py:375-378
// -- as two launches: a deterministic two-stage gradient norm that leaves {norm, clip coefficient} on the device, and one
// batched launch (a warp per weight row, descriptor table) that reads p, g, m, v and the EMA copies once, applies
// clip * AdamW, the EMA / feedback lerps and the per-row re-normalisation, and writes each of them once.
// HBM-bound streaming work: 20 B read + 12 B written per parameter, + 8 B per fp32 EMA (16 B per fp64 EMA).
#include <stdlib.h>
#include "common.cuh"
#include "dualdiffusion_b200.h"
#include "optim_math.cuh"

namespace {

constexpr int kNormChunk = DD_GNORM_CHUNK;    // gradient elements per CTA of the norm kernel
constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum_o(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float block_sum_o(float v, float* red) {
    v = warp_sum_o(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum_o(t);
    __syncthreads();
    return t;
}

template <typename D>
__device__ __forceinline__ int find_desc(const D* __restrict__ descs, int n_descs, int unit) {
    int lo = 0, hi = n_descs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].chunk_begin <= unit) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// stage 1: partials[cta] = sum of squares of one 8192-element chunk of one gradient tensor
__global__ void __launch_bounds__(kThreads) grad_sqnorm_partial_kernel(const dd_gnorm_desc* __restrict__ descs,
                                                                       int n_descs, float* __restrict__ partials) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ float red[32];
    const int di = find_desc(descs, n_descs, (int)blockIdx.x);
    const dd_gnorm_desc d = descs[di];
    const long long begin = (long long)(blockIdx.x - d.chunk_begin) * kNormChunk;
    const long long rem = d.numel - begin;
    const int n = (int)(rem < (long long)kNormChunk ? rem : (long long)kNormChunk);
    const float* g = d.g + begin;
    float ss = 0.f;
    if (n > 0) {
        if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
            const float4* g4 = reinterpret_cast<const float4*>(g);
            const int n4 = n >> 2;
            for (int i = threadIdx.x; i < n4; i += kThreads) {
                const float4 x = __ldg(g4 + i);
                ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
            }
            for (int i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) ss += g[i] * g[i];
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) ss += g[i] * g[i];
        }
    }
    ss = block_sum_o(ss, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = ss;
}

// stage 2 (one CTA): fixed-order double accumulation of the partials -> out[0] = ||g||, out[1] = clip coefficient
// torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1 (NaN propagates, as torch.clamp does)
__global__ void __launch_bounds__(kThreads) grad_norm_finish_kernel(const float* __restrict__ partials, int n,
                                                                    float max_norm, float* __restrict__ out) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ double red[kThreads];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) s += (double)partials[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(red[0]);
        out[0] = norm;
        out[1] = clip_coef_from_norm(norm, max_norm);
    }
}

// One CTA per row of one parameter tensor viewed as [rows][row_len].  Pass 1 updates m, v, p and the EMA copies and
// accumulates the row's sum of squares; pass 2 (normalize != 0) rescales the row: every thread re-reads exactly the
// elements it wrote itself, so no fence is needed and the second read is served from L1/L2.
// SMEM_SEARCH (experiment, DD_OPTIM_SMEM_SEARCH=1): the CTA first copies the row_begin column of the descriptor table to
// shared memory with one round of parallel loads and searches there, instead of ~log2(n_descs) dependent global loads
// per CTA before any useful work (about 9 round trips for the UNet's ~290 tensors; a CTA only streams ~70 KB).
constexpr int kMaxSmemDescs = 2048;
template <bool SMEM_SEARCH>
__global__ void __launch_bounds__(kThreads) optim_step_batched_kernel(const dd_optim_desc* __restrict__ descs, int n_descs,
                                                                      OptimHyperDev h, const float* __restrict__ clip_coef) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ float red[32];
    int lo = 0, hi = n_descs - 1;
    const int unit = blockIdx.x;
    if constexpr (SMEM_SEARCH) {
        __shared__ int s_begin[kMaxSmemDescs];
        for (int i = threadIdx.x; i < n_descs; i += kThreads) s_begin[i] = descs[i].row_begin;
        __syncthreads();
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_begin[mid] <= unit) lo = mid; else hi = mid - 1;
        }
    } else {
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (descs[mid].row_begin <= unit) lo = mid; else hi = mid - 1;
        }
    }
#define DD_ROW_TID threadIdx.x
#define DD_ROW_NT kThreads
#define DD_ROW_SUM(x) block_sum_o(x, red)
#define DD_ROW_EXIT return
#include "optim_row_body.inl"
#undef DD_ROW_EXIT
}

// Persistent experiment (DD_OPTIM_PERSISTENT=1): a few CTAs per SM stride over the rows; the row_begin column is copied to
// shared memory once per CTA.  Same row body as above.
__global__ void __launch_bounds__(kThreads) optim_step_persistent_kernel(const dd_optim_desc* __restrict__ descs, int n_descs,
                                                                         OptimHyperDev h, const float* __restrict__ clip_coef,
                                                                         int total_rows) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    __shared__ float red[32];
    __shared__ int s_begin[kMaxSmemDescs];
    for (int i = threadIdx.x; i < n_descs; i += kThreads) s_begin[i] = descs[i].row_begin;
    __syncthreads();
    for (int unit = blockIdx.x; unit < total_rows; unit += gridDim.x) {
        int lo = 0, hi = n_descs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_begin[mid] <= unit) lo = mid; else hi = mid - 1;
        }
#define DD_ROW_EXIT continue
#include "optim_row_body.inl"
#undef DD_ROW_EXIT
#undef DD_ROW_TID
#undef DD_ROW_NT
#undef DD_ROW_SUM
    }
}

// Warp per row (the default since round 2: 4.09 -> 2.99 ms per sweep of the 293 M-parameter UNet, 0.58 -> 0.79 of the HBM
// roofline): a weight row is a few thousand elements, so a warp can own it -- no block-wide barriers, eight rows in flight
// per CTA, and kOptWarpRows consecutive rows share one descriptor search.  Same row body as the one-CTA-per-row kernels.
constexpr int kOptWarps = 8, kOptWarpRows = 2;
__global__ void __launch_bounds__(kOptWarps * 32) optim_step_warp_kernel(const dd_optim_desc* __restrict__ descs, int n_descs,
                                                                        OptimHyperDev h, const float* __restrict__ clip_coef,
                                                                        int total_rows) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int row0 = (blockIdx.x * kOptWarps + (threadIdx.x >> 5)) * kOptWarpRows;
    if (row0 >= total_rows) return;
    int lo = 0, hi = n_descs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].row_begin <= row0) lo = mid; else hi = mid - 1;
    }
    for (int rr = 0; rr < kOptWarpRows; ++rr) {
        const int unit = row0 + rr;
        if (unit >= total_rows) return;
        while (lo + 1 < n_descs && descs[lo + 1].row_begin <= unit) ++lo;
#define DD_ROW_TID (threadIdx.x & 31)
#define DD_ROW_NT 32
#define DD_ROW_SUM(x) warp_sum_o(x)
#define DD_ROW_EXIT continue
#include "optim_row_body.inl"
#undef DD_ROW_EXIT
#undef DD_ROW_TID
#undef DD_ROW_NT
#undef DD_ROW_SUM
    }
}

}  // namespace

extern "C" int dd_grad_norm_clip(const dd_gnorm_desc* descs_dev, int n_descs, int total_chunks, float* partials_dev,
                                 float max_norm, float* out_norm_coef_dev, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && partials_dev && out_norm_coef_dev && n_descs > 0 && total_chunks > 0,
               "dd_grad_norm_clip: bad arguments");
    DD_CHECK_CUDA(dd_launch_pdl(grad_sqnorm_partial_kernel, dim3(total_chunks), dim3(kThreads), 0, stream, descs_dev, n_descs, partials_dev));
    DD_CHECK_LAUNCH();
    DD_CHECK_CUDA(dd_launch_pdl(grad_norm_finish_kernel, dim3(1), dim3(kThreads), 0, stream, partials_dev, total_chunks, max_norm, out_norm_coef_dev));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_optim_step_batched(const dd_optim_desc* descs_dev, int n_descs, int total_rows,
                                     const dd_optim_hyper* hy, const float* norm_coef_dev, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(descs_dev && hy && n_descs > 0 && total_rows > 0, "dd_optim_step_batched: bad arguments");
    DD_REQUIRE(hy->n_ema >= 0 && hy->n_ema <= DD_OPTIM_MAX_EMA, "dd_optim_step_batched: n_ema=%d out of range", hy->n_ema);
    DD_REQUIRE(hy->bias_correction1 > 0.0 && hy->bias_correction2 > 0.0,
               "dd_optim_step_batched: bias corrections must be positive (step >= 1)");
    const OptimHyperDev h = make_hyper_dev(*hy);
    static const bool smem_search = getenv("DD_OPTIM_SMEM_SEARCH") != nullptr;      // tuning experiments, default off
    static const bool persistent = getenv("DD_OPTIM_PERSISTENT") != nullptr;
    static const bool warp_rows = getenv("DD_OPTIM_CTA_ROWS") == nullptr;      // default; DD_OPTIM_CTA_ROWS=1: one CTA per row
    if (warp_rows) {
        DD_CHECK_CUDA(dd_launch_pdl(optim_step_warp_kernel, dim3(ceil_div(total_rows, kOptWarps * kOptWarpRows)), dim3(kOptWarps * 32), 0, stream, descs_dev,
                                    n_descs, h, norm_coef_dev, total_rows));
    } else if (persistent && n_descs <= kMaxSmemDescs) {
        const int grid = total_rows < dd_num_sms() * 3 ? total_rows : dd_num_sms() * 3;
        DD_CHECK_CUDA(dd_launch_pdl(optim_step_persistent_kernel, dim3(grid), dim3(kThreads), 0, stream, descs_dev, n_descs, h, norm_coef_dev, total_rows));
    } else if (smem_search && n_descs <= kMaxSmemDescs)
        DD_CHECK_CUDA(dd_launch_pdl(optim_step_batched_kernel<true>, dim3(total_rows), dim3(kThreads), 0, stream, descs_dev, n_descs, h, norm_coef_dev));
    else
        DD_CHECK_CUDA(dd_launch_pdl(optim_step_batched_kernel<false>, dim3(total_rows), dim3(kThreads), 0, stream, descs_dev, n_descs, h, norm_coef_dev));
    DD_CHECK_LAUNCH();
    return 0;
}
