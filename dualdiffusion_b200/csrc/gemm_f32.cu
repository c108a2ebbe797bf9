// fp32 CUDA-core GEMM for the audio-format transforms: the MCLT / inverse MCLT of
// MS_MDCT_DualFormat (reference /root/reference/src/utils/mclt.py:87-130, modules/formats/ms_mdct_dual.py:259-318) and
// the min-norm inverse mel filterbanks (frequency_scale.py:130-142, ms_mdct_dual.py:259-270).  The reference's formats
// are fp32-only (format.py:40) and these transforms are tiny next to the UNet (1.4 GFLOP per 45 s stereo item), so they
// run as one strided, batched SIMT GEMM with the surrounding data movement folded in instead of a library call:
//   * operands are addressed through (row, col, batch) element strides, so transposed views cost nothing;
//   * B can be the reflect-padded FRAMES of a raw signal (B[k][n] = raw[reflect(n * hop + k - pad)]): the MCLT needs no
//     materialised frame tensor;
//   * A can be linearised on load ((a * a_scale + a_offset).clip(0) ** a_pow: the mel "unscale") and the result clamped
//     at zero on store (the relu of the inverse mel).
// 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread, fp32 FMA in k order.
#include "common.cuh"
#include "dualdiffusion_b200.h"

namespace {

constexpr int kBM = 64, kBN = 64, kBK = 16, kThreadsG = 256;

struct GemmParams {
    const float* a; long a_m, a_k, a_b;
    const float* b; long b_k, b_n, b_b;
    float* c; long c_m, c_n, c_b;
    int M, N, K;
    int gather_hop, gather_pad; long gather_len;      // gather_hop > 0: B[k][n] = raw[reflect(n * hop + k - pad)]
    float a_scale, a_offset, a_pow;                    // a_pow > 0: A element = clip(a * a_scale + a_offset, 0) ** a_pow
    int relu;
};

__global__ void __launch_bounds__(kThreadsG) gemm_f32_kernel(const GemmParams p) {
    __shared__ float As[kBK][kBM + 4];
    __shared__ float Bs[kBK][kBN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
    const float* A = p.a + (size_t)blockIdx.z * p.a_b;
    const float* B = p.b + (size_t)blockIdx.z * p.b_b;
    float* C = p.c + (size_t)blockIdx.z * p.c_b;
    const bool a_m_fast = p.a_m == 1;                  // which index runs fastest over the threads of a load
    const bool b_n_fast = p.gather_hop == 0 && p.b_n == 1;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += kBK) {
#pragma unroll
        for (int i = 0; i < (kBM * kBK) / kThreadsG; ++i) {
            const int e = tid + i * kThreadsG;
            const int m = a_m_fast ? (e & (kBM - 1)) : (e >> 4), k = a_m_fast ? (e >> 6) : (e & (kBK - 1));
            float v = 0.f;
            if (m0 + m < p.M && k0 + k < p.K) {
                v = __ldg(A + (size_t)(m0 + m) * p.a_m + (size_t)(k0 + k) * p.a_k);
                if (p.a_pow > 0.f) {
                    v = fmaxf(fmaf(v, p.a_scale, p.a_offset), 0.f);
                    if (p.a_pow != 1.f) v = powf(v, p.a_pow);
                }
            }
            As[k][m] = v;
        }
#pragma unroll
        for (int i = 0; i < (kBN * kBK) / kThreadsG; ++i) {
            const int e = tid + i * kThreadsG;
            const int n = b_n_fast ? (e & (kBN - 1)) : (e >> 4), k = b_n_fast ? (e >> 6) : (e & (kBK - 1));
            float v = 0.f;
            if (n0 + n < p.N && k0 + k < p.K) {
                if (p.gather_hop > 0) {
                    long j = (long)(n0 + n) * p.gather_hop + (k0 + k) - p.gather_pad;
                    if (j < 0) j = -j;
                    if (j >= p.gather_len) j = 2 * (p.gather_len - 1) - j;
                    v = __ldg(B + j);
                } else {
                    v = __ldg(B + (size_t)(k0 + k) * p.b_k + (size_t)(n0 + n) * p.b_n);
                }
            }
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            const float v = p.relu ? fmaxf(acc[i][j], 0.f) : acc[i][j];
            C[(size_t)m * p.c_m + (size_t)n * p.c_n] = v;
        }
    }
}

}  // namespace

extern "C" int dd_gemm_f32(const float* a, long a_m, long a_k, long a_batch, const float* b, long b_k, long b_n, long b_batch,
                           float* c, long c_m, long c_n, long c_batch, int M, int N, int K, int batch, int gather_hop,
                           int gather_pad, long gather_len, float a_scale, float a_offset, float a_pow, int relu,
                           void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(a && b && c, "dd_gemm_f32: null pointer");
    DD_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "dd_gemm_f32: bad sizes (M %d N %d K %d batch %d)", M, N,
               K, batch);
    DD_REQUIRE(gather_hop == 0 || (gather_len > 1 && gather_pad < gather_len &&
                                   (long)(N - 1) * gather_hop + K - 1 - gather_pad - gather_len < gather_len - 1),
               "dd_gemm_f32: reflection padding must be shorter than the signal");
    GemmParams p{a, a_m, a_k, a_batch, b, b_k, b_n, b_batch, c, c_m, c_n, c_batch, M, N, K, gather_hop, gather_pad, gather_len,
                 a_scale, a_offset, a_pow, relu};
    const dim3 grid(ceil_div(N, kBN), ceil_div(M, kBM), batch);
    gemm_f32_kernel<<<grid, kThreadsG, 0, stream>>>(p);
    DD_CHECK_LAUNCH();
    return 0;
}
