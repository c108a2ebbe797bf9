// MDCT side of the live audio format: framing / overlap-add / packing kernels around the
// MCLT (reference: /root/reference/src/utils/mclt.py:87-130, modules/formats/ms_mdct_dual.py:259-318).
//
// A 512-sample MCLT block is a 512 x 256 complex DFT-like matrix; with the window, the pre/post phase shifts, the
// 1/mel-density weighting and all scale factors folded in on the host (fp64) the transform of every frame of a batch is one
// plain fp32 library GEMM ([frames x 512] x [512 x 512]), 1.4 GFLOP per 45 s stereo item.  The kernels here do what
// surrounds it: reflect-padded frame extraction, |.| of the complex coefficients, the 50 %-overlap add of the inverse,
// and the linearisation of the mel spectrogram in front of the min-norm inverse mel filterbank.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>

namespace {

inline int grid_for_m(long total, int block, int cap_mult = 16) {
    const long blocks = (total + block - 1) / block;
    return (int)std::max<long>(1, std::min<long>(blocks, (long)dd_num_sms() * cap_mult));
}

// out[s][t][n] = x_reflect[s][t*hop + n - pad_left]   (torch.nn.functional.pad(mode="reflect") + unfold, mclt.py:90-95)
__global__ void frame_reflect_kernel(const float* __restrict__ raw, float* __restrict__ out, int S, long L, int T, int bw,
                                     int hop, int pad_left) {
    const long total = (long)S * T * bw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int n = (int)(idx % bw);
        long r = idx / bw;
        const int t = (int)(r % T), s = (int)(r / T);
        long j = (long)t * hop + n - pad_left;
        if (j < 0) j = -j;
        if (j >= L) j = 2 * (L - 1) - j;
        out[idx] = raw[(size_t)s * L + j];
    }
}

// |re + i im| * scale for coefficient rows stored as [S][2N][T] (rows 0..N-1 real, N..2N-1 imaginary) -> [S][N][T]
__global__ void complex_abs_kernel(const float* __restrict__ y, float* __restrict__ out, int S, int N, int T, float scale) {
    const long total = (long)S * N * T;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long rt = idx % ((long)N * T);
        const long s = idx / ((long)N * T);
        const float re = y[s * 2 * N * T + rt], im = y[s * 2 * N * T + (long)N * T + rt];
        out[idx] = sqrtf(re * re + im * im) * scale;
    }
}

// imclt overlap-add (mclt.py:123-130): frames y[s][t][2N] at hop N, first and last half block cropped:
//   out[s][m] = y[s][m/N + 1][m % N] + y[s][m/N][m % N + N],  m in [0, (T-1) N)
__global__ void mdct_ola_kernel(const float* __restrict__ y, float* __restrict__ out, int S, int T, int N) {
    const long Lout = (long)(T - 1) * N;
    const long total = (long)S * Lout;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long m = idx % Lout, s = idx / Lout;
        const long t0 = m / N;
        const int n = (int)(m - t0 * N);
        const float* ys = y + (size_t)s * T * 2 * N;
        out[idx] = ys[(t0 + 1) * 2 * N + n] + ys[t0 * 2 * N + n + N];
    }
}

// (mel - offset).clip(min=0) ** inv_exponent   (ms_mdct_dual.py:261-265)
__global__ void mel_linearize_kernel(const float* __restrict__ mel, float* __restrict__ out, long n, float offset,
                                     float inv_exponent) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        const float v = fmaxf(mel[idx] - offset, 0.f);
        out[idx] = inv_exponent == 1.f ? v : powf(v, inv_exponent);
    }
}

// ms_mdct_dual_2.py:203-216: out = ((sum_i mel_i[s][f][t] * ww[f][i]) ** exponent + offset) * inv_scale over the n_win
// per-window mel spectrograms mels [n_win][S][F][T] (each already filtered: the blend is per mel filter, after the bank).
__global__ void mel_blend_kernel(const float* __restrict__ mels, const float* __restrict__ ww, int n_win, long S, int F, int T,
                                 float exponent, float offset, float inv_scale, float* __restrict__ out) {
    const long total = S * F * T;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f = (int)((idx / T) % F);
        float acc = 0.f;
        for (int i = 0; i < n_win; ++i) acc += mels[(long)i * total + idx] * __ldg(ww + f * n_win + i);
        out[idx] = (powf(acc, exponent) + offset) * inv_scale;
    }
}

// ms_mdct_dual_2.py:275-289 on coefficient rows y [S][2N][T] (rows 0..N-1 real, N..2N-1 imaginary, unscaled MCLT):
//   |z| -> phase = clip(re / max(|z|, 1e-20), -1, 1) * phase_mul;  psd = ((|z| * inv_density[k]) ** exponent + offset) * inv_scale
__global__ void mdct_phase_psd_kernel(const float* __restrict__ y, const float* __restrict__ inv_density, int S, int N, int T,
                                      float exponent, float offset, float inv_scale, float phase_mul, float* __restrict__ phase,
                                      float* __restrict__ psd) {
    const long total = (long)S * N * T;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long rt = idx % ((long)N * T);
        const long s = idx / ((long)N * T);
        const int k = (int)(rt / T);
        const float re = y[s * 2 * N * T + rt], im = y[s * 2 * N * T + (long)N * T + rt];
        const float mag = sqrtf(re * re + im * im);
        phase[idx] = fminf(fmaxf(re / fmaxf(mag, 1e-20f), -1.f), 1.f) * phase_mul;
        psd[idx] = (powf(mag * __ldg(inv_density + k), exponent) + offset) * inv_scale;
    }
}

}  // namespace

extern "C" int dd_frame_reflect(const float* raw, float* out, int S, long L, int T, int block_width, int hop, int pad_left,
                                void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(raw && out && S > 0 && L > 1 && T > 0 && block_width > 0 && hop > 0, "dd_frame_reflect: bad arguments");
    DD_REQUIRE(pad_left < L && (long)(T - 1) * hop + block_width - pad_left - L < L,
               "dd_frame_reflect: reflection padding must be shorter than the signal");
    const long total = (long)S * T * block_width;
    frame_reflect_kernel<<<grid_for_m(total, 256), 256, 0, stream>>>(raw, out, S, L, T, block_width, hop, pad_left);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_complex_abs(const float* y, float* out, int S, int N, int T, float scale, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(y && out && S > 0 && N > 0 && T > 0, "dd_complex_abs: bad arguments");
    complex_abs_kernel<<<grid_for_m((long)S * N * T, 256), 256, 0, stream>>>(y, out, S, N, T, scale);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_mdct_ola(const float* y, float* out, int S, int T, int N, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(y && out && S > 0 && T > 1 && N > 0, "dd_mdct_ola: bad arguments");
    mdct_ola_kernel<<<grid_for_m((long)S * (T - 1) * N, 256), 256, 0, stream>>>(y, out, S, T, N);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_mel_linearize(const float* mel, float* out, long n, float offset, float inv_exponent, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(mel && out && n >= 0, "dd_mel_linearize: bad arguments");
    if (n == 0) return 0;
    mel_linearize_kernel<<<grid_for_m(n, 256), 256, 0, stream>>>(mel, out, n, offset, inv_exponent);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_mel_blend(const float* mels, const float* ww, int n_win, int S, int F, int T, float exponent, float offset,
                            float inv_scale, float* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(mels && ww && out && n_win >= 1 && S > 0 && F > 0 && T > 0, "dd_mel_blend: bad arguments");
    mel_blend_kernel<<<grid_for_m((long)S * F * T, 256), 256, 0, stream>>>(mels, ww, n_win, S, F, T, exponent, offset, inv_scale, out);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_mdct_phase_psd(const float* y, const float* inv_density, int S, int N, int T, float exponent, float offset,
                                 float inv_scale, float phase_mul, float* phase, float* psd, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(y && inv_density && phase && psd && S > 0 && N > 0 && T > 0, "dd_mdct_phase_psd: bad arguments");
    mdct_phase_psd_kernel<<<grid_for_m((long)S * N * T, 256), 256, 0, stream>>>(y, inv_density, S, N, T, exponent, offset, inv_scale,
                                                                             phase_mul, phase, psd);
    DD_CHECK_LAUNCH();
    return 0;
}
