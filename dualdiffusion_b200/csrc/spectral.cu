// mel-STFT encode and FGLA (fast Griffin-Lim) phase-reconstruction decode of the reference's spectrogram format
// (src/modules/formats/old/spectrogram.py:176-238, frequency_scale.py:127-128, old/phase_recovery.py:40-129),
// replacing torchaudio.Spectrogram / torch.stft / torch.istft (cuFFT) + dense mel matmul + eager elementwise.
//
// Every frame's length-n_fft real FFT is computed inside one CTA as a half-length complex Stockham FFT in shared
// memory (mixed radix 8/5/4/2, so the reference's n_fft = 6400 = 2^8*5^2 and the live format's 4096 both work),
// with windowing, |.|, the triangular-sparse mel filterbank, overlap-add and the Griffin-Lim momentum update
// fused around it.  HBM traffic per FGLA iteration is the algorithmic minimum of SURVEY.md §8(d): phase A reads the
// state T and the magnitudes once, phase B reads T and writes T' once; the waveform round-trips through L2.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>

namespace {

constexpr int kFftThreads = 320;
constexpr int kMaxStagesFft = 8;
constexpr int kEncFrames = 32;      // frames per CTA in the encoder (one 128 B output row segment per filter)
constexpr int kOlaFrames = 32;      // frames per CTA in the inverse STFT (overlap-add accumulated in shared memory)

struct FftPlan {
    int n;                          // complex length = n_fft / 2
    int n_stages;
    int radix[kMaxStagesFft];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = mul_neg_i(csub(b, d));
    a = cadd(t0, t2); b = cadd(t1, t3); c = csub(t0, t2); d = csub(t1, t3);
}

template <int R>
__device__ __forceinline__ void butterfly(float2 (&v)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void butterfly<4>(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
template <>
__device__ __forceinline__ void butterfly<8>(float2 (&v)[8]) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752f;
    o1 = cmul(o1, make_float2(h, -h));          // w8^1
    o2 = mul_neg_i(o2);                         // w8^2
    o3 = cmul(o3, make_float2(-h, -h));         // w8^3
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
template <>
__device__ __forceinline__ void butterfly<5>(float2 (&v)[5]) {
    // 5-point DFT, w = exp(-2 pi i / 5)
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;     // cos(2pi/5), cos(4pi/5)
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;      // sin(2pi/5), sin(4pi/5)
    const float2 a = v[0];
    const float2 p14 = cadd(v[1], v[4]), m14 = csub(v[1], v[4]);
    const float2 p23 = cadd(v[2], v[3]), m23 = csub(v[2], v[3]);
    v[0] = make_float2(a.x + p14.x + p23.x, a.y + p14.y + p23.y);
    const float2 r1 = make_float2(a.x + c1 * p14.x + c2 * p23.x, a.y + c1 * p14.y + c2 * p23.y);
    const float2 r2 = make_float2(a.x + c2 * p14.x + c1 * p23.x, a.y + c2 * p14.y + c1 * p23.y);
    // -i * (s1*m14 + s2*m23) and -i * (s2*m14 - s1*m23)
    const float2 q1 = make_float2(s1 * m14.y + s2 * m23.y, -(s1 * m14.x + s2 * m23.x));
    const float2 q2 = make_float2(s2 * m14.y - s1 * m23.y, -(s2 * m14.x - s1 * m23.x));
    v[1] = cadd(r1, q1); v[4] = csub(r1, q1);
    v[2] = cadd(r2, q2); v[3] = csub(r2, q2);
}

// One Stockham autosort pass of radix R (decimation in time): in -> out.
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out, int n, int ns,
                                              const float2* __restrict__ tw) {
    const int nb = n / R;
    const int tw_step = n / (ns * R);
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int k = j % ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[j + r * nb];
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(tw + r * k * tw_step));
        butterfly<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[j0 + r * ns] = v[r];
    }
}

// Forward complex FFT of length plan.n on shared memory.  Returns the buffer holding the result.
__device__ __forceinline__ float2* fft_forward(const FftPlan& plan, float2* a, float2* b, const float2* tw) {
    int ns = 1;
    float2 *src = a, *dst = b;
    for (int s = 0; s < plan.n_stages; ++s) {
        const int r = plan.radix[s];
        if (r == 8) stockham_pass<8>(src, dst, plan.n, ns, tw);
        else if (r == 5) stockham_pass<5>(src, dst, plan.n, ns, tw);
        else if (r == 4) stockham_pass<4>(src, dst, plan.n, ns, tw);
        else stockham_pass<2>(src, dst, plan.n, ns, tw);
        ns *= r;
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    return src;
}

__device__ __forceinline__ int reflect_index(int j, int len) {
    if (j < 0) j = -j;
    if (j >= len) j = 2 * (len - 1) - j;
    return j;
}

// Windowed frame `t` of a (reflect-padded, centred) signal packed as n_fft/2 complex values, then the real-input
// FFT.  src(j) returns sample j of the un-padded signal.  On return spec[0..n] (n+1 bins) holds the one-sided
// spectrum; `other` is the second scratch buffer.  Both buffers hold n+1 float2.
template <typename SrcFn>
__device__ __forceinline__ float2* stft_frame(const FftPlan& plan, int t, int hop, int len, SrcFn src,
                                              const float* __restrict__ window, const float2* tw, const float2* tw_half,
                                              float2* a, float2* b) {
    const int n = plan.n;
    const int p0 = t * hop - n;                       // first padded-domain sample of the frame, relative to signal
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        const float x0 = src(reflect_index(p0 + 2 * m, len)) * __ldg(window + 2 * m);
        const float x1 = src(reflect_index(p0 + 2 * m + 1, len)) * __ldg(window + 2 * m + 1);
        a[m] = make_float2(x0, x1);
    }
    __syncthreads();
    float2* z = fft_forward(plan, a, b, tw);
    float2* o = (z == a) ? b : a;
    // X[k] = (Z[k] + conj(Z[n-k]))/2 - (i/2) e^{-2 pi i k/(2n)} (Z[k] - conj(Z[n-k])),  k = 0..n  (Z[n] == Z[0])
    for (int k = threadIdx.x; k <= n; k += blockDim.x) {
        const float2 zk = z[k == n ? 0 : k];
        const float2 zc = z[k == 0 ? 0 : n - k];
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
        const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));
        const float2 od = cmul(mul_neg_i(d), __ldg(tw_half + k));
        o[k] = cadd(e, od);
    }
    __syncthreads();
    return o;
}

// ------------------------------------------------------------------------------------------
// mel-STFT encoder
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads)
stft_mel_kernel(const float* __restrict__ raw, int len, const float* __restrict__ window, const float2* __restrict__ tw,
                const float2* __restrict__ tw_half, const __grid_constant__ FftPlan plan, int hop, int n_frames,
                const int* __restrict__ fb_start, const int* __restrict__ fb_count, const int* __restrict__ fb_offset,
                const float* __restrict__ fb_weight, int n_filters, float exponent, float mean, float scale,
                float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    const int n = plan.n;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + (n + 1);
    float* tile = reinterpret_cast<float*>(b + (n + 1));          // [n_filters][kEncFrames + 1]
    const int s = blockIdx.y;
    const int t0 = blockIdx.x * kEncFrames;
    const float* sig = raw + (size_t)s * len;
    for (int f = 0; f < kEncFrames; ++f) {
        const int t = t0 + f;
        if (t >= n_frames) break;
        float2* spec = stft_frame(plan, t, hop, len, [&](int j) { return __ldg(sig + j); }, window, tw, tw_half, a, b);
        float* mag = reinterpret_cast<float*>(spec == a ? b : a);
        for (int k = threadIdx.x; k <= n; k += blockDim.x) mag[k] = sqrtf(spec[k].x * spec[k].x + spec[k].y * spec[k].y);
        __syncthreads();
        for (int m = threadIdx.x; m < n_filters; m += blockDim.x) {
            const int st = fb_start[m], cnt = fb_count[m], off = fb_offset[m];
            float acc = 0.f;
            for (int j = 0; j < cnt; ++j) acc += mag[st + j] * __ldg(fb_weight + off + j);
            const float v = (exponent == 0.25f) ? sqrtf(sqrtf(acc)) : powf(acc, exponent);
            tile[m * (kEncFrames + 1) + f] = (v - mean) * scale;
        }
        __syncthreads();
    }
    const int nf = min(kEncFrames, n_frames - t0);
    for (int i = threadIdx.x; i < n_filters * kEncFrames; i += blockDim.x) {
        const int m = i / kEncFrames, f = i % kEncFrames;
        if (f < nf) out[((size_t)s * n_filters + m) * n_frames + t0 + f] = tile[m * (kEncFrames + 1) + f];
    }
}

// ------------------------------------------------------------------------------------------
// FGLA phase A: A = T/(|T|+1e-16); X = A * M_k; inverse STFT frame, window, overlap-add
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads)
fgla_istft_kernel(const float2* __restrict__ state, const float* __restrict__ mag, int stereo, float interp_t,
                  const float* __restrict__ window, const float2* __restrict__ tw, const float2* __restrict__ tw_half,
                  const __grid_constant__ FftPlan plan, int hop, int n_frames, float* __restrict__ ola, int ola_len) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    const int n = plan.n, bins = n + 1;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + bins;
    float* acc = reinterpret_cast<float*>(b + bins);              // [(kOlaFrames-1)*hop + 2n]
    const int span = (kOlaFrames - 1) * hop + 2 * n;
    const int s = blockIdx.y;
    const int t0 = blockIdx.x * kOlaFrames;
    for (int i = threadIdx.x; i < span; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const float inv_n = 1.f / (float)n;
    for (int f = 0; f < kOlaFrames; ++f) {
        const int t = t0 + f;
        if (t >= n_frames) break;
        const size_t row = ((size_t)s * n_frames + t) * bins;
        const size_t row_other = ((size_t)(s ^ 1) * n_frames + t) * bins;
        // X[k] = angle * magnitude  (phase_recovery.py:84-95); stored in b
        for (int k = threadIdx.x; k < bins; k += blockDim.x) {
            float m = __ldg(mag + row + k);
            if (stereo) {
                const float merged = 0.5f * (m + __ldg(mag + row_other + k));          // :63-64 (L+R)/2
                m = interp_t > 0.f ? merged + interp_t * (m - merged) : merged;        // :86-88 lerp(merged, spec, t)
            }
            float2 ang = make_float2(1.f, 0.f);
            if (state) {
                const float2 tv = state[row + k];
                const float inv = 1.f / (sqrtf(tv.x * tv.x + tv.y * tv.y) + 1e-16f);    // :115
                ang = make_float2(tv.x * inv, tv.y * inv);
            }
            float2 x = make_float2(ang.x * m, ang.y * m);
            if (k == 0 || k == n) x.y = 0.f;            // C2R transforms ignore the imaginary part of DC / Nyquist
            b[k] = x;
        }
        __syncthreads();
        // Z[k] = E[k] + i O[k], E = (X[k] + conj(X[n-k]))/2, O = e^{+2 pi i k/(2n)} (X[k] - conj(X[n-k]))/2;
        // inverse FFT through the forward one: IFFT(Z) = conj(FFT(conj(Z))) / n  ->  a holds conj(Z)
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const float2 xk = b[k], xc = b[n - k];
            const float2 e = make_float2(0.5f * (xk.x + xc.x), 0.5f * (xk.y - xc.y));
            const float2 d = make_float2(0.5f * (xk.x - xc.x), 0.5f * (xk.y + xc.y));
            const float2 w = __ldg(tw_half + k);
            const float2 o = cmul(d, make_float2(w.x, -w.y));
            a[k] = make_float2(e.x - o.y, -(e.y + o.x));                                // conj(E + iO)
        }
        __syncthreads();
        const float2* z = fft_forward(plan, a, b, tw);
        float* dst = acc + f * hop;
        for (int m = threadIdx.x; m < n; m += blockDim.x) {
            const float2 v = z[m];
            dst[2 * m] += v.x * inv_n * __ldg(window + 2 * m);
            dst[2 * m + 1] += -v.y * inv_n * __ldg(window + 2 * m + 1);
        }
        __syncthreads();
    }
    // every output sample is touched by at most two CTAs of the same signal, so the float atomics are
    // order-independent (a+b == b+a) and the result is deterministic
    float* g = ola + (size_t)s * ola_len + (size_t)t0 * hop;
    const int valid = min(span, ola_len - t0 * hop);
    for (int i = threadIdx.x; i < valid; i += blockDim.x) atomicAdd(g + i, acc[i]);
}

// ------------------------------------------------------------------------------------------
// FGLA phase B: rebuilt = STFT(ISTFT(..)); T <- rebuilt - momentum * T   (phase_recovery.py:97-117)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads)
fgla_stft_update_kernel(const float* __restrict__ ola, const float* __restrict__ env, int ola_len, int len,
                        const float* __restrict__ window, const float2* __restrict__ tw,
                        const float2* __restrict__ tw_half, const __grid_constant__ FftPlan plan, int hop, int n_frames,
                        float2* __restrict__ state, float momentum, int first) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    const int n = plan.n, bins = n + 1;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + bins;
    const int s = blockIdx.y;
    const float* o = ola + (size_t)s * ola_len + n;       // trim n_fft/2 (center=True)
    const float* e = env + n;
    for (int f = 0; f < kEncFrames; ++f) {
        const int t = blockIdx.x * kEncFrames + f;
        if (t >= n_frames) break;
        float2* spec = stft_frame(plan, t, hop, len, [&](int j) { return __ldg(o + j) / __ldg(e + j); }, window, tw,
                                  tw_half, a, b);
        float2* row = state + ((size_t)s * n_frames + t) * bins;
        for (int k = threadIdx.x; k < bins; k += blockDim.x) {
            float2 r = spec[k];
            if (!first) {
                const float2 p = row[k];
                r.x -= momentum * p.x; r.y -= momentum * p.y;
            }
            row[k] = r;
        }
        __syncthreads();
    }
}

__global__ void ola_finalize_kernel(const float* __restrict__ ola, const float* __restrict__ env, int ola_len, int half,
                                    int len, float* __restrict__ out, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long s = i / len;
        const int j = (int)(i - s * len);
        out[i] = ola[s * ola_len + half + j] / env[half + j];
    }
}

int make_plan(int n_fft, FftPlan& plan) {
    if (n_fft < 16 || n_fft % 2) return 1;
    int n = n_fft / 2;
    plan.n = n;
    plan.n_stages = 0;
    const int radices[4] = {8, 5, 4, 2};
    for (int ri = 0; ri < 4; ++ri)
        while (n % radices[ri] == 0 && n > 1) {
            if (plan.n_stages == kMaxStagesFft) return 1;
            plan.radix[plan.n_stages++] = radices[ri];
            n /= radices[ri];
        }
    return n == 1 ? 0 : 1;
}

size_t fft_smem_bytes(const FftPlan& plan) { return (size_t)2 * (plan.n + 1) * sizeof(float2); }

}  // namespace

extern "C" int dd_stft_mel(const float* raw, int n_signals, int len, const float* window, const float* twiddles,
                           const float* twiddles_half, int n_fft, int hop, const int* fb_start, const int* fb_count,
                           const int* fb_offset, const float* fb_weight, int n_filters, float exponent, float mean,
                           float scale, float* out, int n_frames, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(raw && window && twiddles && twiddles_half && fb_start && fb_count && fb_offset && fb_weight && out,
               "dd_stft_mel: null pointer");
    FftPlan plan;
    DD_REQUIRE(make_plan(n_fft, plan) == 0, "dd_stft_mel: n_fft=%d is not of the form 2*2^a*5^b", n_fft);
    DD_REQUIRE(len > n_fft / 2, "dd_stft_mel: signal shorter than the reflect padding");
    DD_REQUIRE(n_frames == 1 + len / hop, "dd_stft_mel: n_frames must be 1 + len/hop (center=True)");
    if (n_signals == 0) return 0;
    const size_t smem = fft_smem_bytes(plan) + (size_t)n_filters * (kEncFrames + 1) * sizeof(float);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const dim3 grid(ceil_div(n_frames, kEncFrames), n_signals);
    stft_mel_kernel<<<grid, kFftThreads, smem, stream>>>(raw, len, window, reinterpret_cast<const float2*>(twiddles),
                                                         reinterpret_cast<const float2*>(twiddles_half), plan, hop,
                                                         n_frames, fb_start, fb_count, fb_offset, fb_weight, n_filters,
                                                         exponent, mean, scale, out);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_fgla_istft(const float* state, const float* mag_tk, int n_signals, int n_frames, int stereo,
                             float interp_t, const float* window, const float* twiddles, const float* twiddles_half,
                             int n_fft, int hop, float* ola, int ola_len, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(mag_tk && window && twiddles && twiddles_half && ola, "dd_fgla_istft: null pointer");
    DD_REQUIRE(!stereo || n_signals % 2 == 0, "dd_fgla_istft: stereo needs an even number of signals");
    FftPlan plan;
    DD_REQUIRE(make_plan(n_fft, plan) == 0, "dd_fgla_istft: n_fft=%d is not of the form 2*2^a*5^b", n_fft);
    DD_REQUIRE(ola_len == n_fft + hop * (n_frames - 1), "dd_fgla_istft: ola_len must be n_fft + hop*(n_frames-1)");
    DD_REQUIRE(n_fft <= kOlaFrames * hop, "dd_fgla_istft: n_fft/hop=%d overlaps more than %d frames", n_fft / hop,
               kOlaFrames);
    if (n_signals == 0) return 0;
    DD_CHECK_CUDA(cudaMemsetAsync(ola, 0, (size_t)n_signals * ola_len * sizeof(float), stream));
    const size_t smem = fft_smem_bytes(plan) + (size_t)((kOlaFrames - 1) * hop + n_fft) * sizeof(float);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(fgla_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const dim3 grid(ceil_div(n_frames, kOlaFrames), n_signals);
    fgla_istft_kernel<<<grid, kFftThreads, smem, stream>>>(reinterpret_cast<const float2*>(state), mag_tk, stereo,
                                                           interp_t, window, reinterpret_cast<const float2*>(twiddles),
                                                           reinterpret_cast<const float2*>(twiddles_half), plan, hop,
                                                           n_frames, ola, ola_len);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_fgla_stft_update(const float* ola, const float* env, int n_signals, int n_frames, int len,
                                   const float* window, const float* twiddles, const float* twiddles_half, int n_fft,
                                   int hop, float* state, float momentum, int first, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(ola && env && window && twiddles && twiddles_half && state, "dd_fgla_stft_update: null pointer");
    FftPlan plan;
    DD_REQUIRE(make_plan(n_fft, plan) == 0, "dd_fgla_stft_update: n_fft=%d is not of the form 2*2^a*5^b", n_fft);
    DD_REQUIRE(len == hop * (n_frames - 1), "dd_fgla_stft_update: len must be hop*(n_frames-1)");
    if (n_signals == 0) return 0;
    const size_t smem = fft_smem_bytes(plan);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(fgla_stft_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        smem_set = smem;
    }
    const dim3 grid(ceil_div(n_frames, kEncFrames), n_signals);
    fgla_stft_update_kernel<<<grid, kFftThreads, smem, stream>>>(
        ola, env, n_fft + hop * (n_frames - 1), len, window, reinterpret_cast<const float2*>(twiddles),
        reinterpret_cast<const float2*>(twiddles_half), plan, hop, n_frames, reinterpret_cast<float2*>(state), momentum,
        first);
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_ola_finalize(const float* ola, const float* env, int n_signals, int ola_len, int n_fft, int len,
                               float* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(ola && env && out, "dd_ola_finalize: null pointer");
    const long total = (long)n_signals * len;
    if (total == 0) return 0;
    const int blocks = (int)std::min<long>((total + 255) / 256, (long)dd_num_sms() * 16);
    ola_finalize_kernel<<<blocks, 256, 0, stream>>>(ola, env, ola_len, n_fft / 2, len, out, total);
    DD_CHECK_LAUNCH();
    return 0;
}

