// mel-STFT encode and FGLA (fast Griffin-Lim) phase-reconstruction decode of the reference's spectrogram format
// (src/modules/formats/old/spectrogram.py:176-238, frequency_scale.py:127-128, old/phase_recovery.py:40-129),
// replacing torchaudio.Spectrogram / torch.stft / torch.istft (cuFFT) + dense mel matmul + eager elementwise.
//
// Every frame's length-n_fft real FFT is computed inside one CTA as a half-length complex Stockham FFT in shared
// memory (compile-time mixed-radix plans for the reference's n_fft = 6400 = 2^8*5^2 and the live format's 4096),
// with windowing, |.|, the triangular-sparse mel filterbank, overlap-add and the Griffin-Lim momentum update
// fused around it.  HBM traffic per FGLA iteration is the algorithmic minimum of SURVEY.md §8(d): phase A reads the
// state T and the magnitudes once, phase B reads T and writes T' once; the waveform round-trips through L2.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>

namespace {

constexpr int kEncFrames = 32;      // frames per CTA in the encoder (one 128 B output row segment per filter)
constexpr int kOlaFrames = 32;      // frames per CTA in the inverse STFT (overlap-add accumulated in shared memory)

// Shared-memory FFT buffers are padded by one element every 8 (index i -> i + i/8): the first Stockham passes
// scatter with a stride of `radix` elements, which would otherwise be a 16-way bank conflict on 8-byte elements.
__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 3); }

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

// Compile-time FFT plans (complex length N = n_fft/2, T threads per frame).  Everything that depends on the plan
// -- radices, strides, twiddle steps, loop trip counts -- folds into immediates: the first, runtime-planned version
// of these kernels spent ~230 instructions per complex element on index arithmetic.
template <int N> struct Plan;
template <> struct Plan<3200> {            // n_fft 6400 = 2^8 * 5^2 (reference SpectrogramFormat)
    static constexpr int kThreads = 320, kStages = 5;
    static constexpr int radix(int s) { return s == 0 ? 8 : s == 1 ? 8 : s == 2 ? 5 : s == 3 ? 5 : 2; }
};
template <> struct Plan<2048> {            // n_fft 4096 (live MS_MDCT_DualFormat windows)
    static constexpr int kThreads = 256, kStages = 4;
    static constexpr int radix(int s) { return s == 0 ? 8 : s == 1 ? 8 : s == 2 ? 8 : 4; }
};

// One Stockham autosort pass of radix R with NS = product of the previous radices: in -> out (both padded).
// `tws` is the shared-memory half table exp(-2 pi i m / N), m < N/2 (the other half is its negation): with most
// of the SM's SRAM carved out as shared memory there is next to no L1 left, and __ldg twiddles came from L2.
template <int N, int T, int R, int NS>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out,
                                              const float2* __restrict__ tws) {
    constexpr int NB = N / R, HALF = N / 2, STEP = N / (NS * R);
    constexpr int ITERS = (NB + T - 1) / T;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int j = (int)threadIdx.x + it * T;
        if (NB % T != 0 && j >= NB) break;
        const int k = (NS & (NS - 1)) == 0 ? (j & (NS - 1)) : (NS >= NB ? j : (T % NS == 0 ? (int)threadIdx.x % NS : j % NS));
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[pad_idx(j + r * NB)];
        if (NS > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const int idx = r * k * STEP;                       // < N
                float2 w = tws[idx < HALF ? idx : idx - HALF];
                if (idx >= HALF) { w.x = -w.x; w.y = -w.y; }
                v[r] = cmul(v[r], w);
            }
        }
        butterfly<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[pad_idx(j0 + r * NS)] = v[r];
    }
}

template <int N, int S, int NS>
__device__ __forceinline__ float2* fft_stages(float2* src, float2* dst, const float2* tws) {
    using P = Plan<N>;
    if constexpr (S == P::kStages) {
        return src;
    } else {
        constexpr int R = P::radix(S);
        stockham_pass<N, P::kThreads, R, NS>(src, dst, tws);
        __syncthreads();
        return fft_stages<N, S + 1, NS * R>(dst, src, tws);
    }
}

// Forward complex FFT of length N on (padded) shared memory.  Returns the buffer holding the result.
template <int N>
__device__ __forceinline__ float2* fft_forward(float2* a, float2* b, const float2* tws) {
    return fft_stages<N, 0, 1>(a, b, tws);
}

// Cooperative copy of the half twiddle table into shared memory (once per CTA).
template <int N>
__device__ __forceinline__ void load_twiddles(float2* tws, const float2* __restrict__ tw) {
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) tws[i] = __ldg(tw + i);
    __syncthreads();
}

__device__ __forceinline__ int reflect_index(int j, int len) {
    if (j < 0) j = -j;
    if (j >= len) j = 2 * (len - 1) - j;
    return j;
}

// Windowed frame `t` of a (reflect-padded, centred) signal packed as n_fft/2 complex values, then the real-input
// FFT.  src(j) returns sample j of the un-padded signal.  On return spec[0..n] (n+1 bins) holds the one-sided
// spectrum; `other` is the second scratch buffer.  Both buffers hold n+1 float2.
template <int N, typename SrcFn>
__device__ __forceinline__ float2* stft_frame(int t, int hop, int len, SrcFn src,
                                              const float* __restrict__ window, const float2* tw, const float2* tw_half,
                                              float2* a, float2* b) {
    constexpr int n = N;
    const int p0 = t * hop - n;                       // first padded-domain sample of the frame, relative to signal
    const float2* w2 = reinterpret_cast<const float2*>(window);
    // loads are issued in batches of kGather per thread before any use: one memory round trip per batch
    // instead of one per element (the frame loop is latency-bound otherwise -- measured)
    constexpr int kGather = 5;
    for (int base = threadIdx.x; base < n; base += kGather * blockDim.x) {
        float x0[kGather], x1[kGather];
        float2 w[kGather];
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
            const int m = base + u * blockDim.x;
            if (m < n) {
                x0[u] = src(reflect_index(p0 + 2 * m, len));
                x1[u] = src(reflect_index(p0 + 2 * m + 1, len));
                w[u] = __ldg(w2 + m);
            }
        }
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
            const int m = base + u * blockDim.x;
            if (m < n) a[pad_idx(m)] = make_float2(x0[u] * w[u].x, x1[u] * w[u].y);
        }
    }
    __syncthreads();
    float2* z = fft_forward<N>(a, b, tw);
    float2* o = (z == a) ? b : a;
    // X[k] = (Z[k] + conj(Z[n-k]))/2 - (i/2) e^{-2 pi i k/(2n)} (Z[k] - conj(Z[n-k])),  k = 0..n  (Z[n] == Z[0])
    // (z is padded, the spectrum o is written densely)
    constexpr int kPost = 6;
    for (int base = threadIdx.x; base <= n; base += kPost * blockDim.x) {
        float2 wh[kPost];
#pragma unroll
        for (int u = 0; u < kPost; ++u) {
            const int k = base + u * blockDim.x;
            if (k <= n) wh[u] = __ldg(tw_half + k);
        }
#pragma unroll
        for (int u = 0; u < kPost; ++u) {
            const int k = base + u * blockDim.x;
            if (k <= n) {
                const float2 zk = z[pad_idx(k == n ? 0 : k)];
                const float2 zc = z[pad_idx(k == 0 ? 0 : n - k)];
                const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
                const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));
                o[k] = cadd(e, cmul(mul_neg_i(d), wh[u]));
            }
        }
    }
    __syncthreads();
    return o;
}

// ------------------------------------------------------------------------------------------
// mel-STFT encoder
// ------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Plan<N>::kThreads)
stft_mel_kernel(const float* __restrict__ raw, int len, const float* __restrict__ window,
                const float* __restrict__ window2, const float* __restrict__ coef1, const float* __restrict__ coef2,
                const float2* __restrict__ tw, const float2* __restrict__ tw_half, int hop, int n_frames,
                const int* __restrict__ fb_start, const int* __restrict__ fb_count, const int* __restrict__ fb_offset,
                const float* __restrict__ fb_weight, int n_filters, float exponent, float mean, float scale,
                float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    constexpr int n = N;
    constexpr int buf = pad_idx(n) + 8;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + buf;
    float2* tws = b + buf;                                        // [n/2] twiddle half table
    float* mag = reinterpret_cast<float*>(tws + n / 2);           // [n + 8] blended magnitudes
    float* tile = mag + n + 8;                                    // [n_filters][kEncFrames + 1]
    load_twiddles<N>(tws, tw);
    const int s = blockIdx.y;
    const int t0 = blockIdx.x * kEncFrames;
    const float* sig = raw + (size_t)s * len;
    for (int f = 0; f < kEncFrames; ++f) {
        const int t = t0 + f;
        if (t >= n_frames) break;
        // |STFT| with the first window (times an optional per-bin coefficient); the live MS_MDCT_DualFormat blends a
        // second, narrower-window STFT per bin (ms_mdct_dual.py:249-256): mag = |X1|*coef1 + |X2|*coef2
        float2* spec = stft_frame<N>(t, hop, len, [&](int j) { return __ldg(sig + j); }, window, tws, tw_half, a, b);
        for (int k = threadIdx.x; k <= n; k += blockDim.x) {
            const float m1 = sqrtf(spec[k].x * spec[k].x + spec[k].y * spec[k].y);
            mag[k] = coef1 ? m1 * __ldg(coef1 + k) : m1;
        }
        __syncthreads();
        if (window2) {
            spec = stft_frame<N>(t, hop, len, [&](int j) { return __ldg(sig + j); }, window2, tws, tw_half, a, b);
            for (int k = threadIdx.x; k <= n; k += blockDim.x)
                mag[k] += sqrtf(spec[k].x * spec[k].x + spec[k].y * spec[k].y) * __ldg(coef2 + k);
            __syncthreads();
        }
        for (int m = threadIdx.x; m < n_filters; m += blockDim.x) {
            const int st = fb_start[m], cnt = fb_count[m], off = fb_offset[m];
            float acc = 0.f;
            for (int j = 0; j < cnt; ++j) acc += mag[st + j] * __ldg(fb_weight + off + j);
            const float v = (exponent == 0.25f) ? sqrtf(sqrtf(acc)) : (exponent == 1.f ? acc : powf(acc, exponent));
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
template <int N>
__global__ void __launch_bounds__(Plan<N>::kThreads)
fgla_istft_kernel(const float2* __restrict__ state, const float* __restrict__ mag, int stereo, float interp_t,
                  const float* __restrict__ window, const float2* __restrict__ tw, const float2* __restrict__ tw_half,
                  int hop, int n_frames, float* __restrict__ ola, int ola_len) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    constexpr int n = N, bins = N + 1;
    constexpr int buf = pad_idx(n) + 8;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + buf;
    float2* tws = b + buf;                                        // [n/2] twiddle half table
    float* acc = reinterpret_cast<float*>(tws + n / 2);           // [(kOlaFrames-1)*hop + 2n]
    load_twiddles<N>(tws, tw);
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
        // X[k] = angle * magnitude  (phase_recovery.py:84-95); stored in b.  Loads batched kBatch per thread.
        constexpr int kBatch = 6;
        for (int base = threadIdx.x; base < bins; base += kBatch * blockDim.x) {
            float2 tv[kBatch];
            float m0[kBatch], m1[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * blockDim.x;
                if (k < bins) {
                    m0[u] = __ldg(mag + row + k);
                    m1[u] = stereo ? __ldg(mag + row_other + k) : 0.f;
                    tv[u] = state ? __ldg(state + row + k) : make_float2(1.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * blockDim.x;
                if (k < bins) {
                    float m = m0[u];
                    if (stereo) {
                        const float merged = 0.5f * (m + m1[u]);                           // :63-64 (L+R)/2
                        m = interp_t > 0.f ? merged + interp_t * (m - merged) : merged;    // :86-88 lerp(merged, spec, t)
                    }
                    float2 ang = tv[u];
                    if (state) {
                        const float inv = 1.f / (sqrtf(ang.x * ang.x + ang.y * ang.y) + 1e-16f);   // :115
                        ang = make_float2(ang.x * inv, ang.y * inv);
                    }
                    float2 x = make_float2(ang.x * m, ang.y * m);
                    if (k == 0 || k == n) x.y = 0.f;    // C2R transforms ignore the imaginary part of DC / Nyquist
                    b[k] = x;
                }
            }
        }
        __syncthreads();
        // Z[k] = E[k] + i O[k], E = (X[k] + conj(X[n-k]))/2, O = e^{+2 pi i k/(2n)} (X[k] - conj(X[n-k]))/2;
        // inverse FFT through the forward one: IFFT(Z) = conj(FFT(conj(Z))) / n  ->  a holds conj(Z)
        for (int base = threadIdx.x; base < n; base += kBatch * blockDim.x) {
            float2 wh[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * blockDim.x;
                if (k < n) wh[u] = __ldg(tw_half + k);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * blockDim.x;
                if (k < n) {
                    const float2 xk = b[k], xc = b[n - k];
                    const float2 e = make_float2(0.5f * (xk.x + xc.x), 0.5f * (xk.y - xc.y));
                    const float2 d = make_float2(0.5f * (xk.x - xc.x), 0.5f * (xk.y + xc.y));
                    const float2 o = cmul(d, make_float2(wh[u].x, -wh[u].y));
                    a[pad_idx(k)] = make_float2(e.x - o.y, -(e.y + o.x));               // conj(E + iO)
                }
            }
        }
        __syncthreads();
        const float2* z = fft_forward<N>(a, b, tws);
        float* dst = acc + f * hop;
        for (int base = threadIdx.x; base < n; base += kBatch * blockDim.x) {
            float2 wv[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int m = base + u * blockDim.x;
                if (m < n) wv[u] = __ldg(reinterpret_cast<const float2*>(window) + m);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int m = base + u * blockDim.x;
                if (m < n) {
                    const float2 v = z[pad_idx(m)];
                    float2* d2 = reinterpret_cast<float2*>(dst) + m;
                    float2 cur = *d2;
                    cur.x += v.x * inv_n * wv[u].x;
                    cur.y += -v.y * inv_n * wv[u].y;
                    *d2 = cur;
                }
            }
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
template <int N>
__global__ void __launch_bounds__(Plan<N>::kThreads)
fgla_stft_update_kernel(const float* __restrict__ ola, const float* __restrict__ env, int ola_len, int len,
                        const float* __restrict__ window, const float2* __restrict__ tw,
                        const float2* __restrict__ tw_half, int hop, int n_frames,
                        float2* __restrict__ state, float momentum, int first) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    constexpr int n = N, bins = N + 1;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + pad_idx(n) + 8;
    float2* tws = b + pad_idx(n) + 8;
    load_twiddles<N>(tws, tw);
    const int s = blockIdx.y;
    const float* o = ola + (size_t)s * ola_len + n;       // trim n_fft/2 (center=True)
    const float* e = env + n;
    for (int f = 0; f < kEncFrames; ++f) {
        const int t = blockIdx.x * kEncFrames + f;
        if (t >= n_frames) break;
        float2* spec = stft_frame<N>(t, hop, len, [&](int j) { return __ldg(o + j) / __ldg(e + j); }, window, tws,
                                     tw_half, a, b);
        float2* row = state + ((size_t)s * n_frames + t) * bins;
        constexpr int kBatch = 6;
        for (int base = threadIdx.x; base < bins; base += kBatch * blockDim.x) {
            float2 pv[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * blockDim.x;
                pv[u] = (!first && k < bins) ? row[k] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * blockDim.x;
                if (k < bins) {
                    const float2 r = spec[k];
                    row[k] = make_float2(r.x - momentum * pv[u].x, r.y - momentum * pv[u].y);
                }
            }
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

bool supported_n_fft(int n_fft) { return n_fft == 6400 || n_fft == 4096; }
size_t fft_smem_bytes(int n) { return ((size_t)2 * (pad_idx(n) + 8) + n / 2) * sizeof(float2); }

template <typename K>
cudaError_t raise_smem_limit(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

extern "C" int dd_stft_mel(const float* raw, int n_signals, int len, const float* window, const float* window2,
                           const float* coef1, const float* coef2, const float* twiddles,
                           const float* twiddles_half, int n_fft, int hop, const int* fb_start, const int* fb_count,
                           const int* fb_offset, const float* fb_weight, int n_filters, float exponent, float mean,
                           float scale, float* out, int n_frames, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(raw && window && twiddles && twiddles_half && fb_start && fb_count && fb_offset && fb_weight && out,
               "dd_stft_mel: null pointer");
    DD_REQUIRE(supported_n_fft(n_fft), "dd_stft_mel: n_fft=%d unsupported (6400, 4096)", n_fft);
    DD_REQUIRE(!window2 || coef2, "dd_stft_mel: a second window needs its per-bin coefficients");
    DD_REQUIRE(len > n_fft / 2, "dd_stft_mel: signal shorter than the reflect padding");
    DD_REQUIRE(n_frames == 1 + len / hop, "dd_stft_mel: n_frames must be 1 + len/hop (center=True)");
    if (n_signals == 0) return 0;
    const size_t smem = fft_smem_bytes(n_fft / 2) + (size_t)(n_fft / 2 + 8 + n_filters * (kEncFrames + 1)) * sizeof(float);
    const dim3 grid(ceil_div(n_frames, kEncFrames), n_signals);
    const float2* tw = reinterpret_cast<const float2*>(twiddles);
    const float2* twh = reinterpret_cast<const float2*>(twiddles_half);
#define DD_ENC(N_)                                                                                                 \
    do {                                                                                                           \
        DD_CHECK_CUDA(raise_smem_limit(stft_mel_kernel<N_>, smem));                                                \
        stft_mel_kernel<N_><<<grid, Plan<N_>::kThreads, smem, stream>>>(raw, len, window, window2, coef1, coef2,   \
                                                                         tw, twh, hop, n_frames,                    \
                                                                         fb_start, fb_count, fb_offset, fb_weight,  \
                                                                         n_filters, exponent, mean, scale, out);    \
    } while (0)
    if (n_fft == 6400) DD_ENC(3200); else DD_ENC(2048);
#undef DD_ENC
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_fgla_istft(const float* state, const float* mag_tk, int n_signals, int n_frames, int stereo,
                             float interp_t, const float* window, const float* twiddles, const float* twiddles_half,
                             int n_fft, int hop, float* ola, int ola_len, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(mag_tk && window && twiddles && twiddles_half && ola, "dd_fgla_istft: null pointer");
    DD_REQUIRE(!stereo || n_signals % 2 == 0, "dd_fgla_istft: stereo needs an even number of signals");
    DD_REQUIRE(supported_n_fft(n_fft), "dd_fgla_istft: n_fft=%d unsupported (6400, 4096)", n_fft);
    DD_REQUIRE(ola_len == n_fft + hop * (n_frames - 1), "dd_fgla_istft: ola_len must be n_fft + hop*(n_frames-1)");
    DD_REQUIRE(n_fft <= kOlaFrames * hop, "dd_fgla_istft: n_fft/hop=%d overlaps more than %d frames", n_fft / hop,
               kOlaFrames);
    if (n_signals == 0) return 0;
    DD_CHECK_CUDA(cudaMemsetAsync(ola, 0, (size_t)n_signals * ola_len * sizeof(float), stream));
    const size_t smem = fft_smem_bytes(n_fft / 2) + (size_t)((kOlaFrames - 1) * hop + n_fft) * sizeof(float);
    const dim3 grid(ceil_div(n_frames, kOlaFrames), n_signals);
    const float2* tw = reinterpret_cast<const float2*>(twiddles);
    const float2* twh = reinterpret_cast<const float2*>(twiddles_half);
#define DD_ISTFT(N_)                                                                                               \
    do {                                                                                                           \
        DD_CHECK_CUDA(raise_smem_limit(fgla_istft_kernel<N_>, smem));                                              \
        fgla_istft_kernel<N_><<<grid, Plan<N_>::kThreads, smem, stream>>>(reinterpret_cast<const float2*>(state),  \
                                                                           mag_tk, stereo, interp_t, window, tw,    \
                                                                           twh, hop, n_frames, ola, ola_len);       \
    } while (0)
    if (n_fft == 6400) DD_ISTFT(3200); else DD_ISTFT(2048);
#undef DD_ISTFT
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_fgla_stft_update(const float* ola, const float* env, int n_signals, int n_frames, int len,
                                   const float* window, const float* twiddles, const float* twiddles_half, int n_fft,
                                   int hop, float* state, float momentum, int first, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(ola && env && window && twiddles && twiddles_half && state, "dd_fgla_stft_update: null pointer");
    DD_REQUIRE(supported_n_fft(n_fft), "dd_fgla_stft_update: n_fft=%d unsupported (6400, 4096)", n_fft);
    DD_REQUIRE(len == hop * (n_frames - 1), "dd_fgla_stft_update: len must be hop*(n_frames-1)");
    if (n_signals == 0) return 0;
    const size_t smem = fft_smem_bytes(n_fft / 2);
    const dim3 grid(ceil_div(n_frames, kEncFrames), n_signals);
    const float2* tw = reinterpret_cast<const float2*>(twiddles);
    const float2* twh = reinterpret_cast<const float2*>(twiddles_half);
    const int ola_len = n_fft + hop * (n_frames - 1);
#define DD_UPD(N_)                                                                                                 \
    do {                                                                                                           \
        DD_CHECK_CUDA(raise_smem_limit(fgla_stft_update_kernel<N_>, smem));                                        \
        fgla_stft_update_kernel<N_><<<grid, Plan<N_>::kThreads, smem, stream>>>(                                   \
            ola, env, ola_len, len, window, tw, twh, hop, n_frames, reinterpret_cast<float2*>(state), momentum, first); \
    } while (0)
    if (n_fft == 6400) DD_UPD(3200); else DD_UPD(2048);
#undef DD_UPD
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

