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

constexpr int kEncFrames = 64;      // frames per CTA in the FGLA forward STFT
constexpr int kMelFrames = 8;       // frames per CTA in the encoder (one 32 B output sector per filter; 2 CTAs/SM fit)
constexpr int kOlaFrames = 64;      // frames per CTA in the inverse STFT (overlap-add ring in shared memory)

// Only the buffer between the first two Stockham passes is XOR-swizzled (inside aligned groups of 16 complex
// values): pass 0 scatters with a stride of `radix` elements -- a 16-way conflict on 8-byte elements when stored
// plainly -- while every read is a contiguous run, conflict-free under any in-group permutation.  Later passes
// write runs of >= 8 elements and need nothing.  (History: `i + i/8` padding cost an extra wavefront on every
// contiguous access, 47% of all shared wavefronts in ncu; swizzling every buffer cost ~40 integer ops per point.)
template <bool S> __device__ __forceinline__ int swz(int i) { return S ? (i ^ ((i >> 4) & 15)) : i; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float wr, float wi) {
    return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
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
    o1 = cmulc(o1, h, -h);                      // w8^1
    o2 = mul_neg_i(o2);                         // w8^2
    o3 = cmulc(o3, -h, -h);                     // w8^3
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
template <>
__device__ __forceinline__ void butterfly<10>(float2 (&v)[10]) {
    // 10 = 5 x 2 in registers: X[k1 + 5 k2] = E[k1] + (-1)^k2 w10^k1 O[k1]
    float2 e[5] = {v[0], v[2], v[4], v[6], v[8]}, o[5] = {v[1], v[3], v[5], v[7], v[9]};
    butterfly<5>(e);
    butterfly<5>(o);
    o[1] = cmulc(o[1], 0.80901699437494742f, -0.58778525229247313f);
    o[2] = cmulc(o[2], 0.30901699437494742f, -0.95105651629515357f);
    o[3] = cmulc(o[3], -0.30901699437494742f, -0.95105651629515357f);
    o[4] = cmulc(o[4], -0.80901699437494742f, -0.58778525229247313f);
#pragma unroll
    for (int k = 0; k < 5; ++k) { v[k] = cadd(e[k], o[k]); v[k + 5] = csub(e[k], o[k]); }
}

// Compile-time FFT plans (complex length N = n_fft/2, T threads per frame).  Everything that depends on the plan
// -- radices, strides, twiddle steps, loop trip counts -- folds into immediates: the first, runtime-planned version
// of these kernels spent ~230 instructions per complex element on index arithmetic.
template <int N> struct Plan;
template <> struct Plan<3200> {            // n_fft 6400 = 2^8 * 5^2 (reference SpectrogramFormat)
    static constexpr int kThreads = 320, kStages = 4;
    static constexpr int radix(int s) { return s == 0 ? 8 : s == 1 ? 8 : s == 2 ? 5 : 10; }
};
template <> struct Plan<2048> {            // n_fft 4096 (live MS_MDCT_DualFormat windows)
    static constexpr int kThreads = 256, kStages = 4;
    static constexpr int radix(int s) { return s == 0 ? 8 : s == 1 ? 8 : s == 2 ? 8 : 4; }
};

// Per-thread twiddle registers.  In pass S (NS = product of the earlier radices) a thread's butterflies use
// w^(r k), k = j % NS, and j = tid + it*T: whenever T % NS == 0 (or the pass has a single iteration) k is the same
// for every frame AND every iteration, so w^k and w^2k are loaded from the global table once per CTA and stay in
// registers; the higher powers are products at most two multiplications deep.  (History: a shared-memory table
// cost 12.8 KB and its strided reads were 2- to 8-way bank-conflicted; __ldg twiddles came from L2 every pass.)
template <int N> struct TwRegs {
    using P = Plan<N>;
    static constexpr int ns(int S) { int v = 1; for (int s = 0; s < S; ++s) v *= P::radix(s); return v; }
    static constexpr int iters(int S) { return (N / P::radix(S) + P::kThreads - 1) / P::kThreads; }
    static constexpr int sets(int S) { return S == 0 ? 0 : (P::kThreads % ns(S) == 0 ? 1 : iters(S)); }
    static constexpr int base(int S) { int v = 0; for (int s = 1; s < S; ++s) v += sets(s); return v; }
    static constexpr int kSets = base(P::kStages);
    float2 w1[kSets], w2[kSets];
};

// (asm volatile: a plain __ldg is rematerialisable, and under register pressure the compiler re-issued the global
// loads inside every pass instead of keeping the values -- visible as long-scoreboard stalls on the twiddle multiply)
__device__ __forceinline__ float2 ld_pinned(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

template <int N, int S>
__device__ __forceinline__ void tw_init(TwRegs<N>& tw, const float2* __restrict__ table) {
    using P = Plan<N>;
    using TW = TwRegs<N>;
    if constexpr (S < P::kStages) {
        constexpr int R = P::radix(S), NS = TW::ns(S), STEP = N / (NS * R);
#pragma unroll
        for (int it = 0; it < TW::sets(S); ++it) {
            const int k = ((int)threadIdx.x + it * P::kThreads) % NS;
            tw.w1[TW::base(S) + it] = ld_pinned(table + k * STEP);                    // k*STEP < N/R <= N/2
            tw.w2[TW::base(S) + it] = R > 2 ? ld_pinned(table + 2 * k * STEP) : make_float2(1.f, 0.f);
        }
        tw_init<N, S + 1>(tw, table);
    }
}

// Only the buffer between the first two Stockham passes is XOR-swizzled (see swz()).
template <int S> struct SmemLoad {
    const float2* buf;
    __device__ __forceinline__ float2 operator()(int i) const { return buf[swz<S == 1>(i)]; }
};
template <int S> struct SmemStore {
    float2* buf;
    __device__ __forceinline__ void operator()(int i, float2 v) const { buf[swz<S == 0>(i)] = v; }
};

// One Stockham autosort pass S of radix R with NS = product of the previous radices.  load(i) yields input element i,
// store(i, v) consumes output element i: the first pass of a forward STFT reads the windowed signal straight from the
// sample ring and the last pass of the inverse STFT accumulates straight into the overlap-add ring, which saves one
// shared-memory round trip and one barrier per frame each.
template <int N, int S, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void fft_pass(const TwRegs<N>& tw, LoadFn load, StoreFn store) {
    using P = Plan<N>;
    using TW = TwRegs<N>;
    constexpr int T = P::kThreads, R = P::radix(S), NS = TW::ns(S), NB = N / R;
    constexpr int ITERS = (NB + T - 1) / T;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int j = (int)threadIdx.x + it * T;
        if (NB % T != 0 && j >= NB) break;
        const int k = (NS & (NS - 1)) == 0 ? (j & (NS - 1)) : (NS >= NB ? j : (T % NS == 0 ? (int)threadIdx.x % NS : j % NS));
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = load(j + r * NB);
        if constexpr (S > 0) {
            constexpr int kBase = TW::base(S);
            const int set = kBase + (TW::sets(S) == 1 ? 0 : it);
            float2 w[R];
            w[1] = tw.w1[set];
            if constexpr (R > 2) w[2] = tw.w2[set];
#pragma unroll
            for (int r = 3; r < R; ++r) {
                const int a = r == 3 ? 1 : r == 4 ? 2 : r < 9 ? 4 : 8;
                w[r] = cmul(w[a], w[r - a]);
            }
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
        }
        butterfly<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) store(j0 + r * NS, v[r]);
    }
}

// Passes S .. kStages-2 between the two shared-memory buffers, a barrier after each.  Returns the buffer holding
// the input of the last pass.
template <int N, int S>
__device__ __forceinline__ float2* fft_middle(const TwRegs<N>& tw, float2* src, float2* dst) {
    if constexpr (S >= Plan<N>::kStages - 1) {
        return src;
    } else {
        fft_pass<N, S>(tw, SmemLoad<S>{src}, SmemStore<S>{dst});
        __syncthreads();
        return fft_middle<N, S + 1>(tw, dst, src);
    }
}

__device__ __forceinline__ int reflect_index(int j, int len) {
    if (j < 0) j = -j;
    if (j >= len) j = 2 * (len - 1) - j;
    return j;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Pull `bytes` starting at p (any alignment) into L2, one 128 B line per thread per step.
template <int T>
__device__ __forceinline__ void prefetch_span_l2(const void* p, size_t bytes) {
    const char* c = static_cast<const char*>(p);
    for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)T * 128) prefetch_l2(c + off);
}

// The forward STFT kernels keep the last n_fft signal samples of their frame sequence in a shared-memory ring
// (sample q of the CTA's padded-domain span lives at q % n_fft): consecutive frames overlap by n_fft - hop samples,
// so only `hop` new samples per frame come from global memory (reflect padding, envelope division) instead of
// n_fft -- that gather was 31% of all stall samples before.  fetch(q) returns span sample q.
template <int N, typename FetchFn>
__device__ __forceinline__ void ring_fill(float* ring, FetchFn fetch) {
    constexpr int RING = 2 * N, T = Plan<N>::kThreads;
    constexpr int kU = 5;
    for (int base = threadIdx.x; base < RING; base += kU * T) {
        float x[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) if (base + u * T < RING) x[u] = fetch(base + u * T);
#pragma unroll
        for (int u = 0; u < kU; ++u) if (base + u * T < RING) ring[base + u * T] = x[u];
    }
}

// Half-length complex FFT of the windowed frame f (local index; its samples are ring span [f*hop, f*hop + n_fft)).
// The first pass reads ring * window directly (`win` may point to shared or global memory).  While the second pass
// runs, the `hop` samples the NEXT frame adds are fetched (if `more`) and dropped into the ring slots this frame no
// longer needs.  Z ends up in `b` (kStages is even), with a barrier behind it; the caller unpacks the real-input
// spectrum with spectrum_pairs().  The next call may start without a barrier: its first pass writes `a`.
template <int N, typename FetchFn>
__device__ __forceinline__ const float2* stft_fft(const TwRegs<N>& tw, float* ring, int f, int hop, bool more, FetchFn fetch,
                                                  const float2* win, float2* a, float2* b) {
    using P = Plan<N>;
    static_assert(P::kStages % 2 == 0 && P::kStages >= 4, "buffer rotation below assumes an even number of passes");
    constexpr int T = P::kThreads, RING = 2 * N;
    const int ring0 = (f * hop) % RING;               // even (hop is even): float2-aligned
    fft_pass<N, 0>(tw, [&](int i) {
        int p = ring0 + 2 * i;
        if (p >= RING) p -= RING;
        const float2 x = *reinterpret_cast<const float2*>(ring + p);
        const float2 w = win[i];
        return make_float2(x.x * w.x, x.y * w.y);
    }, SmemStore<0>{a});
    __syncthreads();
    constexpr int kNew = 2;
    decltype(fetch(0)) nx[kNew];
#pragma unroll
    for (int u = 0; u < kNew; ++u) {
        const int i = (int)threadIdx.x + u * T;
        if (more && i < hop) nx[u] = fetch(f * hop + RING + i);
    }
    fft_pass<N, 1>(tw, SmemLoad<1>{a}, SmemStore<1>{b});
#pragma unroll
    for (int u = 0; u < kNew; ++u) {
        const int i = (int)threadIdx.x + u * T;
        if (more && i < hop) { int p = ring0 + i; if (p >= RING) p -= RING; ring[p] = nx[u]; }
    }
    if (more)
        for (int i = (int)threadIdx.x + kNew * T; i < hop; i += T) ring[(ring0 + i) % RING] = fetch(f * hop + RING + i);
    __syncthreads();
    float2* src = fft_middle<N, 2>(tw, b, a);
    float2* dst = (src == a) ? b : a;
    fft_pass<N, P::kStages - 1>(tw, SmemLoad<P::kStages - 1>{src}, SmemStore<P::kStages - 1>{dst});
    __syncthreads();
    return dst;
}

// Real-input spectrum from the half-length FFT Z, two bins per step:
//   X[k]   = E + (-i) D w_k,  X[n-k] = conj(E) - i conj(D w_k),   E = (Z[k] + conj(Z[n-k]))/2, D = (Z[k] - conj(Z[n-k]))/2,
//   w_k = e^{-2 pi i k/(2n)}  (w_{n-k} = -conj(w_k)),  Z[n] == Z[0];  k = 0 yields X[0] and X[n] (DC / Nyquist).
// pre(k) is fetched for a whole batch before any arithmetic (global operands of the sink, e.g. the previous
// Griffin-Lim state); sink(k, X[k], pre(k)) consumes each bin -- nothing is staged through shared memory.
template <int N, typename PreFn, typename SinkFn>
__device__ __forceinline__ void spectrum_pairs(const float2* z, const float2* __restrict__ tw_half, PreFn pre, SinkFn sink) {
    constexpr int n = N, T = Plan<N>::kThreads, HALFN = N / 2;
    constexpr int kPost = 5;
    using PreT = decltype(pre(0));
    for (int base = threadIdx.x; base < HALFN; base += kPost * T) {
        float2 wh[kPost];
        PreT pa[kPost], pb[kPost];
#pragma unroll
        for (int u = 0; u < kPost; ++u) {
            const int k = base + u * T;
            if (k < HALFN) { wh[u] = __ldg(tw_half + k); pa[u] = pre(k); pb[u] = pre(n - k); }
        }
#pragma unroll
        for (int u = 0; u < kPost; ++u) {
            const int k = base + u * T;
            if (k < HALFN) {
                const float2 zk = z[k];
                const float2 zc = z[k == 0 ? 0 : n - k];
                const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
                const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));
                const float2 q = cmul(d, wh[u]);
                sink(k, make_float2(e.x + q.y, e.y - q.x), pa[u]);
                sink(n - k, make_float2(e.x - q.y, -e.y - q.x), pb[u]);
            }
        }
    }
    if (threadIdx.x == 0) {                           // k = n/2 pairs with itself: w = -i, X = conj(Z[n/2])
        const float2 zk = z[HALFN];
        sink(HALFN, make_float2(zk.x, -zk.y), pre(HALFN));
    }
}

// ------------------------------------------------------------------------------------------
// mel-STFT encoder
// ------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Plan<N>::kThreads, 2)
stft_mel_kernel(const float* __restrict__ raw, int len, const float* __restrict__ window,
                const float* __restrict__ window2, const float* __restrict__ coef1, const float* __restrict__ coef2,
                const float2* __restrict__ tw, const float2* __restrict__ tw_half, int hop, int n_frames,
                const int* __restrict__ fb_start, const int* __restrict__ fb_count, const int* __restrict__ fb_offset,
                const float* __restrict__ fb_weight, int n_filters, float exponent, float mean, float scale,
                float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    constexpr int n = N;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + n;
    float* ring = reinterpret_cast<float*>(b + n);                // [2n] signal ring
    float* mag = ring + 2 * n;                                    // [n + 8] blended magnitudes
    float* tile = mag + n + 8;                                    // [n_filters][kMelFrames + 1]
    TwRegs<N> twr;
    tw_init<N, 1>(twr, tw);
    const int s = blockIdx.y;
    const int t0 = blockIdx.x * kMelFrames;
    const int nf = min(kMelFrames, n_frames - t0);
    const float* sig = raw + (size_t)s * len;
    const int p_base = t0 * hop - n;                  // signal index of span sample 0 (center=True reflect padding)
    auto fetch = [&](int q) { return __ldg(sig + reflect_index(p_base + q, len)); };
    ring_fill<N>(ring, fetch);
    __syncthreads();
    // the windows stay in global memory here (L2-resident; the 8 reads of a butterfly are issued together)
    const float2* win1 = reinterpret_cast<const float2*>(window);
    const float2* win2 = reinterpret_cast<const float2*>(window2);
    for (int f = 0; f < nf; ++f) {
        // |STFT| with the first window (times an optional per-bin coefficient); the live MS_MDCT_DualFormat blends a
        // second, narrower-window STFT per bin (ms_mdct_dual.py:249-256): mag = |X1|*coef1 + |X2|*coef2.  A bin is
        // handled by the same thread in both passes, so the accumulation needs no barrier.
        {
            const float2* z = stft_fft<N>(twr, ring, f, hop, !window2 && f + 1 < nf, fetch, win1, a, b);
            spectrum_pairs<N>(z, tw_half, [&](int k) { return coef1 ? __ldg(coef1 + k) : 1.f; },
                              [&](int k, float2 x, float c) { mag[k] = sqrtf(x.x * x.x + x.y * x.y) * c; });
        }
        if (window2) {
            const float2* z = stft_fft<N>(twr, ring, f, hop, f + 1 < nf, fetch, win2, a, b);
            spectrum_pairs<N>(z, tw_half, [&](int k) { return __ldg(coef2 + k); },
                              [&](int k, float2 x, float c) { mag[k] += sqrtf(x.x * x.x + x.y * x.y) * c; });
        }
        __syncthreads();
        for (int m = threadIdx.x; m < n_filters; m += blockDim.x) {
            const int st = fb_start[m], cnt = fb_count[m], off = fb_offset[m];
            float acc = 0.f;
            for (int j = 0; j < cnt; ++j) acc += mag[st + j] * __ldg(fb_weight + off + j);
            const float v = (exponent == 0.25f) ? sqrtf(sqrtf(acc)) : (exponent == 1.f ? acc : powf(acc, exponent));
            tile[m * (kMelFrames + 1) + f] = (v - mean) * scale;
        }
        // the next frame's first write to `mag` sits behind its FFT barriers
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_filters * kMelFrames; i += blockDim.x) {
        const int m = i / kMelFrames, f = i % kMelFrames;
        if (f < nf) out[((size_t)s * n_filters + m) * n_frames + t0 + f] = tile[m * (kMelFrames + 1) + f];
    }
}

// ------------------------------------------------------------------------------------------
// FGLA phase A: A = T/(|T|+1e-16); X = A * M_k; inverse STFT frame, window, overlap-add
// ------------------------------------------------------------------------------------------
struct FglaBin { float2 t; float m, mo; };

template <int N>
__global__ void __launch_bounds__(Plan<N>::kThreads, 2)
fgla_istft_kernel(const float2* __restrict__ state, const float* __restrict__ mag, int stereo, float interp_t,
                  const float* __restrict__ window, const float2* __restrict__ tw, const float2* __restrict__ tw_half,
                  int hop, int n_frames, float* __restrict__ ola, int ola_len) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    using P = Plan<N>;
    constexpr int n = N, bins = N + 1, T = P::kThreads, HALFN = N / 2, RING = 2 * N;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + n;
    float2* win = b + n;                                          // [n] synthesis window (sample pairs)
    float* ring = reinterpret_cast<float*>(win + n);              // [2n] overlap-add ring: sample p lives at p % 2n
    TwRegs<N> twr;
    tw_init<N, 1>(twr, tw);
    const int s = blockIdx.y;
    const int t0 = blockIdx.x * kOlaFrames;
    const int nf = min(kOlaFrames, n_frames - t0);
    for (int i = threadIdx.x; i < n; i += T) win[i] = __ldg(reinterpret_cast<const float2*>(window) + i);
    for (int i = threadIdx.x; i < RING; i += T) ring[i] = 0.f;
    // (ordered before their first use by the barriers of the first frame)
    const float inv_n = 1.f / (float)n;
    float* g = ola + (size_t)s * ola_len + (size_t)t0 * hop;      // global position of this CTA's local sample 0

    // X[k] = angle * magnitude  (phase_recovery.py:84-95)
    auto load_bin = [&](size_t row, size_t row_other, int k) {
        FglaBin v;
        v.m = __ldg(mag + row + k);
        v.mo = stereo ? __ldg(mag + row_other + k) : 0.f;
        v.t = state ? __ldg(state + row + k) : make_float2(1.f, 0.f);
        return v;
    };
    auto make_x = [&](const FglaBin& v, int k) {
        float m = v.m;
        if (stereo) {
            const float merged = 0.5f * (m + v.mo);                                // :63-64 (L+R)/2
            m = interp_t > 0.f ? merged + interp_t * (m - merged) : merged;        // :86-88 lerp(merged, spec, t)
        }
        float2 ang = v.t;
        if (state) {
            // :115 T / (|T| + 1e-16); the epsilon only matters at |T| ~ 0, where both forms give 0
            const float inv = rsqrtf(fmaxf(ang.x * ang.x + ang.y * ang.y, 1e-32f));
            ang = make_float2(ang.x * inv, ang.y * inv);
        }
        float2 x = make_float2(ang.x * m, ang.y * m);
        if (k == 0 || k == n) x.y = 0.f;        // C2R transforms ignore the imaginary part of DC / Nyquist
        return x;
    };
    // conj(Z[k]) and conj(Z[n-k]) of the packed half-length spectrum Z = E + iO from the bin pair (X[k], X[n-k]):
    // E = (X[k] + conj(X[n-k]))/2, O = e^{+2 pi i k/(2n)} (X[k] - conj(X[n-k]))/2; Z[n-k] = conj(E) + i conj(O).
    // The inverse FFT runs through the forward one: IFFT(Z) = conj(FFT(conj(Z))) / n.
    auto pack_pair = [&](float2 xk, float2 xc, float2 wh, float2& zk, float2& zc) {
        const float2 e = make_float2(0.5f * (xk.x + xc.x), 0.5f * (xk.y - xc.y));
        const float2 d = make_float2(0.5f * (xk.x - xc.x), 0.5f * (xk.y + xc.y));
        const float2 o = cmul(d, make_float2(wh.x, -wh.y));
        zk = make_float2(e.x - o.y, -(e.y + o.x));
        zc = make_float2(e.x + o.y, e.y - o.x);
    };
    auto prefetch_rows = [&](int t) {            // HBM -> L2 one frame ahead of the loads (nothing re-reads these rows)
        const size_t row = ((size_t)s * n_frames + t) * bins;
        if (state) prefetch_span_l2<T>(state + row, bins * sizeof(float2));
        prefetch_span_l2<T>(mag + row, bins * sizeof(float));
        if (stereo) prefetch_span_l2<T>(mag + ((size_t)(s ^ 1) * n_frames + t) * bins, bins * sizeof(float));
    };

    for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        const size_t row = ((size_t)s * n_frames + t) * bins;
        const size_t row_other = ((size_t)(s ^ 1) * n_frames + t) * bins;
        if (f + 1 < nf) prefetch_rows(t + 1);
        constexpr int kBatch = 5;
        for (int base = threadIdx.x; base < HALFN; base += kBatch * T) {
            FglaBin vk[kBatch], vc[kBatch];
            float2 wh[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * T;
                if (k < HALFN) {
                    vk[u] = load_bin(row, row_other, k);
                    vc[u] = load_bin(row, row_other, n - k);
                    wh[u] = __ldg(tw_half + k);
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int k = base + u * T;
                if (k < HALFN) {
                    float2 zk, zc;
                    pack_pair(make_x(vk[u], k), make_x(vc[u], n - k), wh[u], zk, zc);
                    a[k] = zk;
                    if (k > 0) a[n - k] = zc;
                }
            }
        }
        if (threadIdx.x == 0) {                       // k = n/2 pairs with itself
            const float2 x = make_x(load_bin(row, row_other, HALFN), HALFN);
            float2 zk, zc;
            pack_pair(x, x, __ldg(tw_half + HALFN), zk, zc);
            a[HALFN] = zk;
        }
        __syncthreads();
        fft_pass<N, 0>(twr, SmemLoad<0>{a}, SmemStore<0>{b});
        __syncthreads();
        float2* src = fft_middle<N, 1>(twr, b, a);
        // Last pass: windowed overlap-add straight into the ring.  The first `hop` samples of the frame are final
        // for this CTA once added (later frames start further right): they go straight to global memory and their
        // ring slots, re-used by the next frame's tail, are cleared.  The next frame's accumulation (other threads,
        // same slots) sits behind its packing + FFT barriers; so does its first write to `a` / `b`.
        const int ring0 = (f * hop) % RING;
        fft_pass<N, P::kStages - 1>(twr, SmemLoad<P::kStages - 1>{src}, [&](int m, float2 v) {
            int p = ring0 + 2 * m;
            if (p >= RING) p -= RING;
            float2* slot = reinterpret_cast<float2*>(ring + p);
            const float2 w = win[m];
            float2 cur = *slot;
            cur.x += v.x * inv_n * w.x;
            cur.y += -v.y * inv_n * w.y;
            if (2 * m < hop) {
                float* dst = g + (size_t)f * hop + 2 * m;
                atomicAdd(dst, cur.x);
                atomicAdd(dst + 1, cur.y);
                cur = make_float2(0.f, 0.f);
            }
            *slot = cur;
        });
    }
    __syncthreads();
    // tail: samples [nf*hop, (nf-1)*hop + 2n) of this CTA.  Every output sample is touched by at most two CTAs of
    // the same signal (kOlaFrames*hop >= n_fft), so the float atomics are order-independent (0 + a + b == 0 + b + a)
    // and the result is deterministic.
    const int tail0 = nf * hop, tail = RING - hop;
    for (int i = threadIdx.x; i < tail; i += T) atomicAdd(g + tail0 + i, ring[(tail0 + i) % RING]);
}

// ------------------------------------------------------------------------------------------
// FGLA phase B: rebuilt = STFT(ISTFT(..)); T <- rebuilt - momentum * T   (phase_recovery.py:97-117)
// ------------------------------------------------------------------------------------------

template <int N>
__global__ void __launch_bounds__(Plan<N>::kThreads, 2)
fgla_stft_update_kernel(const float* __restrict__ ola, const float* __restrict__ env, int ola_len, int len,
                        const float* __restrict__ window, const float2* __restrict__ tw,
                        const float2* __restrict__ tw_half, int hop, int n_frames,
                        float2* __restrict__ state, float momentum, int first) {
    extern __shared__ __align__(16) uint8_t smem_fft[];
    constexpr int n = N, bins = N + 1, T = Plan<N>::kThreads;
    float2* a = reinterpret_cast<float2*>(smem_fft);
    float2* b = a + n;
    float2* win = b + n;                                  // [n] analysis window (sample pairs)
    float* ring = reinterpret_cast<float*>(win + n);      // [2n] signal ring
    TwRegs<N> twr;
    tw_init<N, 1>(twr, tw);
    const int s = blockIdx.y;
    const float* o = ola + (size_t)s * ola_len + n;       // trim n_fft/2 (center=True)
    const float* e = env + n;
    const int t0 = blockIdx.x * kEncFrames;
    const int nf = min(kEncFrames, n_frames - t0);
    const int p_base = t0 * hop - n;
    for (int i = threadIdx.x; i < n; i += T) win[i] = __ldg(reinterpret_cast<const float2*>(window) + i);
    ring_fill<N>(ring, [&](int q) { const int j = reflect_index(p_base + q, len); return __ldg(o + j) / __ldg(e + j); });
    __syncthreads();
    // per-frame refill: the division waits until the ring store, so the loads fly during an FFT pass
    struct Lazy { float num, den; __device__ __forceinline__ operator float() const { return num / den; } };
    auto fetch = [&](int q) { const int j = reflect_index(p_base + q, len); return Lazy{__ldg(o + j), __ldg(e + j)}; };
    for (int f = 0; f < nf; ++f) {
        const int t = t0 + f;
        float2* row = state + ((size_t)s * n_frames + t) * bins;
        if (!first && f + 1 < nf) prefetch_span_l2<T>(row + bins, bins * sizeof(float2));    // HBM -> L2 a frame ahead
        const float2* z = stft_fft<N>(twr, ring, f, hop, f + 1 < nf, fetch, win, a, b);
        spectrum_pairs<N>(z, tw_half, [&](int k) { return first ? make_float2(0.f, 0.f) : row[k]; },
                          [&](int k, float2 x, float2 prev) {
                              row[k] = make_float2(x.x - momentum * prev.x, x.y - momentum * prev.y);
                          });
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
size_t fft_smem_bytes(int n) { return (size_t)2 * n * sizeof(float2); }      // the two FFT buffers

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
    DD_REQUIRE(hop > 0 && hop % 2 == 0 && hop <= n_fft, "dd_stft_mel: hop must be even and at most n_fft");
    const size_t smem = fft_smem_bytes(n_fft / 2) + (size_t)(n_fft + n_fft / 2 + 8 + n_filters * (kMelFrames + 1)) * sizeof(float);
    const dim3 grid(ceil_div(n_frames, kMelFrames), n_signals);
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
    DD_REQUIRE(hop > 0 && hop % 2 == 0 && hop <= n_fft, "dd_fgla_istft: hop must be even and at most n_fft");
    const size_t smem = fft_smem_bytes(n_fft / 2) + (size_t)2 * n_fft * sizeof(float);      // + window + OLA ring
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
    DD_REQUIRE(hop > 0 && hop % 2 == 0 && hop <= n_fft, "dd_fgla_stft_update: hop must be even and at most n_fft");
    const size_t smem = fft_smem_bytes(n_fft / 2) + (size_t)2 * n_fft * sizeof(float);      // + window + signal ring
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

