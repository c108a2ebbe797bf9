// Grouped 3x3 MPConv with the three horizontal taps stacked along the UMMA N dimension (tcgen05 + TMEM + TMA).
//
// Replaces `F.conv2d(x, w, padding=1, groups=g)` of MPConv.forward (/root/reference/src/modules/mp_tools.py:369) for the
// res0 / res1 convolutions of the EDM2 blocks (unet_edm2_b4.py:119-131) at the levels whose images are >= 8 rows tall.
//
// Why: with groups = 8 a group's GEMM is narrow (Cout/g = 32..128).  An SS-mode 128 x N x 16 UMMA reads its A slice
// (128 pixels x 16 channels = 4 KB) from shared memory whatever N is, so at N = 32 / 64 the 128 B/clk shared-memory port
// -- not the tensor pipe -- bounds it: 40 / 48 cycles against 16 / 32 (measured, profiles/r02_umma_issue_rates.log).
// Here one UMMA computes the partial sums of the three taps (dy, dx = -1 | 0 | +1) from the SAME pixel rows:
//     D_t[pixel p, co] = sum_ci x[p + dy*W, ci] * W[co][dy][t][ci],   t = dx + 1,   N = 3 * n_co  (96 or 192),
// i.e. A is read once for three taps and N >= 96 runs at 86-100 % of the tensor rate.  The shift along w moves to the
// epilogue:  out[h, w] = D_0[h, w-1] + D_1[h, w] + D_2[h, w+1].  A tile is 4 image rows x 32 columns whose first and last
// column are halo, one image row per TMEM lane quadrant, so the shift is a +-1 lane warp shuffle and every warp produces
// 30 output pixels of its own row (6.25 % of the MMA work is halo).  Vertical taps are 32-row (4 KB) shifts of the UMMA
// descriptor into one (4+2) x 32 pixel x 64 channel TMA box, as in the halo kernel.
//
//   warp 0      : TMA producer -- weight panel (resident per CTA, rows ordered [dy][t][co]), activation boxes, and for the
//                 residual epilogue the residual tile, all through mbarrier rings
//   warps 1, 2  : UMMA issuers on alternate units (immediates-only issue loop).  The tensor pipe queues about one UMMA
//                 ahead of the issuing thread, so a single issuer's per-tile bookkeeping (mbarrier probes ~120 cycles each,
//                 commits, descriptor arithmetic: 300-450 cycles) drains it -- profiles/r02_issue_overhead.log; with two
//                 issuers one warp's UMMAs cover the other's bookkeeping.  Warp 1 owns the TMEM allocation.
//   warps 3..18 : epilogue, 32 output channels of one image row per warp (4 groups of 4 warps on alternate tiles for
//                 n_co = 32, 2 groups of 8 for n_co = 64): tcgen05.ld x3 -> shuffle-add -> fused Block.forward glue
//                 (emb-gain * mp_silu | mp_sum residual + clip; packed f32x2 arithmetic) -> bf16 -> swizzled staging slab
//                 -> per-warp TMA store
// When Cin/g = 32 one 64-channel activation box serves the two groups it spans (two sub-tiles per box, no wasted half).
#include "common.cuh"
#include "conv_dx.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int kTileW = 32, kTileH = 4, kOutW = kTileW - 2;
constexpr uint32_t kABytes = (kTileH + 2) * kTileW * 128;      // one activation box: 6 x 32 pixels x 64 channels
constexpr int kMaxAStages = 6, kMaxAcc = 4, kMaxRes = 4;
constexpr int kEpiWarps = 16;           // 32 output channels per warp: 4 groups x 4 warps (n_co = 32) or 2 groups x 8 (n_co = 64)
constexpr int kEpiCh = 32;
constexpr int kMmaWarps = 2;             // UMMA-issuing warps on alternate units (one hides the other's bookkeeping)
constexpr int kFirstEpiWarp = 1 + kMmaWarps;
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);
constexpr uint32_t kSmemLimit = 226u * 1024u;

struct FastDiv {
    unsigned long long m;
    int d;
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = d;
    f.m = ((1ull << 44) + (unsigned long long)d - 1ull) / (unsigned long long)d;
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return (int)(((unsigned long long)(unsigned)n * f.m) >> 44); }

struct DxParams {
    int B, H, W, Cout;
    int cin_g, cout_g;
    int n_co;               // output channels per sub-tile (32 or 64); the accumulator is 3 * n_co columns wide
    int nsub;               // sub-tiles per activation box: 2 when Cin/g == 32 (the box spans two groups), else 1
    int kchunks, ks_last;   // 64-channel boxes per unit / 16-channel steps in the last one
    int tiles_w, tiles_h, m_tiles, npg, units;
    FastDiv fd_m, fd_npg, fd_tw, fd_th;
    int a_stages, nbuf, res_stages;
    int mma_warps, ring;    // UMMA-issuing warps in use (1 or 2) and the activation stages each of them owns
    int ngroups;            // epilogue groups (tiles in the epilogue at once); 16 / ngroups warps share a tile
    uint32_t b_row_bytes, b_dx_bytes, b_blk_bytes, b_sub_bytes;
    uint32_t off_a, off_res, off_slab, res_stride, res_bytes, slab_bytes;
    int epi, epi2;
    float alpha, beta, clip;
    const float* scale;
    const float* scale2;
    unsigned long long* trace;      // DD_CONV_TRACE=1: [4][12][64] clock64 stamps (slots as in conv_igemm.cu's TraceSlot)
};

template <bool TRACE>
__device__ __forceinline__ void trace_stamp(const DxParams& p, int slot, uint32_t item) {
    if constexpr (TRACE) {
        if (blockIdx.x < 4u && item < 64u && p.trace != nullptr)
            p.trace[((size_t)blockIdx.x * 12 + slot) * 64 + item] = (unsigned long long)clock64();
    }
}

struct Unit {
    int panel, b, h0, w0;
};
__device__ __forceinline__ Unit decode_unit(const DxParams& p, int u) {
    Unit t;
    t.panel = fdiv(u, p.fd_m);
    const int m = u - t.panel * p.m_tiles;
    const int r1 = fdiv(m, p.fd_tw);
    t.w0 = (m - r1 * p.tiles_w) * kOutW;
    t.b = fdiv(r1, p.fd_th);
    t.h0 = (r1 - t.b * p.tiles_h) * kTileH;
    return t;
}
// first output channel of sub-tile s of a panel / first input channel of the panel's activation boxes
__device__ __forceinline__ int panel_co0(const DxParams& p, int panel, int s) {
    const int gq = fdiv(panel, p.fd_npg), j = panel - gq * p.npg;
    return (gq * p.nsub + s) * p.cout_g + j * p.n_co;
}
__device__ __forceinline__ int panel_ci0(const DxParams& p, int panel) { return fdiv(panel, p.fd_npg) * p.nsub * p.cin_g; }

// Position in the unit sequence (panel-major, then batch item, tile row, tile column), advanced without divisions.
struct Walker {
    int panel, m, b, h0, w0;
    __device__ __forceinline__ void init(const DxParams& p, int u) {
        const Unit t = decode_unit(p, u);
        panel = t.panel; m = u - t.panel * p.m_tiles; b = t.b; h0 = t.h0; w0 = t.w0;
    }
    __device__ __forceinline__ bool next(const DxParams& p) {        // true when the next unit starts a new weight panel
        w0 += kOutW;
        if (w0 >= p.tiles_w * kOutW) {
            w0 = 0; h0 += kTileH;
            if (h0 >= p.tiles_h * kTileH) { h0 = 0; ++b; }
        }
        if (++m == p.m_tiles) { m = 0; b = 0; ++panel; return true; }
        return false;
    }
};

// Packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): half the issue slots of the epilogue arithmetic.
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// mp_silu of two values: x * sigmoid(x) / 0.596 = a * tanh(x / 2) + a with a = x * (0.5 / 0.596); one MUFU per value
__device__ __forceinline__ float2 mp_silu2(float2 x) {
    const float2 h = mul2(x, make_float2(0.5f, 0.5f));
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
    const float2 a = mul2(x, make_float2(0.5f / 0.596f, 0.5f / 0.596f));
    return fma2(a, t, a);
}
__device__ __forceinline__ float clampf(float x, float c) { return fminf(fmaxf(x, -c), c); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 pack_bf16x8(const float2* v) {
    return make_uint4(pack_bf16x2(v[0].x, v[0].y), pack_bf16x2(v[1].x, v[1].y), pack_bf16x2(v[2].x, v[2].y),
                      pack_bf16x2(v[3].x, v[3].y));
}
// Byte offset of 16 B chunk `chunk` of row `row` in a TMA tile whose rows are one swizzle span wide (128 B: SWIZZLE_128B,
// 64 B: SWIZZLE_64B); the tile base is 1024 B aligned, so the XOR pattern is a function of the row index alone.
__device__ __forceinline__ uint32_t swz_off(int row, int chunk, uint32_t row_bytes) {
    return row_bytes == 128u ? (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4))
                             : (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_addr, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 3 vertical taps x NKS 16-channel steps of one activation box: every descriptor is an immediate away from the base.
template <int NKS>
__device__ __forceinline__ void issue_box(uint64_t a_desc0, uint64_t b_desc0, uint32_t b_blk16, uint32_t d_tmem, uint32_t idesc,
                                          uint32_t acc_first) {
    uint64_t b_dy = b_desc0;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const uint64_t a_dy = a_desc0 + (uint64_t)(dy * ((kTileW * 128) >> 4));      // + 32 pixel rows
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            if (dy == 0 && ks == 0) ptx::umma_bf16_ss(d_tmem, a_dy, b_dy, idesc, acc_first);
            else ptx::umma_bf16_ss_acc(d_tmem, a_dy + 2 * ks, b_dy + 2 * ks, idesc);
        }
        b_dy += b_blk16;
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_dx_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                  const __grid_constant__ CUtensorMap tmR, const __grid_constant__ DxParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[kMaxAStages];
    __shared__ __align__(8) uint64_t a_empty[kMaxAStages];
    __shared__ __align__(8) uint64_t b_full, b_empty[kMmaWarps];     // one 'panel consumed' barrier per issuing warp
    __shared__ __align__(8) uint64_t acc_full[kMaxAcc];
    __shared__ __align__(8) uint64_t acc_empty[kMaxAcc];
    __shared__ __align__(8) uint64_t res_full[kMaxRes];
    __shared__ __align__(8) uint64_t res_empty[kMaxRes];
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) trace_stamp<TRACE>(p, 0, 63);          // kernel entry (items 62 / 63 are never real tiles)
    const int u_begin = (int)((long)blockIdx.x * p.units / gridDim.x);
    const int u_end = (int)((long)(blockIdx.x + 1) * p.units / gridDim.x);
    const bool has_res = p.epi == DD_EPI_RESIDUAL;
    const int ncols = 3 * p.n_co;

#ifdef DD_MBAR_DEBUG
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("barriers: a_full 0x%x a_empty 0x%x b_full 0x%x b_empty 0x%x acc_full 0x%x acc_empty 0x%x res_full 0x%x res_empty 0x%x units %d..%d\n",
               ptx::smem_u32(a_full), ptx::smem_u32(a_empty), ptx::smem_u32(&b_full), ptx::smem_u32(b_empty), ptx::smem_u32(acc_full),
               ptx::smem_u32(acc_empty), ptx::smem_u32(res_full), ptx::smem_u32(res_empty), u_begin, u_end);
#endif
    if (warp == 0) {            // spread over the lanes: the set-up is on every launch's critical path
        const uint32_t group_warps = (uint32_t)(kEpiWarps / p.ngroups);
        if (lane == 0) { ptx::prefetch_tensormap(&tmA); ptx::mbar_init(&b_full, 1); }
        if (lane == 1) ptx::prefetch_tensormap(&tmB);
        if (lane == 2) ptx::prefetch_tensormap(&tmO);
        if (lane == 3 && p.epi2 != DD_EPI2_NONE) ptx::prefetch_tensormap(&tmO2);
        if (lane == 4 && has_res) ptx::prefetch_tensormap(&tmR);
        if (lane < kMmaWarps) ptx::mbar_init(&b_empty[lane], 1);
        if (lane >= 8 && lane - 8 < p.a_stages) { ptx::mbar_init(&a_full[lane - 8], 1); ptx::mbar_init(&a_empty[lane - 8], 1); }
        if (lane >= 16 && lane - 16 < p.nbuf) { ptx::mbar_init(&acc_full[lane - 16], 1); ptx::mbar_init(&acc_empty[lane - 16], group_warps); }
        if (lane >= 24 && lane - 24 < p.res_stages) { ptx::mbar_init(&res_full[lane - 24], 1); ptx::mbar_init(&res_empty[lane - 24], group_warps); }
        ptx::mbar_fence_init();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) trace_stamp<TRACE>(p, 2, 63);          // barriers initialised
    }
    // Every role walks the same unit sequence with incremental counters: no integer division and no tile decode in the
    // per-tile control path (the issuing warp's bookkeeping between two tiles must stay shorter than the work queued in
    // the tensor pipe, or the pipe idles -- profiles/r02_trace_dx_v1.log).
    Walker wk;
    wk.init(p, u_begin);
    // The first weight panel needs nothing but its own barrier: it is requested before the CTA-wide set-up (TMEM
    // allocation, __syncthreads) completes, so that it streams in behind the rest of the prologue.
    auto issue_panel = [&](int panel) {
        if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(&b_full, (uint32_t)p.nsub * p.b_sub_bytes);
        __syncwarp();
        for (int s = 0; s < p.nsub; ++s) {
            const int co0 = panel_co0(p, panel, s);
            uint8_t* dst = smem + (size_t)s * p.b_sub_bytes;
            // one box per 64-channel chunk: (ci, co, tap) = (64, n_co, 9) lands as [tap][co][ci], i.e. the three
            // (kc, dy) blocks with rows [t][co] (tap = dy * 3 + t)
            for (int kc = 0; kc < p.kchunks; ++kc, dst += 9u * p.b_dx_bytes) {
                if (ptx::elect_one()) ptx::tma_load_3d(dst, &tmB, &b_full, kc * 64, co0, 0);
                __syncwarp();
            }
        }
    };
    if (warp == 0 && u_begin < u_end) issue_panel(wk.panel);
    if (threadIdx.x == 0) trace_stamp<TRACE>(p, 3, 63);          // first weight panel requested
    if (warp == 1) {
        if (lane == 0) trace_stamp<TRACE>(p, 4, 63);          // TMEM allocation starts
        ptx::tmem_alloc(&tmem_base_slot, 512);
        ptx::tmem_relinquish();
        if (lane == 0) trace_stamp<TRACE>(p, 5, 63);          // TMEM allocated
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    ptx::grid_launch_dependents();
    if (threadIdx.x == 0) trace_stamp<TRACE>(p, 1, 63);          // set-up done (barriers, TMEM)

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // Warp-uniform loop, one elected lane issues: a single-lane (`lane == 0`) loop makes the compiler wrap every UTMALDG
        // in a lane-serialisation loop (R2UR + ELECT + BRA.U.ANY), which costs the producer ~250 cycles per box.
        {
            // Each issuing warp owns its own ring of `ring` activation stages (stage = warp * ring + position): a parity
            // wait is only sound for a waiter that observes every phase of its barrier in order, and two warps taking
            // alternate fills of one ring can lap each other (tools/sim_dx_protocol.py reproduces the deadlock).
            uint32_t rpos = 0, rphase = 0, rpos_other = 0, rphase_other = 0, b_par = 0, rs = 0, rph = 0;    // [turn] / [other warp]
            int turn = 0;
            bool new_panel = true, first = true;
            for (int u = u_begin; u < u_end; ++u) {
                if (new_panel && !first) {                    // (the first panel was requested during the set-up)
                    for (int w = 0; w < p.mma_warps; ++w) ptx::mbar_wait(&b_empty[w], b_par);      // old panel fully consumed
                    b_par ^= 1;
                    issue_panel(wk.panel);
                }
                if (first) {                                  // weights do not depend on the previous kernel
                    ptx::grid_dependency_wait();
                    first = false;
                    if (lane == 0) trace_stamp<TRACE>(p, 7, 63);      // previous grid complete
                }
                const int ci0 = panel_ci0(p, wk.panel);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    const uint32_t stage = (uint32_t)(turn * p.ring) + rpos;
                    ptx::mbar_wait(&a_empty[stage], rphase ^ 1);
                    if (kc == 0 && lane == 0) trace_stamp<TRACE>(p, 0, (uint32_t)(u - u_begin) * p.nsub);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&a_full[stage], kABytes);
                        ptx::tma_load_4d(smem + p.off_a + (size_t)stage * kABytes, &tmA, &a_full[stage], ci0 + kc * 64, wk.w0 - 1,
                                         wk.h0 - 1, wk.b);
                    }
                    __syncwarp();
                    if (++rpos == (uint32_t)p.ring) { rpos = 0; rphase ^= 1; }
                }
                if (p.mma_warps == 2) {                   // next unit belongs to the other warp: swap the ring cursors
                    uint32_t t0 = rpos; rpos = rpos_other; rpos_other = t0;
                    t0 = rphase; rphase = rphase_other; rphase_other = t0;
                    turn ^= 1;
                }
                if (lane == 0) trace_stamp<TRACE>(p, 1, (uint32_t)(u - u_begin) * p.nsub);
                if (has_res) {
                    for (int s = 0; s < p.nsub; ++s) {
                        ptx::mbar_wait(&res_empty[rs], rph ^ 1);
                        if (ptx::elect_one()) {
                            ptx::mbar_arrive_expect_tx(&res_full[rs], p.res_bytes);
                            ptx::tma_load_4d(smem + p.off_res + (size_t)rs * p.res_stride, &tmR, &res_full[rs],
                                             panel_co0(p, wk.panel, s), wk.w0, wk.h0, wk.b);
                        }
                        __syncwarp();
                        if (++rs == (uint32_t)p.res_stages) { rs = 0; rph ^= 1; }
                    }
                }
                new_panel = wk.next(p);
            }
        }
    } else if (warp < kFirstEpiWarp) {
        // ------------------------------ UMMA issuers ------------------------------
        // Both warps walk the whole unit sequence (ring positions advance for every unit) and issue alternate units.
        const int me = warp - 1;
        const uint32_t ring_base = (uint32_t)(me * p.ring);
        const uint32_t idesc = ptx::make_idesc_bf16(128, ncols);
        const uint32_t b_base = ptx::smem_u32(smem), a_base = ptx::smem_u32(smem + p.off_a);
        const uint32_t b_blk16 = p.b_blk_bytes >> 4;
        const bool b128 = p.b_row_bytes == 128u;
        uint32_t stage = 0, phase = 0, buf = 0, aph = 0, item = 0, panel_idx = 0;
        int m = wk.m, turn = 0;
        bool have_panel = false, issued_in_panel = false;
        for (int u = u_begin; u < u_end; ++u) {
            const bool panel_ends = (u + 1 == u_end) || (m + 1 == p.m_tiles);
            // mbarrier waits are parity based: a waiter must observe EVERY phase of a barrier in order, or a phase it skipped
            // makes the next one look complete.  Each issuing warp therefore waits for every weight panel, also one in
            // which it happens to issue nothing; activation stages and accumulator buffers are never shared between the
            // two warps (own ring; nbuf / nsub even).
            if (me >= p.mma_warps) break;
            if (!have_panel) {
                ptx::mbar_wait(&b_full, panel_idx & 1);
                have_panel = true;
            }
            if (turn == me) {
                issued_in_panel = true;
                uint32_t st = stage, ph = phase, bf = buf, bph = aph;
                for (int s = 0; s < p.nsub; ++s) {
                    // probe the accumulator buffer first: the answer is consumed after the operand wait below
                    const bool acc_free = ptx::mbar_test_wait(&acc_empty[bf], bph ^ 1);
                    const uint32_t d_tmem = tmem_base + bf * (uint32_t)ncols;
                    st = stage; ph = phase;
                    uint32_t b_addr = b_base + (uint32_t)s * p.b_sub_bytes;
                    for (int kc = 0; kc < p.kchunks; ++kc, b_addr += 3u * p.b_blk_bytes) {
                        if (s == 0) ptx::mbar_wait(&a_full[ring_base + st], ph);
                        if (kc == 0) {
                            if (!acc_free) ptx::mbar_wait(&acc_empty[bf], bph ^ 1);
                            if (lane == 0) trace_stamp<TRACE>(p, 2, item + s);
                        }
                        ptx::tcgen05_fence_after();
                        if (kc == 0 && lane == 0) trace_stamp<TRACE>(p, 3, item + s);
                        if (ptx::elect_one()) {
                            const uint64_t b_desc0 = b128 ? ptx::make_kmajor_desc_sw128(b_addr, 1024) : ptx::make_kmajor_desc(b_addr, 64);
                            // sub-tile s of a two-group box reads channels [32 s, 32 s + 32) of the box: + 64 B = 4 units
                            const uint64_t a_desc0 = ptx::make_kmajor_desc_sw128(a_base + (ring_base + st) * kABytes, 1024) +
                                                     (uint64_t)(p.nsub == 2 ? 4 * s : 0);
                            const int nks = p.nsub == 2 ? 2 : ((kc == p.kchunks - 1) ? p.ks_last : 4);
                            const uint32_t acc_first = kc > 0 ? 1u : 0u;
                            if (nks == 4) issue_box<4>(a_desc0, b_desc0, b_blk16, d_tmem, idesc, acc_first);
                            else if (nks == 2) issue_box<2>(a_desc0, b_desc0, b_blk16, d_tmem, idesc, acc_first);
                            else if (nks == 1) issue_box<1>(a_desc0, b_desc0, b_blk16, d_tmem, idesc, acc_first);
                            else issue_box<3>(a_desc0, b_desc0, b_blk16, d_tmem, idesc, acc_first);
                            if (s == p.nsub - 1) ptx::umma_commit(&a_empty[ring_base + st]);      // box consumed by all its sub-tiles
                            if (kc == p.kchunks - 1) ptx::umma_commit(&acc_full[bf]);
                        }
                        __syncwarp();
                        if (++st == (uint32_t)p.ring) { st = 0; ph ^= 1; }
                    }
                    if (lane == 0) trace_stamp<TRACE>(p, 4, item + s);
                    if (++bf == (uint32_t)p.nbuf) { bf = 0; bph ^= 1; }
                }
            }
            if (turn == me)                                   // this warp's ring advances with its own units only
                for (int kc = 0; kc < p.kchunks; ++kc)
                    if (++stage == (uint32_t)p.ring) { stage = 0; phase ^= 1; }
            for (int s = 0; s < p.nsub; ++s)
                if (++buf == (uint32_t)p.nbuf) { buf = 0; aph ^= 1; }
            item += (uint32_t)p.nsub;
            if (++turn == p.mma_warps) turn = 0;
            if (panel_ends) {
                // the weight panel may be overwritten once BOTH warps' UMMAs that read it have completed
                if (issued_in_panel) { if (ptx::elect_one()) ptx::umma_commit(&b_empty[me]); }
                else if (lane == 0) ptx::mbar_arrive(&b_empty[me]);
                __syncwarp();
                have_panel = false; issued_in_panel = false; ++panel_idx;
            }
            if (++m == p.m_tiles) m = 0;
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        ptx::grid_dependency_wait();
        const int e = warp - kFirstEpiWarp;
        const int quad = warp & 3;                   // TMEM lane quadrant == image row of the tile
        const int group = p.ngroups == 4 ? (e >> 2) : (e >> 3);
        const int cofs = p.ngroups == 4 ? 0 : ((e >> 2) & 1) * kEpiCh;      // this warp's 32 channels within the sub-tile
        const bool two = p.epi2 != DD_EPI2_NONE;
        const uint32_t rb = (uint32_t)p.n_co * 2u;   // bytes per pixel row of a residual tile
        constexpr uint32_t kSlabRow = kEpiCh * 2;    // staging rows: 32 channels = 64 B (SWIZZLE_64B)
        // one 2 KB staging slab per warp; a second output goes through the same slab once the first one's TMA store has
        // read it (its 32 channels wait in registers): doubling the slabs cost the kernel two activation stages and with
        // them the second UMMA-issuing warp
        const uint32_t slab = ptx::smem_u32(smem) + p.off_slab + (uint32_t)e * p.slab_bytes;
        const int srow = min(max(lane - 1, 0), kOutW - 1);       // staging row of this lane (lanes 0 / 31 are halo)
        const bool out_lane = lane >= 1 && lane <= kOutW;
        const int rrow = quad * kOutW + srow;
        const float2 alpha2 = make_float2(p.alpha, p.alpha), beta2 = make_float2(p.beta, p.beta);
        const bool tracer = (e % (kEpiWarps / p.ngroups)) == 0;             // first warp of each epilogue group
        bool store_pending = false;
        uint32_t item = 0, buf = 0, aph = 0, rs = 0, rph = 0;
        int turn = 0;
        for (int u = u_begin; u < u_end; ++u) {
            for (int s = 0; s < p.nsub; ++s, ++item) {
                const uint32_t my_buf = buf, my_aph = aph, my_rs = rs, my_rph = rph;
                const bool mine = turn == group;
                if (++buf == (uint32_t)p.nbuf) { buf = 0; aph ^= 1; }
                if (++rs == (uint32_t)p.res_stages) { rs = 0; rph ^= 1; }
                if (++turn == p.ngroups) turn = 0;
                if (!mine) continue;
                const int h = wk.h0 + quad;
                const int co0 = panel_co0(p, wk.panel, s) + cofs;
                if (tracer && lane == 0) trace_stamp<TRACE>(p, 7, item);
                ptx::mbar_wait(&acc_full[my_buf], my_aph);
                ptx::tcgen05_fence_after();
                if (tracer && lane == 0) trace_stamp<TRACE>(p, 5, item);
                if (has_res) ptx::mbar_wait(&res_full[my_rs], my_rph);
                const bool raw_first = p.epi2 == DD_EPI2_RAW;      // pass 1 stages the accumulator itself, pass 2 the activation
                if (h < p.H) {
                    if (store_pending) {                    // the previous tile's TMA store must be done reading the slab
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
                    }
                    if (tracer && lane == 0) trace_stamp<TRACE>(p, 8, item);       // slab free again
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + my_buf * (uint32_t)ncols + (uint32_t)cofs;
                    const uint32_t res_tile = ptx::smem_u32(smem) + p.off_res + my_rs * p.res_stride;
#pragma unroll
                    for (int c0 = 0; c0 < kEpiCh; c0 += 16) {
                        uint32_t d0[16], d1[16], d2[16];
                        ptx::tmem_ld_32x16(taddr + c0, d0);
                        ptx::tmem_ld_32x16(taddr + p.n_co + c0, d1);
                        ptx::tmem_ld_32x16(taddr + 2 * p.n_co + c0, d2);
                        const int ch = co0 + c0;
                        float4 sc[4];
                        if (p.epi == DD_EPI_SCALE_SILU && !raw_first) {    // per-channel emb gain: in flight while the TMEM loads complete
                            const float4* sp = reinterpret_cast<const float4*>(p.scale + (size_t)wk.b * p.Cout + ch);
#pragma unroll
                            for (int i = 0; i < 4; ++i) sc[i] = __ldg(sp + i);
                        }
                        ptx::tmem_ld_wait();
                        float2 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 left = make_float2(__shfl_up_sync(0xffffffffu, __uint_as_float(d0[2 * i]), 1),
                                                            __shfl_up_sync(0xffffffffu, __uint_as_float(d0[2 * i + 1]), 1));
                            const float2 right = make_float2(__shfl_down_sync(0xffffffffu, __uint_as_float(d2[2 * i]), 1),
                                                             __shfl_down_sync(0xffffffffu, __uint_as_float(d2[2 * i + 1]), 1));
                            v[i] = add2(add2(make_float2(__uint_as_float(d1[2 * i]), __uint_as_float(d1[2 * i + 1])), left), right);
                        }
                        const int lc = c0 >> 3;              // first of this chunk's two 16 B pieces within the staging row
                        if (p.epi == DD_EPI_SCALE_SILU && !raw_first) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                v[2 * i] = mp_silu2(mul2(v[2 * i], make_float2(sc[i].x, sc[i].y)));
                                v[2 * i + 1] = mp_silu2(mul2(v[2 * i + 1], make_float2(sc[i].z, sc[i].w)));
                            }
                        } else if (has_res) {
                            const int rc = (cofs + c0) >> 3;
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const uint4 q = ld_shared_v4(res_tile + swz_off(rrow, rc + i, rb));
                                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 x = fma2(alpha2, v[4 * i + j], mul2(beta2, unpack_bf16x2(w4[j])));
                                    v[4 * i + j] = make_float2(clampf(x.x, p.clip), clampf(x.y, p.clip));
                                }
                            }
                        }
                        if (out_lane) {
                            st_shared_v4(slab + swz_off(srow, lc, kSlabRow), pack_bf16x8(v));
                            st_shared_v4(slab + swz_off(srow, lc + 1, kSlabRow), pack_bf16x8(v + 4));
                        }
                    }
                    if (tracer && lane == 0) trace_stamp<TRACE>(p, 9, item);       // arithmetic + staging done
                    ptx::fence_proxy_async_smem();          // generic-proxy slab writes -> visible to the TMA store
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(raw_first ? &tmO2 : &tmO, slab, co0, wk.w0, h, wk.b);
                        bulk_commit();
                    }
                }
                // all TMEM / residual reads of this warp are complete (wait::ld, ld.shared above): hand the buffers back
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&acc_empty[my_buf]);
                    if (has_res) ptx::mbar_arrive(&res_empty[my_rs]);
                }
                if (h < p.H) {
                    if (two) {
                        // Second output, derived from the first as it sits (bf16) in the slab once its TMA store has read it:
                        // mp_silu(out) | out * scale2 | (raw first) the activation mp_silu(acc * scale).  The reference
                        // applies these to bf16 tensors too (conv2d / mp_sum outputs under autocast), and neither extra
                        // staging memory nor registers held across the tile are needed.
                        if (lane == 0) bulk_wait_read0();
                        __syncwarp();
                        if (out_lane) {
                            const float* scp = raw_first ? p.scale : p.scale2;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t addr = slab + swz_off(srow, i, kSlabRow);
                                const uint4 q = ld_shared_v4(addr);
                                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
                                float2 x[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) x[j] = unpack_bf16x2(w4[j]);
                                if (p.epi2 != DD_EPI2_SILU && (p.epi2 == DD_EPI2_SCALE || p.epi == DD_EPI_SCALE_SILU)) {
                                    const float4* sp = reinterpret_cast<const float4*>(scp + (size_t)wk.b * p.Cout + co0 + 8 * i);
                                    const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
                                    x[0] = mul2(x[0], make_float2(s0.x, s0.y)); x[1] = mul2(x[1], make_float2(s0.z, s0.w));
                                    x[2] = mul2(x[2], make_float2(s1.x, s1.y)); x[3] = mul2(x[3], make_float2(s1.z, s1.w));
                                }
                                if (p.epi2 == DD_EPI2_SILU || (raw_first && p.epi == DD_EPI_SCALE_SILU)) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) x[j] = mp_silu2(x[j]);
                                }
                                st_shared_v4(addr, pack_bf16x8(x));
                            }
                        }
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d(raw_first ? &tmO : &tmO2, slab, co0, wk.w0, h, wk.b);
                            bulk_commit();
                        }
                    }
                    if (tracer && lane == 0) trace_stamp<TRACE>(p, 10, item);      // store(s) issued
                    store_pending = true;
                }
                if (tracer && lane == 0) trace_stamp<TRACE>(p, 6, item);
            }
            wk.next(p);
        }
        if (store_pending && lane == 0) bulk_wait0();
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) trace_stamp<TRACE>(p, 6, 63);          // about to exit
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC bf16 tensor [B][H][W][C] as a (C, W, H, B) map with box (bc, bw, bh, 1)
bool encode_nhwc(PFN_encodeTiled encode, CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bc, int bw, int bh,
                 CUtensorMapSwizzle swz) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int dd_launch_conv3x3_dx(const DxConvArgs& a, cudaStream_t stream) {
    static const bool disabled = getenv("DD_DISABLE_DX") != nullptr;          // tuning / A-B experiments only
    if (disabled) return -1;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    static const int min_h = getenv("DD_DX_MIN_H") ? atoi(getenv("DD_DX_MIN_H")) : 2;      // tuning: images shorter than a tile
    if (a.H < min_h || a.W < 16 || a.Cin < 64 || cin_g % 32 != 0 || cout_g % 32 != 0) return -1;
    if (a.epi != DD_EPI_NONE && a.epi != DD_EPI_SCALE_SILU && a.epi != DD_EPI_RESIDUAL) return -1;
    if (a.epi2 == DD_EPI2_RAW && a.epi == DD_EPI_RESIDUAL) return -1;      // (the raw-first second pass has no residual tile left)
    PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(dd_tensormap_encode_fn());
    DD_REQUIRE(encode != nullptr, "dd_mpconv_forward: cuTensorMapEncodeTiled unavailable (driver too old?)");

    DxParams p{};
    p.B = a.B; p.H = a.H; p.W = a.W; p.Cout = a.Cout;
    p.cin_g = cin_g; p.cout_g = cout_g;
    p.nsub = (cin_g == 32 && a.groups % 2 == 0) ? 2 : 1;
    p.kchunks = (cin_g + 63) / 64;
    p.ks_last = (cin_g - 64 * (p.kchunks - 1)) / 16;
    p.b_row_bytes = cin_g == 32 ? 64u : 128u;
    p.epi = a.epi; p.epi2 = a.epi2;
    p.alpha = a.alpha; p.beta = a.beta; p.clip = a.clip;
    p.scale = a.scale; p.scale2 = a.scale2;
    const bool has_res = a.epi == DD_EPI_RESIDUAL, two = a.epi2 != DD_EPI2_NONE;

    // output channels per sub-tile: 64 (N = 192, tensor-bound) when the weight panel leaves room for >= 3 activation
    // boxes in flight, else 32 (N = 96)
    static const int force_nco = getenv("DD_DX_NCO") ? atoi(getenv("DD_DX_NCO")) : 0;
    int chosen = 0;
    for (int n_co : {64, 32}) {
        if (cout_g % n_co != 0 || (force_nco && n_co != force_nco)) continue;
        // accumulator buffers must return to the same issuing warp every lap (parity waits, see the kernel): with two
        // sub-tiles per unit that needs the four buffers of n_co = 32
        if (n_co == 64 && p.nsub == 2 && kMmaWarps > 1) continue;
        const uint32_t b_total = (uint32_t)p.nsub * p.kchunks * 9u * n_co * p.b_row_bytes;
        const uint32_t slab = 32u * kEpiCh * 2u;        // per epilogue warp and output: 32 rows x 32 channels
        const uint32_t slabs = (uint32_t)kEpiWarps * slab;            // (a second output shares the slab)
        const uint32_t res_stride = (((uint32_t)(kTileH * kOutW * n_co * 2) + 1023u) / 1024u) * 1024u;
        const int res_stages = has_res ? (n_co == 64 ? 2 : 4) : 0;      // == epilogue groups: a group always meets the same stage
        const uint32_t fixed = b_total + slabs + res_stages * res_stride + 1024u;
        if (fixed + 2u * kABytes > kSmemLimit) continue;
        const int a_stages = std::min<int>(kMaxAStages, (int)((kSmemLimit - fixed) / kABytes));
        if (n_co == 64 && a_stages < 3 && cout_g % 32 == 0 && !force_nco) continue;       // prefer the narrower panel
        p.n_co = n_co; p.a_stages = a_stages; p.res_stages = std::max(1, res_stages);
        p.b_dx_bytes = (uint32_t)n_co * p.b_row_bytes;
        p.b_blk_bytes = 3u * p.b_dx_bytes;
        p.b_sub_bytes = (uint32_t)p.kchunks * 3u * p.b_blk_bytes;
        p.off_a = b_total;
        p.off_res = p.off_a + (uint32_t)a_stages * kABytes;
        p.res_stride = res_stride;
        p.res_bytes = (uint32_t)(kTileH * kOutW * n_co * 2);
        p.off_slab = p.off_res + (uint32_t)res_stages * res_stride;
        p.slab_bytes = slab;
        chosen = n_co;
        break;
    }
    if (!chosen) return -1;
    p.mma_warps = p.a_stages >= 4 ? kMmaWarps : 1;        // each issuing warp needs a ring of >= 2 activation stages
    p.ring = p.a_stages / p.mma_warps;
    p.nbuf = chosen == 64 ? 2 : 4;
    p.ngroups = chosen == 64 ? 2 : 4;
    p.npg = cout_g / p.n_co;
    p.tiles_w = (a.W + kOutW - 1) / kOutW;
    p.tiles_h = (a.H + kTileH - 1) / kTileH;
    p.m_tiles = p.tiles_w * p.tiles_h * a.B;
    const int panels = (a.groups / p.nsub) * p.npg;
    p.units = panels * p.m_tiles;
    p.fd_m = make_fastdiv(p.m_tiles); p.fd_npg = make_fastdiv(p.npg);
    p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);

    CUtensorMap tmA, tmB, tmO, tmO2, tmR;
    const CUtensorMapSwizzle res_swz = p.n_co == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    DD_REQUIRE(encode_nhwc(encode, &tmA, a.x, a.B, a.H, a.W, a.Cin, 64, kTileW, kTileH + 2, CU_TENSOR_MAP_SWIZZLE_128B),
               "dd_mpconv_forward: activation tensor map encode failed");
    {
        // weights [Cout][9][cin_g] viewed as (ci, co, tap): a box (64 | 32 channels, n_co, 9 taps) is one chunk of a panel
        cuuint64_t dims[3] = {(cuuint64_t)cin_g, (cuuint64_t)a.Cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)9 * cin_g * 2, (cuuint64_t)cin_g * 2};
        cuuint32_t box[3] = {p.b_row_bytes / 2, (cuuint32_t)p.n_co, 9};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a.w), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE,
                            p.b_row_bytes == 128u ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: weight tensor map encode failed (CUresult %d)", (int)r);
    }
    DD_REQUIRE(encode_nhwc(encode, &tmO, a.out, a.B, a.H, a.W, a.Cout, kEpiCh, kOutW, 1, CU_TENSOR_MAP_SWIZZLE_64B),
               "dd_mpconv_forward: output tensor map encode failed");
    DD_REQUIRE(encode_nhwc(encode, &tmO2, two ? a.out2 : a.out, a.B, a.H, a.W, a.Cout, kEpiCh, kOutW, 1, CU_TENSOR_MAP_SWIZZLE_64B),
               "dd_mpconv_forward: second output tensor map encode failed");
    DD_REQUIRE(encode_nhwc(encode, &tmR, has_res ? a.residual : a.out, a.B, a.H, a.W, a.Cout, p.n_co, kOutW, kTileH, res_swz),
               "dd_mpconv_forward: residual tensor map encode failed");

    const size_t smem_bytes = (size_t)p.off_slab + (size_t)kEpiWarps * p.slab_bytes + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_dx_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        DD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_dx_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
        attr_done = true;
    }
    if (getenv("DD_DEBUG_CONV"))
        fprintf(stderr, "[conv dx] B%d %dx%d %d->%d g%d: n_co %d nsub %d kchunks %d ks_last %d units %d a_stages %d res_stages %d smem %zu\n",
                a.B, a.H, a.W, a.Cin, a.Cout, a.groups, p.n_co, p.nsub, p.kchunks, p.ks_last, p.units, p.a_stages,
                has_res ? p.res_stages : 0, smem_bytes);
    const int grid = std::min(p.units, dd_num_sms());
    static const bool trace_on = getenv("DD_CONV_TRACE") != nullptr;       // diagnostic instantiation, tools/trace_halo.py
    if (trace_on) {
        p.trace = dd_conv_trace_buffer(true);
        DD_REQUIRE(p.trace != nullptr, "dd_mpconv_forward: could not allocate the trace buffer");
        DD_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, 4 * 12 * 64 * sizeof(unsigned long long), stream));
        int* meta = dd_conv_trace_meta();
        meta[0] = p.units * p.nsub; meta[1] = grid; meta[2] = p.n_co; meta[3] = p.ring * p.mma_warps; meta[4] = p.nbuf;
        meta[5] = kEpiWarps; meta[6] = p.kchunks; meta[7] = 2;
        DD_CHECK_CUDA(dd_launch_pdl(conv3x3_dx_kernel<true>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, tmO, tmO2, tmR, p));
        DD_CHECK_LAUNCH();
        return 0;
    }
    DD_CHECK_CUDA(dd_launch_pdl(conv3x3_dx_kernel<false>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, tmO, tmO2, tmR, p));
    DD_CHECK_LAUNCH();
    return 0;
}
