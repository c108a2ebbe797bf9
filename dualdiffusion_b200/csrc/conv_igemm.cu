// MPConv as an implicit-GEMM on the Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the reference's `F.conv2d(x, w, padding=k//2, groups=g)` in MPConv.forward
// (/root/reference/src/modules/mp_tools.py:369) for the stride-1 1x1 and 3x3 convolutions the
// EDM2 UNet uses, with the block-level elementwise work (emb-gain * mp_silu, mp_sum residual,
// clip; unet_edm2_b4.py:119-131,157) folded into the epilogue.
//
// Layout: activations NHWC bf16 ([B][H][W][C]); weights pre-scaled bf16 [Cout][tap][Cin/g]
// (written by dd_weight_prep).  GEMM view per group g:  D[pixels, Cout_g] = sum over taps of
// X_shifted[pixels, Cin_g] * W_tap[Cout_g, Cin_g]^T.
//
// One CTA owns a 128-row tile of output pixels (a wt x ht x bt box of the image, so that a
// filter tap is a pure coordinate shift of a 4-D TMA box and zero padding comes from TMA's
// out-of-bounds fill) and an n_tile-wide slice of one group's output channels.
//   warp 0     : TMA producer   (A: 4-D box of activations, B: 2-D box of weights)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x n_tile x 16)
//   warps 2..9 : epilogue       (tcgen05.ld -> registers -> fused elementwise -> global); two warps per
//                TMEM lane quadrant, interleaved over 16-column chunks
// A pipeline stage carries `sub` (1 or 2) consecutive (tap, k-chunk) operand pairs so that narrow tiles
// (n_tile <= 64 or 32-channel k-chunks) amortise the mbarrier round trip over more MMA work.
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1; the kernel is persistent over a static round-robin tile schedule.
#include "common.cuh"
#include "conv_dx.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kMaxAccBufs = 8;
constexpr int DD_EPI_HEAD = 3;   // internal: EDM output head (dd_conv_out)

// Division by a launch constant without the ~40-instruction integer divide (ncu: the per-tile tile decode was half of
// the epilogue warps' instruction stream): q = (n * ceil(2^44 / d)) >> 44, exact for n * d < 2^44 (tile counts < 2^20).
struct FastDiv {
    unsigned long long m;
    int d;
};
__host__ inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = d;
    f.m = ((1ull << 44) + (unsigned long long)d - 1ull) / (unsigned long long)d;
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return (int)(((unsigned long long)(unsigned)n * f.m) >> 44); }

struct ConvParams {
    int B, H, W, Cin, Cout;
    int kh, kw, taps;
    int cin_g, cout_g;
    int wt, ht, bt;                 // pixel box of one M tile (wt*ht*bt <= 128)
    int tiles_w, tiles_h, tiles_b, m_tiles;
    FastDiv fd_m_tiles, fd_npg, fd_tw, fd_th;       // divisions of the per-tile index decode
    int n_tile, n_tiles_per_group;
    int kchunks;                    // cin_g / KC
    int k_iters;                    // taps * kchunks
    int sub;                        // operand pairs per pipeline stage (1 or 2)
    int a_split;                    // > 0: input channels [a_split, Cin) are read from a second activation tensor
    int num_tiles;
    int stages;
    uint32_t a_bytes, b_bytes;      // bytes landed per stage by the two TMA boxes
    uint32_t tmem_cols;
    int nacc;                       // independent accumulators per tile (power of two), summed in the epilogue
    int nbuf;                       // TMEM accumulator buffers (power of two): MMAs run up to nbuf tiles ahead of the epilogue
    int dbg_taps;                   // experiment: number of taps actually issued (9 = all)
    int dbg_nostore;                // experiment: skip the epilogue's global stores
    int prefetch_res;               // epilogue warps request the next tile's residual rows into L2 a tile ahead
    uint32_t stage_off;             // byte offset (from the aligned dynamic smem base) of the epilogue warps' 4 KB staging
                                    // slabs; 0 = direct (un-staged) epilogue
    // halo kernel (3x3, weights stationary)
    int ks_last;                    // 16-channel MMA steps in the last 64-channel chunk
    int a_stages;
    uint32_t b_block_bytes;         // n_tile * 128
    uint32_t halo_bytes, halo_stride;   // landed bytes / stage stride of one halo tile
    // epilogue
    int epi;                        // DD_EPI_*
    int epi2;                       // DD_EPI2_*
    float alpha, beta, clip;
    const float* scale;             // [B][Cout]   (DD_EPI_SCALE_SILU)
    const float* scale2;            // [B][Cout]   (DD_EPI2_SCALE)
    const __nv_bfloat16* residual;  // [B][H][W][Cout]
    __nv_bfloat16* out;
    __nv_bfloat16* out2;
    // DD_EPI_HEAD (EDM output preconditioning, fp32 NCHW)
    int head_cout;
    float sigma_data;
    const float* sigma;             // [B]
    const float* x_in;              // [B][head_cout][H][W]
    const float* x_ref;             // [B][head_cout+1][H][W] or null
    float* d_out;                   // [B][head_cout][H][W]
    // role timeline of the halo kernel (DD_CONV_TRACE=1, conv3x3_halo_kernel<EW, true> only; appended last so that the
    // offsets of every other field -- and with them the default instantiations' code -- stay what they were)
    unsigned long long* trace;      // [kTraceCtas][kTraceSlots][kTraceTiles] clock64 stamps, or null
};

struct TileCoord {
    int g, n_idx, b0, h0, w0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
    TileCoord t;
    const int rest = fdiv(tile, p.fd_m_tiles);
    const int m_idx = tile - rest * p.m_tiles;
    t.g = fdiv(rest, p.fd_npg);
    t.n_idx = rest - t.g * p.n_tiles_per_group;
    const int r1 = fdiv(m_idx, p.fd_tw);                 // (b, h) tile index
    t.w0 = (m_idx - r1 * p.tiles_w) * p.wt;
    const int r2 = fdiv(r1, p.fd_th);                    // b tile index
    t.h0 = (r1 - r2 * p.tiles_h) * p.ht;
    t.b0 = r2 * p.bt;
    return t;
}


__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const float (&v)[16]) {
    uint4* op = reinterpret_cast<uint4*>(dst);
    op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                       pack_bf16x2(v[14], v[15]));
}

// Epilogue of one 128 x n_tile accumulator tile for one thread (= one output pixel row of the tile):
// TMEM -> registers -> fused elementwise (Block.forward glue) -> global.  `taddr` addresses this warp's lane
// quadrant and the tile's accumulator buffer; with EW == 8 the two warps of a quadrant interleave 16-column chunks.
template <int CH>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[CH], bool half) {
    if constexpr (CH == 16) {
        ptx::tmem_ld_32x16(taddr, r);
    } else {
        if (!half) ptx::tmem_ld_32x32(taddr, r);
        else ptx::tmem_ld_32x16(taddr, reinterpret_cast<uint32_t(&)[16]>(r));
    }
}

template <int EW>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, bool valid, int b, int h, int w,
                                              int ch0) {
    constexpr int kChunk = EW >= 8 ? 16 : 32;   // columns per tcgen05.ld
    const size_t pix = ((size_t)b * p.H + h) * p.W + w;
    uint32_t rn[kChunk];
    // software pipeline: the TMEM load of chunk i+1 is in flight while chunk i is converted and stored
    tmem_ld_chunk<kChunk>(taddr, rn, kChunk == 32 && 32 > p.n_tile);
    for (int c0 = 0; c0 < p.n_tile; c0 += kChunk) {
              uint32_t r[kChunk];
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < kChunk; ++i) r[i] = rn[i];
              // (tuning experiment DD_FORCE_NACC) k-steps round-robined over `nacc` accumulators are summed here
              for (int u = 1; u < p.nacc; ++u) {
                  uint32_t r2[kChunk];
                  tmem_ld_chunk<kChunk>(taddr + u * p.n_tile + c0, r2, kChunk == 32 && c0 + 32 > p.n_tile);
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < kChunk; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
              }
              if (c0 + kChunk < p.n_tile)
                  tmem_ld_chunk<kChunk>(taddr + c0 + kChunk, rn, kChunk == 32 && c0 + 2 * kChunk > p.n_tile);
              if (!valid) continue;
#pragma unroll
              for (int sub16 = 0; sub16 < kChunk / 16; ++sub16) {
                const int c = c0 + sub16 * 16;
                if (c >= p.n_tile) break;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[sub16 * 16 + i]);
                const int ch = ch0 + c;
                if (p.epi2 == DD_EPI2_RAW) store_bf16x16(p.out2 + pix * p.Cout + ch, v);   // pre-activation, for backward
                if (p.epi == DD_EPI_HEAD) {
                    // D = c_skip*x_in + c_out*F(x) (unet_edm2_b4.py:291), optional x_ref blend (:293-294)
                    const float sg = __ldg(p.sigma + b), sd2 = p.sigma_data * p.sigma_data;
                    const float c_skip = sd2 / (sg * sg + sd2), c_out = sg * p.sigma_data * rsqrtf(sg * sg + sd2);
                    const size_t plane = (size_t)p.H * p.W, hw = (size_t)h * p.W + w;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int co = ch + i;
                        if (co < p.head_cout) {
                            const size_t o = ((size_t)b * p.head_cout + co) * plane + hw;
                            float d = c_skip * __ldg(p.x_in + o) + c_out * v[i];
                            if (p.x_ref) {
                                const float rr = __ldg(p.x_ref + ((size_t)b * (p.head_cout + 1) + co) * plane + hw);
                                const float tt = __ldg(p.x_ref + ((size_t)b * (p.head_cout + 1) + p.head_cout) * plane + hw);
                                d = (rr + tt * (d - rr)) * rsqrtf((1.f - tt) * (1.f - tt) + tt * tt);
                            }
                            p.d_out[o] = d;
                        }
                    }
                    continue;
                }
                if (p.epi == DD_EPI_SCALE_SILU) {
                    const float4* sc = reinterpret_cast<const float4*>(p.scale + (size_t)b * p.Cout + ch);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 s = __ldg(sc + i);
                        v[4 * i + 0] = mp_silu_fast(v[4 * i + 0] * s.x);
                        v[4 * i + 1] = mp_silu_fast(v[4 * i + 1] * s.y);
                        v[4 * i + 2] = mp_silu_fast(v[4 * i + 2] * s.z);
                        v[4 * i + 3] = mp_silu_fast(v[4 * i + 3] * s.w);
                    }
                } else if (p.epi == DD_EPI_RESIDUAL) {
                    // (requesting these rows two chunks ahead measured no gain: the fused residual epilogue is bound by
                    // the scattered 32-byte sector traffic itself, not by its latency -- tools/exp_epi.py)
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + ch);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint4 q = __ldg(rp + i);
                        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = unpack_bf16x2(u[j]);
                            const float a0 = p.alpha * v[8 * i + 2 * j + 0] + p.beta * f.x;
                            const float a1 = p.alpha * v[8 * i + 2 * j + 1] + p.beta * f.y;
                            v[8 * i + 2 * j + 0] = fminf(fmaxf(a0, -p.clip), p.clip);
                            v[8 * i + 2 * j + 1] = fminf(fmaxf(a1, -p.clip), p.clip);
                        }
                    }
                }
                if (!p.dbg_nostore) store_bf16x16(p.out + pix * p.Cout + ch, v);
                if (p.epi2 != DD_EPI2_NONE && p.epi2 != DD_EPI2_RAW) {
                    if (p.epi2 == DD_EPI2_SILU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = mp_silu_fast(v[i]);
                    } else {
                        const float4* sc = reinterpret_cast<const float4*>(p.scale2 + (size_t)b * p.Cout + ch);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 s = __ldg(sc + i);
                            v[4 * i + 0] *= s.x; v[4 * i + 1] *= s.y; v[4 * i + 2] *= s.z; v[4 * i + 3] *= s.w;
                        }
                    }
                    store_bf16x16(p.out2 + pix * p.Cout + ch, v);
                }
              }
            }
}

// Lean single-output epilogues, one per mode (compile-time): accumulator -> fused glue -> bf16 -> global and nothing else.
// The general epilogue_tile re-tests its modes and re-derives its addresses for every 16 columns: 147 instructions per 32
// columns, a third of them register copies of its software pipeline (ncu source page of the level-0 1x1 layer: 71 % of the
// kernel's instruction stream), which made the epilogue warps the limiter.  Here two 16-column loads are in flight per
// iteration, the row's addresses are formed once, and a 32-column group costs ~40 instructions plus the mode's arithmetic.
template <int MODE>
__device__ __forceinline__ void lean_group16(const ConvParams& p, const uint32_t (&r)[16], __nv_bfloat16* dst,
                                             const float* sc, const __nv_bfloat16* res) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    if constexpr (MODE == DD_EPI_SCALE_SILU) {
        const float4* s4 = reinterpret_cast<const float4*>(sc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 q = __ldg(s4 + i);
            v[4 * i + 0] = mp_silu_fast(v[4 * i + 0] * q.x);
            v[4 * i + 1] = mp_silu_fast(v[4 * i + 1] * q.y);
            v[4 * i + 2] = mp_silu_fast(v[4 * i + 2] * q.z);
            v[4 * i + 3] = mp_silu_fast(v[4 * i + 3] * q.w);
        }
    } else if constexpr (MODE == DD_EPI_RESIDUAL) {
        const uint4* rp = reinterpret_cast<const uint4*>(res);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint4 q = __ldg(rp + i);
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(u[j]);
                const float a0 = p.alpha * v[8 * i + 2 * j + 0] + p.beta * f.x;
                const float a1 = p.alpha * v[8 * i + 2 * j + 1] + p.beta * f.y;
                v[8 * i + 2 * j + 0] = fminf(fmaxf(a0, -p.clip), p.clip);
                v[8 * i + 2 * j + 1] = fminf(fmaxf(a1, -p.clip), p.clip);
            }
        }
    }
    store_bf16x16(dst, v);
}
template <int MODE, bool PAIR>       // PAIR: two 16-column loads in flight (the 4-warp kernels have the registers for it)
__device__ __forceinline__ void epilogue_tile_lean(const ConvParams& p, uint32_t taddr, bool valid, int b, int h, int w,
                                                   int ch0) {
    const size_t pix = ((size_t)b * p.H + h) * p.W + w;
    __nv_bfloat16* dst = p.out + pix * p.Cout + ch0;
    const float* sc = MODE == DD_EPI_SCALE_SILU ? p.scale + (size_t)b * p.Cout + ch0 : nullptr;
    const __nv_bfloat16* res = MODE == DD_EPI_RESIDUAL ? p.residual + pix * p.Cout + ch0 : nullptr;
    const int n = p.n_tile;                       // a multiple of 16
    if constexpr (PAIR) {
        for (int c0 = 0; c0 < n; c0 += 32) {
            uint32_t r0[16], r1[16];
            const bool two = c0 + 16 < n;
            ptx::tmem_ld_32x16(taddr + c0, r0);
            if (two) ptx::tmem_ld_32x16(taddr + c0 + 16, r1);
            ptx::tmem_ld_wait();
            if (valid) {
                lean_group16<MODE>(p, r0, dst + c0, sc + c0, res + c0);
                if (two) lean_group16<MODE>(p, r1, dst + c0 + 16, sc + c0 + 16, res + c0 + 16);
            }
        }
    } else {
        for (int c0 = 0; c0 < n; c0 += 16) {
            uint32_t r0[16];
            ptx::tmem_ld_32x16(taddr + c0, r0);
            ptx::tmem_ld_wait();
            if (valid) lean_group16<MODE>(p, r0, dst + c0, sc + c0, res + c0);
        }
    }
}
// mode dispatch, once per tile; false = not a lean case (second output, head, experiments): the caller takes the general path
template <int EW>
__device__ __forceinline__ bool epilogue_tile_lean_any(const ConvParams& p, uint32_t taddr, bool valid, int b, int h, int w,
                                                       int ch0) {
    if (p.epi2 != DD_EPI2_NONE || p.nacc != 1 || p.dbg_nostore || p.stage_off != 0) return false;
    if (p.epi == DD_EPI_NONE) epilogue_tile_lean<DD_EPI_NONE, EW == 4>(p, taddr, valid, b, h, w, ch0);
    else if (p.epi == DD_EPI_SCALE_SILU) epilogue_tile_lean<DD_EPI_SCALE_SILU, EW == 4>(p, taddr, valid, b, h, w, ch0);
    else if (p.epi == DD_EPI_RESIDUAL) epilogue_tile_lean<DD_EPI_RESIDUAL, EW == 4>(p, taddr, valid, b, h, w, ch0);
    else return false;
    return true;
}

// The fused residual epilogues are bound by memory-level parallelism, not by bandwidth: a warp has only the 2-4 KB
// of residual it is about to consume in flight, and with ~1 us of loaded HBM latency 8-12 warps x 2-4 KB per SM is what
// Little's law gives for the measured ~2.5 TB/s (profiles/r01_ncu_epi_variants_summary.csv: DRAM 40 %, L2 20 %, tensor
// pipe 24 % busy).  Each epilogue thread therefore asks L2 for the residual row it will need for its NEXT tile while it
// waits for the current tile's accumulator, turning the later loads into L2 hits.
__device__ __forceinline__ void prefetch_residual_row(const ConvParams& p, int pixv, int ch0) {
    if (pixv < 0) return;
    const char* a = reinterpret_cast<const char*>(p.residual + (size_t)pixv * p.Cout + ch0);
    for (int off = 0; off < p.n_tile * 2; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + off));
}

// ---------------------------------------------------------------------------------
// Shared-memory-staged epilogue.  In the direct epilogue above a thread owns one pixel row of the tile, so every
// global access of a warp touches 32 different 128 B lines (16 B each): the LSU needs 32 passes per instruction and
// the fused residual / second-output epilogues become the limiter of the wide layers (tools/exp_epi.py).  Here each
// epilogue warp owns a private 4 KB slab (32 rows x 128 B, no cross-warp synchronisation): results are written to
// the slab in the row-per-thread layout and read back so that 8 (4, 2) adjacent lanes cover the 128 (64, 32)
// contiguous bytes of one pixel before they go to global memory; the residual takes the same route in reverse and is
// overwritten in place.  With a second output the slab holds 32 channels of each output (chunks 0-3 / 4-7 of a row),
// otherwise 64 channels.  The 16 B chunk index is XORed with f(row) = (row&1)<<2 | (row>>1)&3, which makes the
// row-per-thread accesses and the 8- and 4-lanes-per-row accesses bank-conflict free.
// Arithmetic is identical to epilogue_tile (bit-identical outputs: tools/dev_check_epi_staged.py).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t slab_addr(uint32_t slab, int row, int chunk) {
    const int f = ((row & 1) << 2) | ((row >> 1) & 3);
    return slab + (uint32_t)(row * 128 + ((chunk ^ f) << 4));
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* v) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// `pixv` = this thread's (row's) linear output pixel index, or -1 when the row lies outside the tensor; `slab` = shared
// address of this warp's 4 KB staging slab.  All 32 lanes must call (shuffles / __syncwarp inside).
template <int EW>
__device__ __forceinline__ void epilogue_tile_staged(const ConvParams& p, uint32_t taddr, int pixv, int b, int ch0,
                                                     uint32_t slab, int lane) {
    constexpr int kChunk = EW >= 8 ? 16 : 32;   // columns per tcgen05.ld
    const bool two = p.epi2 != DD_EPI2_NONE;
    uint32_t rn[kChunk];
    tmem_ld_chunk<kChunk>(taddr, rn, kChunk == 32 && 32 > p.n_tile);
    for (int slab0 = 0; slab0 < p.n_tile;) {
        const int rem = p.n_tile - slab0;
        const int lpr_log2 = (!two && rem >= 64) ? 3 : (rem >= 32 ? 2 : 1);    // lanes per pixel row: 8 / 4 / 2
        const int sw = 8 << lpr_log2;                                            // slab width in channels: 64 / 32 / 16
        const int lpr = 1 << lpr_log2, rstep = 32 >> lpr_log2;
        const int ck = lane & (lpr - 1), r_in = lane >> lpr_log2;
        const size_t col = (size_t)(ch0 + slab0 + ck * 8);
        if (p.epi == DD_EPI_RESIDUAL) {
            // coalesced residual rows -> slab (all loads in flight before the first shared store)
            for (int i0 = 0; i0 < lpr; i0 += 4) {
                uint4 q[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i0 + i < lpr) {
                        const int px = __shfl_sync(0xffffffffu, pixv, (i0 + i) * rstep + r_in);
                        q[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (px >= 0) q[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)px * p.Cout + col));
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i0 + i < lpr) st_shared_v4(slab_addr(slab, (i0 + i) * rstep + r_in, ck), q[i]);
            }
            __syncwarp();
        }
        for (int c0 = slab0; c0 < slab0 + sw; c0 += kChunk) {
            uint32_t r[kChunk];
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < kChunk; ++i) r[i] = rn[i];
            if (c0 + kChunk < p.n_tile)
                tmem_ld_chunk<kChunk>(taddr + c0 + kChunk, rn, kChunk == 32 && c0 + 2 * kChunk > p.n_tile);
#pragma unroll
            for (int sub16 = 0; sub16 < kChunk / 16; ++sub16) {
                const int c = c0 + sub16 * 16;
                if (c >= slab0 + sw) break;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[sub16 * 16 + i]);
                const int ch = ch0 + c;
                const int lc = (c - slab0) >> 3;                     // first of this sub-chunk's two 16 B chunks
                const uint32_t a0 = slab_addr(slab, lane, lc), a1 = slab_addr(slab, lane, lc + 1);
                const uint32_t a2 = slab_addr(slab, lane, lc + 4), a3 = slab_addr(slab, lane, lc + 5);
                if (p.epi2 == DD_EPI2_RAW) {                         // pre-activation, for backward
                    st_shared_v4(a2, pack_bf16x8(v));
                    st_shared_v4(a3, pack_bf16x8(v + 8));
                }
                if (p.epi == DD_EPI_SCALE_SILU) {
                    const float4* sc = reinterpret_cast<const float4*>(p.scale + (size_t)b * p.Cout + ch);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 s = __ldg(sc + i);
                        v[4 * i + 0] = mp_silu_fast(v[4 * i + 0] * s.x);
                        v[4 * i + 1] = mp_silu_fast(v[4 * i + 1] * s.y);
                        v[4 * i + 2] = mp_silu_fast(v[4 * i + 2] * s.z);
                        v[4 * i + 3] = mp_silu_fast(v[4 * i + 3] * s.w);
                    }
                } else if (p.epi == DD_EPI_RESIDUAL) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint4 q = ld_shared_v4(i == 0 ? a0 : a1);
                        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = unpack_bf16x2(u[j]);
                            const float x0 = p.alpha * v[8 * i + 2 * j + 0] + p.beta * f.x;
                            const float x1 = p.alpha * v[8 * i + 2 * j + 1] + p.beta * f.y;
                            v[8 * i + 2 * j + 0] = fminf(fmaxf(x0, -p.clip), p.clip);
                            v[8 * i + 2 * j + 1] = fminf(fmaxf(x1, -p.clip), p.clip);
                        }
                    }
                }
                st_shared_v4(a0, pack_bf16x8(v));
                st_shared_v4(a1, pack_bf16x8(v + 8));
                if (p.epi2 == DD_EPI2_SILU || p.epi2 == DD_EPI2_SCALE) {
                    if (p.epi2 == DD_EPI2_SILU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = mp_silu_fast(v[i]);
                    } else {
                        const float4* sc = reinterpret_cast<const float4*>(p.scale2 + (size_t)b * p.Cout + ch);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 s = __ldg(sc + i);
                            v[4 * i + 0] *= s.x; v[4 * i + 1] *= s.y; v[4 * i + 2] *= s.z; v[4 * i + 3] *= s.w;
                        }
                    }
                    st_shared_v4(a2, pack_bf16x8(v));
                    st_shared_v4(a3, pack_bf16x8(v + 8));
                }
            }
        }
        __syncwarp();
        // slab -> global: `lpr` adjacent lanes write the contiguous bytes of one pixel
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < lpr) {
                const int rl = i * rstep + r_in;
                const int px = __shfl_sync(0xffffffffu, pixv, rl);
                const uint4 o = ld_shared_v4(slab_addr(slab, rl, ck));
                uint4 o2 = o;
                if (two) o2 = ld_shared_v4(slab_addr(slab, rl, ck + 4));
                if (px >= 0) {
                    *reinterpret_cast<uint4*>(p.out + (size_t)px * p.Cout + col) = o;
                    if (two) *reinterpret_cast<uint4*>(p.out2 + (size_t)px * p.Cout + col) = o2;
                }
            }
        }
        __syncwarp();
        slab0 += sw;
    }
}

// EW = number of epilogue warps: 4 (one per TMEM lane quadrant, 32-column chunks) or 8 (two per quadrant,
// interleaved 16-column chunks; for epilogues with transcendental math).
template <int KC, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvParams p) {
    constexpr uint32_t kRowBytes = KC * 2;           // one swizzle span per row
    constexpr uint32_t kABufBytes = kTileM * kRowBytes;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[kMaxAccBufs];
    __shared__ __align__(8) uint64_t tmem_empty_bar[kMaxAccBufs];
    __shared__ uint32_t tmem_base_slot;

    // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B swizzle atom
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_buf_bytes = (uint32_t)p.n_tile * kRowBytes;
    const uint32_t pair_bytes = kABufBytes + b_buf_bytes;          // one (A box, B box) operand pair
    const uint32_t stage_bytes = pair_bytes * (uint32_t)p.sub;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;

    if (warp == 0) {            // one barrier pair per lane: the set-up is on every launch's critical path
        if (lane == 0) {
            ptx::prefetch_tensormap(&tmA);
            ptx::prefetch_tensormap(&tmB);
        }
        if (lane == 1 && p.a_split > 0) ptx::prefetch_tensormap(&tmA2);
        if (lane < p.stages) {
            ptx::mbar_init(&full_bar[lane], 1);
            ptx::mbar_init(&empty_bar[lane], 1);
        }
        if (lane < p.nbuf) {
            ptx::mbar_init(&tmem_full_bar[lane], 1);
            ptx::mbar_init(&tmem_empty_bar[lane], 4);    // the four warps (lane quadrants) of one epilogue group
        }
        ptx::mbar_fence_init();
        ptx::fence_proxy_async_smem();
        __syncwarp();
    }
    // Weights never depend on the preceding kernel and need nothing but their own barriers: the B boxes of the first
    // pipeline fill are requested here, before the CTA-wide set-up (TMEM allocation, __syncthreads) completes and before
    // griddepcontrol.wait, so they stream in behind the prologue / while the previous kernel drains.
    struct KIter {
        int tap, kc, dy, dx;
        __device__ __forceinline__ void reset(const ConvParams& q) { tap = 0; kc = 0; dy = -(q.kh / 2); dx = -(q.kw / 2); }
        __device__ __forceinline__ void next(const ConvParams& q) {
            if (++kc == q.kchunks) {
                kc = 0; ++tap;
                if (++dx > q.kw / 2) { dx = -(q.kw / 2); ++dy; }
            }
        }
    };
    int prefetched = 0;                          // stages of the first tile whose B boxes are already in flight
    if (warp == 0) {
        const int tile = blockIdx.x;
        if (tile < p.num_tiles) {
            const TileCoord t = decode_tile(p, tile);
            const int b_row = t.g * p.cout_g + t.n_idx * p.n_tile;
            KIter ki;
            ki.reset(p);
            for (int it0 = 0; it0 < p.k_iters && prefetched < p.stages; it0 += p.sub, ++prefetched) {
                const int cnt = min(p.sub, p.k_iters - it0);
                if (ptx::elect_one())
                    ptx::mbar_arrive_expect_tx(&full_bar[prefetched], (uint32_t)cnt * (p.a_bytes + p.b_bytes));
                __syncwarp();
                for (int j = 0; j < cnt; ++j) {
                    if (ptx::elect_one())
                        ptx::tma_load_2d(smem + prefetched * stage_bytes + j * pair_bytes + kABufBytes, &tmB,
                                         &full_bar[prefetched], ki.tap * p.cin_g + ki.kc * KC, b_row);
                    __syncwarp();
                    ki.next(p);
                }
            }
        }
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_slot, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    ptx::grid_launch_dependents();      // PDL: the next kernel may start its own prologue / weight prefetch now

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // The whole warp runs the (warp-uniform) loop and one elected lane issues.  With a single-lane (`lane == 0`) loop
        // the compiler cannot keep the TMA operands in uniform registers: it wraps every UTMALDG in a lane-serialisation
        // loop (R2UR + ELECT + BRA.U.ANY, ~250 cycles per box), which made the producer -- not L2 -- the limiter of the
        // short k-loops of the small levels.  (tap, k-chunk) advance incrementally: no integer division per box.
        // Activations (A boxes) are only touched after griddepcontrol.wait; the first B boxes are already in flight.
        uint32_t stage = 0, phase = 0;
        ptx::grid_dependency_wait();
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            const int a_c0 = t.g * p.cin_g;
            const int b_row = t.g * p.cout_g + t.n_idx * p.n_tile;
            KIter ki;
            ki.reset(p);
            for (int it0 = 0; it0 < p.k_iters; it0 += p.sub) {
                const int cnt = min(p.sub, p.k_iters - it0);
                const bool b_done = prefetched > 0;              // this stage's barrier is armed, B already issued
                if (!b_done) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (ptx::elect_one())
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)cnt * (p.a_bytes + p.b_bytes));
                    __syncwarp();
                } else {
                    --prefetched;
                }
                for (int j = 0; j < cnt; ++j) {
                    uint8_t* a_dst = smem + stage * stage_bytes + j * pair_bytes;
                    if (ptx::elect_one()) {
                        // (dd_mpconv_forward_cat: input channels >= a_split come from the second tensor -- the conv reads the
                        // two operands of an mp_cat without the concatenation ever being written)
                        const int c = a_c0 + ki.kc * KC;
                        const bool second = p.a_split > 0 && c >= p.a_split;
                        ptx::tma_load_4d(a_dst, second ? &tmA2 : &tmA, &full_bar[stage], second ? c - p.a_split : c, t.w0 + ki.dx,
                                         t.h0 + ki.dy, t.b0);
                        if (!b_done)
                            ptx::tma_load_2d(a_dst + kABufBytes, &tmB, &full_bar[stage], ki.tap * p.cin_g + ki.kc * KC, b_row);
                    }
                    __syncwarp();
                    ki.next(p);
                }
                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        // The whole warp runs the (warp-uniform) control flow so that addresses and descriptors live in
        // uniform registers; one elected lane issues the tcgen05 instructions.
        {
            const uint32_t idesc = ptx::make_idesc_bf16(kTileM, p.n_tile);
            const uint32_t smem_base = ptx::smem_u32(smem);
            uint32_t stage = 0, phase = 0;
            uint32_t local = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
                const uint32_t acc = local & (uint32_t)(p.nbuf - 1);
                const uint32_t acc_phase = (local / (uint32_t)p.nbuf) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                ptx::tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.n_tile;
                for (int it0 = 0; it0 < p.k_iters; it0 += p.sub) {
                    const int cnt = min(p.sub, p.k_iters - it0);
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tcgen05_fence_after();
                    if (ptx::elect_one()) {
                        const uint32_t s_addr = smem_base + stage * stage_bytes;
                        for (int j = 0; j < cnt; ++j) {
                            const uint64_t a_desc = ptx::make_kmajor_desc(s_addr + j * pair_bytes, kRowBytes);
                            const uint64_t b_desc = ptx::make_kmajor_desc(s_addr + j * pair_bytes + kABufBytes, kRowBytes);
                            // only the first UMMA of a tile overwrites the accumulator; +32 B per 16-channel step = +2 units
                            ptx::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (it0 + j) > 0 ? 1u : 0u);
#pragma unroll
                            for (int ks = 1; ks < KC / 16; ++ks)
                                ptx::umma_bf16_ss_acc(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc);
                        }
                        ptx::umma_commit(&empty_bar[stage]);    // frees the smem stage once the MMAs retire
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one()) ptx::umma_commit(&tmem_full_bar[acc]);   // accumulator complete -> epilogue
                __syncwarp();
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        ptx::grid_dependency_wait();                // epilogue reads (residual, scales) / writes depend on prior kernels
        // EW == 8: two groups of four warps (one per TMEM lane quadrant) take alternate tiles
        const int group = (warp - 2) >> 2, ngroups = EW / 4;
        const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;           // row of the 128-row tile == box-linear pixel index
        const int ww = row % p.wt;
        const int hh = (row / p.wt) % p.ht;
        const int bb = row / (p.wt * p.ht);
        uint32_t local = group;
        for (int tile = blockIdx.x + group * gridDim.x; tile < p.num_tiles; tile += ngroups * gridDim.x, local += ngroups) {
            const TileCoord t = decode_tile(p, tile);
            const uint32_t acc = local & (uint32_t)(p.nbuf - 1);
            const uint32_t acc_phase = (local / (uint32_t)p.nbuf) & 1;
            const int b = t.b0 + bb, h = t.h0 + hh, w = t.w0 + ww;
            const bool valid = (bb < p.bt) && (b < p.B) && (h < p.H) && (w < p.W);
            const int ch0 = t.g * p.cout_g + t.n_idx * p.n_tile;
            if (p.prefetch_res && tile + ngroups * (int)gridDim.x < p.num_tiles) {
                const TileCoord tn = decode_tile(p, tile + ngroups * (int)gridDim.x);
                const int bn = tn.b0 + bb, hn = tn.h0 + hh, wn = tn.w0 + ww;
                const bool vn = (bb < p.bt) && (bn < p.B) && (hn < p.H) && (wn < p.W);
                prefetch_residual_row(p, vn ? (bn * p.H + hn) * p.W + wn : -1, tn.g * p.cout_g + tn.n_idx * p.n_tile);
            }

            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)(p.n_tile * p.nacc);
            if (epilogue_tile_lean_any<EW>(p, taddr, valid, b, h, w, ch0)) {
            } else if (p.stage_off) {
                const uint32_t slab = ptx::smem_u32(smem) + p.stage_off + (uint32_t)(warp - 2) * 4096u;
                epilogue_tile_staged<EW>(p, taddr, valid ? (b * p.H + h) * p.W + w : -1, valid ? b : 0, ch0, slab, lane);
            } else {
                epilogue_tile<EW>(p, taddr, valid, b, h, w, ch0);
            }
            // all TMEM reads of this buffer have completed (wait::ld above): hand it back
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}


// ---------------------------------------------------------------------------------
// 3x3 convolution, halo variant (image rows >= 16): the 9 filter taps are shifted *views* of one
// shared-memory halo tile instead of 9 separate TMA boxes, and the weight panel stays resident.
//
// M tile = 16 groups of 8 consecutive pixels: 8 x 16 pixels of one image, or (images of 8..15 rows) 8 x 8
// pixels of two batch items with the image rows of the two items interleaved.  Per 64-channel chunk one TMA
// box of (8+2) x bt x (ht+2) pixels x 64 channels (23-26 KB, 128B-swizzled rows, smem order w, b, h) is loaded.
// For tap (dy,dx) the A operand of UMMA is that same tile read from row dy*bt*10+dx with a stride of 10 rows
// between 8-row groups -- operand traffic from L2 drops ~6x versus re-loading a box per tap.  The B operand
// (9 taps x all chunks x n_tile output channels of one group) is loaded once per weight panel; every CTA walks
// a contiguous range of (panel, m-tile) pairs so that it reloads weights at most once or twice per launch.
// ---------------------------------------------------------------------------------
constexpr int kHaloW = 8, kHaloPitch = kHaloW + 2;
constexpr int kMaxAStages = 6;

struct HaloTile {
    int panel, g, n_idx, b, h0, w0;
};

__device__ __forceinline__ HaloTile decode_halo_tile(const ConvParams& p, int tile) {
    HaloTile t;
    t.panel = fdiv(tile, p.fd_m_tiles);
    const int m = tile - t.panel * p.m_tiles;
    t.g = fdiv(t.panel, p.fd_npg);
    t.n_idx = t.panel - t.g * p.n_tiles_per_group;
    const int r1 = fdiv(m, p.fd_tw);
    t.w0 = (m - r1 * p.tiles_w) * kHaloW;
    const int r2 = fdiv(r1, p.fd_th);
    t.h0 = (r1 - r2 * p.tiles_h) * p.ht;
    t.b = r2 * p.bt;
    return t;
}

// Issue the 9 taps x NKS 16-channel steps of one 64-channel chunk.  Descriptor arithmetic is in 16 B units:
// a tap is a (dy*RP + dx)-row shift of the halo tile (8 units per 128 B row), a k-step is +2 units.  The issuing
// warp's instruction stream is the limiter for narrow tiles (ncu: ~10 uniform-datapath instructions per UTCHMMA made a
// 36-UMMA chunk cost ~3500 cycles against ~1500 of tensor-pipe time), so everything here is an immediate: the row
// pitch RP is a template parameter, the accumulate flag is constant except for the very first UMMA of a tile, and each
// descriptor is one 64-bit add away from the chunk's base.
template <int NKS, int RP>
__device__ __forceinline__ void halo_issue_taps(uint64_t a_desc0, uint64_t b_desc0, uint32_t b_block16, uint32_t d_tmem,
                                                uint32_t idesc, uint32_t acc_first) {
    uint64_t b_tap = b_desc0;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const uint64_t a_tap = a_desc0 + (uint64_t)(((tap / 3) * RP + (tap % 3)) * 8);
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            if (tap == 0 && ks == 0) ptx::umma_bf16_ss(d_tmem, a_tap, b_tap, idesc, acc_first);
            else ptx::umma_bf16_ss_acc(d_tmem, a_tap + 2 * ks, b_tap + 2 * ks, idesc);
        }
        b_tap += b_block16;
    }
}

template <int RP>
__device__ __forceinline__ void halo_issue_chunk(int nks, uint64_t a_desc0, uint64_t b_desc0, uint32_t b_block16,
                                                 uint32_t d_tmem, uint32_t idesc, uint32_t acc_first) {
    if (nks == 4) halo_issue_taps<4, RP>(a_desc0, b_desc0, b_block16, d_tmem, idesc, acc_first);
    else if (nks == 2) halo_issue_taps<2, RP>(a_desc0, b_desc0, b_block16, d_tmem, idesc, acc_first);
    else if (nks == 1) halo_issue_taps<1, RP>(a_desc0, b_desc0, b_block16, d_tmem, idesc, acc_first);
    else halo_issue_taps<3, RP>(a_desc0, b_desc0, b_block16, d_tmem, idesc, acc_first);
}

// ---- role timeline (diagnostic instantiation only) ----
// Per-tile clock64 stamps of the three roles of the first kTraceCtas CTAs.  All roles of a CTA run on one SM, so the
// stamps are directly comparable; tools/trace_halo.py turns them into wait / busy intervals per role and tile.
constexpr int kTraceCtas = 4, kTraceSlots = 12, kTraceTiles = 64;
enum TraceSlot { kTrTmaFree = 0,      // producer: stage free (a_empty wait done), about to issue the tile's first box
                 kTrTmaIssued = 1,    // producer: last box of the tile issued
                 kTrMmaAcc = 2,       // MMA warp: accumulator buffer free (tmem_empty wait done)
                 kTrMmaOperands = 3,  // MMA warp: the tile's first halo box has landed (a_full wait done)
                 kTrMmaIssued = 4,    // MMA warp: last UMMA + commit issued
                 kTrEpiFull = 5,      // epilogue (quadrant-0 warp of the tile's group): accumulator complete
                 kTrEpiDone = 6,      // epilogue: tile stored, accumulator released
                 kTrEpiStart = 7 };   // epilogue: began waiting for this tile
template <bool TRACE>
__device__ __forceinline__ void trace_stamp(const ConvParams& p, int slot, uint32_t local) {
    if constexpr (TRACE) {
        if (blockIdx.x < (unsigned)kTraceCtas && local < (uint32_t)kTraceTiles && p.trace != nullptr)
            p.trace[((size_t)blockIdx.x * kTraceSlots + slot) * kTraceTiles + local] = (unsigned long long)clock64();
    }
}

template <int EW, bool TRACE = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[kMaxAStages];
    __shared__ __align__(8) uint64_t a_empty[kMaxAStages];
    __shared__ __align__(8) uint64_t b_full, b_empty;
    __shared__ __align__(8) uint64_t tmem_full_bar[kMaxAccBufs];
    __shared__ __align__(8) uint64_t tmem_empty_bar[kMaxAccBufs];
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_smem = smem;                                              // [kchunks][9] blocks of n_tile x 128 B
    uint8_t* a_smem = smem + (size_t)p.kchunks * 9 * p.b_block_bytes;    // a_stages x kHaloStride

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int t_begin = (int)((long)blockIdx.x * p.num_tiles / gridDim.x);
    const int t_end = (int)((long)(blockIdx.x + 1) * p.num_tiles / gridDim.x);

    if (warp == 0) {            // one barrier pair per lane: the set-up is on every launch's critical path
        if (lane == 0) {
            ptx::prefetch_tensormap(&tmA);
            ptx::prefetch_tensormap(&tmB);
            ptx::mbar_init(&b_full, 1);
            ptx::mbar_init(&b_empty, 1);
        }
        if (lane < p.a_stages) {
            ptx::mbar_init(&a_full[lane], 1);
            ptx::mbar_init(&a_empty[lane], 1);
        }
        if (lane >= 16 && lane - 16 < p.nbuf) {
            ptx::mbar_init(&tmem_full_bar[lane - 16], 1);
            ptx::mbar_init(&tmem_empty_bar[lane - 16], 4);       // the four warps (lane quadrants) of one epilogue group
        }
        ptx::mbar_fence_init();
        ptx::fence_proxy_async_smem();
        __syncwarp();
    }
    // one box per 64-channel chunk of the weight panel: (ci, co, tap) = (64, n_tile, 9) lands as [tap][co][ci]
    auto issue_panel = [&](const HaloTile& t) {
        if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(&b_full, (uint32_t)p.kchunks * 9u * p.b_block_bytes);
        __syncwarp();
        const int b_row = t.g * p.cout_g + t.n_idx * p.n_tile;
        uint8_t* dst = b_smem;
        for (int kc = 0; kc < p.kchunks; ++kc, dst += 9u * p.b_block_bytes) {
            if (ptx::elect_one()) ptx::tma_load_3d(dst, &tmB, &b_full, kc * 64, b_row, 0);
            __syncwarp();
        }
    };
    // the first panel needs nothing but its own barrier: requested before the CTA-wide set-up completes
    if (warp == 0 && t_begin < t_end) issue_panel(decode_halo_tile(p, t_begin));
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_slot, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    ptx::grid_launch_dependents();      // PDL: the next kernel may start its own prologue / weight prefetch now

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // warp-uniform loop, one elected lane issues (see conv_igemm_kernel: a single-lane loop costs a lane-serialisation
        // loop around every UTMALDG)
        {
            uint32_t stage = 0, phase = 0, b_par = 0;
            int cur_panel = -1;
            bool waited = false;
            for (int tile = t_begin; tile < t_end; ++tile) {
                const uint32_t tr_local = (uint32_t)(tile - t_begin);
                const HaloTile t = decode_halo_tile(p, tile);
                if (t.panel != cur_panel) {
                    if (cur_panel >= 0) {                      // (the first panel was requested during the set-up)
                        ptx::mbar_wait(&b_empty, b_par);       // old panel fully consumed
                        b_par ^= 1;
                        issue_panel(t);
                    }
                    cur_panel = t.panel;
                }
                // the weight panel does not depend on the previous kernel; activations do (PDL)
                if (!waited) { ptx::grid_dependency_wait(); waited = true; }
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    ptx::mbar_wait(&a_empty[stage], phase ^ 1);
                    if (kc == 0 && lane == 0) trace_stamp<TRACE>(p, kTrTmaFree, tr_local);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&a_full[stage], p.halo_bytes);
                        ptx::tma_load_4d(a_smem + (size_t)stage * p.halo_stride, &tmA, &a_full[stage], t.g * p.cin_g + kc * 64,
                                         t.w0 - 1, t.b, t.h0 - 1);      // tensor-map dims are (C, W, B, H)
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)p.a_stages) { stage = 0; phase ^= 1; }
                }
                if (lane == 0) trace_stamp<TRACE>(p, kTrTmaIssued, tr_local);
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        {
            const uint32_t idesc = ptx::make_idesc_bf16(kTileM, p.n_tile);
            const uint32_t b_base = ptx::smem_u32(b_smem), a_base = ptx::smem_u32(a_smem);
            const uint32_t b_block16 = p.b_block_bytes >> 4;
            uint32_t stage = 0, phase = 0, b_par = 0, local = 0;
            int cur_panel = -1;
            for (int tile = t_begin; tile < t_end; ++tile, ++local) {
                const int panel = fdiv(tile, p.fd_m_tiles);
                if (panel != cur_panel) {
                    ptx::mbar_wait(&b_full, b_par);
                    b_par ^= 1;
                    cur_panel = panel;
                }
                const uint32_t acc = local & (uint32_t)(p.nbuf - 1);
                const uint32_t acc_phase = (local / (uint32_t)p.nbuf) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                ptx::tcgen05_fence_after();
                if (lane == 0) trace_stamp<TRACE>(p, kTrMmaAcc, local);
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.n_tile;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    ptx::mbar_wait(&a_full[stage], phase);
                    ptx::tcgen05_fence_after();
                    if (kc == 0 && lane == 0) trace_stamp<TRACE>(p, kTrMmaOperands, local);
                    if (ptx::elect_one()) {
                        const uint64_t a_desc0 = ptx::make_kmajor_desc_sw128(a_base + stage * p.halo_stride, kHaloPitch * 128);
                        const uint64_t b_desc0 = ptx::make_kmajor_desc_sw128(b_base + (uint32_t)(kc * 9) * p.b_block_bytes, 1024);
                        const int nks = (kc == p.kchunks - 1) ? p.ks_last : 4;
                        const uint32_t acc_first = kc > 0 ? 1u : 0u;
                        if (p.bt == 1) halo_issue_chunk<kHaloPitch>(nks, a_desc0, b_desc0, b_block16, d_tmem, idesc, acc_first);
                        else halo_issue_chunk<2 * kHaloPitch>(nks, a_desc0, b_desc0, b_block16, d_tmem, idesc, acc_first);
                        ptx::umma_commit(&a_empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)p.a_stages) { stage = 0; phase ^= 1; }
                }
                const bool panel_ends = (tile + 1 == t_end) || (fdiv(tile + 1, p.fd_m_tiles) != panel);
                if (ptx::elect_one()) {
                    ptx::umma_commit(&tmem_full_bar[acc]);
                    if (panel_ends) ptx::umma_commit(&b_empty);
                }
                __syncwarp();
                if (lane == 0) trace_stamp<TRACE>(p, kTrMmaIssued, local);
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        ptx::grid_dependency_wait();
        const int group = (warp - 2) >> 2, ngroups = EW / 4;      // EW == 8: two warp groups take alternate tiles
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int grp = row >> 3, ww = row & 7;       // group index = h * bt + b
        const int hh = grp / p.bt, bb = grp - hh * p.bt;
        uint32_t local = group;
        for (int tile = t_begin + group; tile < t_end; tile += ngroups, local += ngroups) {
            const HaloTile t = decode_halo_tile(p, tile);
            const uint32_t acc = local & (uint32_t)(p.nbuf - 1);
            const uint32_t acc_phase = (local / (uint32_t)p.nbuf) & 1;
            const int h = t.h0 + hh, w = t.w0 + ww, b = t.b + bb;
            const bool valid = (hh < p.ht) && (h < p.H) && (w < p.W) && (b < p.B);
            const int ch0 = t.g * p.cout_g + t.n_idx * p.n_tile;
            if (p.prefetch_res && tile + ngroups < t_end) {
                const HaloTile tn = decode_halo_tile(p, tile + ngroups);
                const int hn = tn.h0 + hh, wn = tn.w0 + ww, bn = tn.b + bb;
                const bool vn = (hh < p.ht) && (hn < p.H) && (wn < p.W) && (bn < p.B);
                prefetch_residual_row(p, vn ? (bn * p.H + hn) * p.W + wn : -1, tn.g * p.cout_g + tn.n_idx * p.n_tile);
            }
            if (quad == 0 && lane == 0) trace_stamp<TRACE>(p, kTrEpiStart, local);
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            if (quad == 0 && lane == 0) trace_stamp<TRACE>(p, kTrEpiFull, local);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)(p.n_tile * p.nacc);
            if (epilogue_tile_lean_any<EW>(p, taddr, valid, b, h, w, ch0)) {
            } else if (p.stage_off) {
                const uint32_t slab = ptx::smem_u32(smem) + p.stage_off + (uint32_t)(warp - 2) * 4096u;
                epilogue_tile_staged<EW>(p, taddr, valid ? (b * p.H + h) * p.W + w : -1, valid ? b : 0, ch0, slab, lane);
            } else {
                epilogue_tile<EW>(p, taddr, valid, b, h, w, ch0);
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
            if (quad == 0 && lane == 0) trace_stamp<TRACE>(p, kTrEpiDone, local);
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<PFN_encodeTiled>(sym);
        }
    }
    return fn;
}

// Launch with programmatic stream serialization (PDL): the kernel may begin before its stream predecessor has
// finished; everything that depends on the predecessor sits behind griddepcontrol.wait inside the kernels.
template <typename Kernel, typename... Args>
cudaError_t launch_pdl(Kernel kernel, int grid, int block, size_t smem, cudaStream_t stream, const Args&... args) {
    static const bool no_pdl = getenv("DD_DISABLE_PDL") != nullptr;      // tuning experiments only
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Pick the pixel box of an M tile: wt*ht*bt <= 128, minimising the number of tiles.
void choose_box(int B, int H, int W, int& wt, int& ht, int& bt) {
    long best_tiles = -1;
    int best_rows = 0;
    wt = ht = bt = 1;
    for (int w = 1; w <= std::min(W, kTileM); ++w) {
        for (int h = 1; h <= std::min(H, kTileM / w); ++h) {
            const int b = std::min(B, kTileM / (w * h));
            const long tiles = (long)ceil_div(W, w) * ceil_div(H, h) * ceil_div(B, b);
            const int rows = w * h * b;
            const bool better = best_tiles < 0 || tiles < best_tiles ||
                                (tiles == best_tiles && (w > wt || (w == wt && rows < best_rows)));
            if (better) { best_tiles = tiles; best_rows = rows; wt = w; ht = h; bt = b; }
        }
    }
}

// Cost model (cycles) used to pick the output-channel tile: a k-iteration is bound either by operand
// delivery from L2 (~36 B/cycle/SM measured on B200 with all SMs pulling) or by the tensor pipe
// (128 x n x 16 MACs = n/2 cycles per UMMA).
int choose_n_tile(int cout_g, long m_tiles, int groups, int k_iters, int KC, int num_sms) {
    int best = 16;
    double best_cost = 1e30;
    for (int n = 16; n <= std::min(cout_g, 256); n += 16) {
        if (cout_g % n) continue;
        const long tiles = m_tiles * groups * (cout_g / n);
        const long waves = (tiles + num_sms - 1) / num_sms;
        const double t_l2 = (128.0 + n) * KC * 2.0 / 36.0;
        const double t_mma = n * KC / 32.0;
        const double t_iter = std::max(t_l2, t_mma) + 30.0;
        const double t_epi = 64.0 + n * 6.0;                     // ~ per-tile epilogue (overlapped unless last)
        const double cost = 2500.0 + waves * (k_iters * t_iter) + t_epi + (waves - 1) * 200.0;
        if (cost < best_cost * 0.999 || (cost <= best_cost * 1.001 && n > best)) { best_cost = cost; best = n; }
    }
    return best;
}

// Epilogue warps: 4 for the plain / head epilogues, 8-12 for the fused ones (tools/exp_epi.py).
int choose_epi_warps(const ConvParams& p) {
    int ew = (p.epi != DD_EPI_NONE && p.epi != DD_EPI_HEAD) ? 8 : 4;
    if (ew == 8 && p.nbuf >= 4) ew = 12;      // fused epilogues are the limiter (tools/exp_epi.py): a third warp group
    if (const char* f = getenv("DD_FORCE_EPI_WARPS")) { const int v = atoi(f); ew = v == 12 ? 12 : (v == 8 ? 8 : 4); }   // tuning
    while (ew > 4 && ew / 4 > p.nbuf) ew -= 4;      // every epilogue warp group needs its own accumulator buffer
    return ew;
}

// Shared-memory-staged epilogue (epilogue_tile_staged).  Measured on B200 against the direct epilogue, bit-identical
// outputs (tools/dev_check_epi_staged.py, profiles/r01_epi_staged_ab.log): two-output epilogues gain 5-30 % on every
// BASELINE-size layer (512->1024 3x3 residual + silu copy 123 -> 104 us, 256->512 67 -> 45 us, silu + raw copy 48 -> 35 us);
// single-output ones are within +-5 % either way (both variants move the same bytes at ~3 TB/s of algorithmic traffic,
// i.e. coalescing was not what limits them), so they keep the direct path.  DD_EPI_STAGED=0/1 forces one variant (tuning).
constexpr uint32_t kSlabBytes = 4096;
constexpr bool kPrefetchResidualDefault = false;
int want_residual_prefetch(const ConvParams& p) {
    bool on = kPrefetchResidualDefault;
    if (const char* f = getenv("DD_EPI_PREFETCH")) on = atoi(f) != 0;
    return (on && p.epi == DD_EPI_RESIDUAL) ? 1 : 0;
}
bool want_staged_epilogue(const ConvParams& p) {
    bool on = p.epi2 != DD_EPI2_NONE;
    if (const char* f = getenv("DD_EPI_STAGED")) on = atoi(f) != 0;
    return on && p.epi != DD_EPI_HEAD && p.nacc == 1;
}

int fill_and_launch(ConvParams& p, const void* x, const void* w_prepped, int groups, cudaStream_t stream,
                    const void* x2 = nullptr, int c1 = 0) {
    PFN_encodeTiled encode = get_encode_fn();
    DD_REQUIRE(encode != nullptr, "dd_mpconv_forward: cuTensorMapEncodeTiled unavailable (driver too old?)");
    const int B = p.B, H = p.H, W = p.W, Cin = p.Cin, Cout = p.Cout;
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    p.cin_g = cin_g; p.cout_g = cout_g;
    choose_box(B, H, W, p.wt, p.ht, p.bt);
    p.tiles_w = ceil_div(W, p.wt); p.tiles_h = ceil_div(H, p.ht); p.tiles_b = ceil_div(B, p.bt);
    p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int num_sms = dd_num_sms();
    const int KC = (cin_g % 64 == 0) ? 64 : 32;
    p.kchunks = cin_g / KC;
    p.k_iters = p.taps * p.kchunks;
    p.n_tile = choose_n_tile(cout_g, p.m_tiles, groups, p.k_iters, KC, num_sms);
    p.n_tiles_per_group = cout_g / p.n_tile;
    p.num_tiles = p.m_tiles * groups * p.n_tiles_per_group;
    p.fd_m_tiles = make_fastdiv(p.m_tiles); p.fd_npg = make_fastdiv(p.n_tiles_per_group);
    p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);
    const uint32_t row_bytes = KC * 2;
    p.a_bytes = (uint32_t)(p.wt * p.ht * p.bt) * row_bytes;
    p.b_bytes = (uint32_t)p.n_tile * row_bytes;
    const uint32_t pair_bytes = kTileM * row_bytes + p.b_bytes;
    p.sub = (pair_bytes <= 24u * 1024u && p.k_iters >= 2) ? 2 : 1;
    const uint32_t stage_bytes = pair_bytes * p.sub;
    p.nacc = 1;      // (splitting the accumulation chain over several TMEM accumulators measured no gain)
    p.nbuf = 2;
    while (p.nbuf < kMaxAccBufs && 2 * p.nbuf * p.n_tile * p.nacc <= 512) p.nbuf *= 2;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.nbuf * p.n_tile * p.nacc)) cols <<= 1;
    p.tmem_cols = cols;
    const int ew = choose_epi_warps(p);
    p.prefetch_res = want_residual_prefetch(p);
    const bool staged = want_staged_epilogue(p);
    const uint32_t slabs_bytes = staged ? (uint32_t)ew * kSlabBytes : 0u;
    p.stages = std::max(2, std::min<int>(kMaxStages, (int)((200u * 1024u - slabs_bytes) / stage_bytes)));
    p.stage_off = staged ? (uint32_t)p.stages * stage_bytes : 0u;

    CUtensorMap tmA, tmA2, tmB;
    const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    p.a_split = x2 ? c1 : 0;
    DD_REQUIRE(x2 == nullptr || (groups == 1 && c1 > 0 && c1 < Cin && c1 % KC == 0),
               "dd_mpconv_forward_cat: the first operand's %d channels must be a multiple of the %d-channel K chunk", c1, KC);
    auto encode_act = [&](CUtensorMap* tm, const void* ptr, int C) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.wt, (cuuint32_t)p.ht, (cuuint32_t)p.bt};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    {
        CUresult r = encode_act(&tmA, x, x2 ? c1 : Cin);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: activation tensor map encode failed (CUresult %d)", (int)r);
        r = encode_act(&tmA2, x2 ? x2 : x, x2 ? Cin - c1 : Cin);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: second activation tensor map encode failed (CUresult %d)", (int)r);
    }
    {
        const cuuint64_t ktot = (cuuint64_t)p.taps * cin_g;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)p.n_tile};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_prepped), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: weight tensor map encode failed (CUresult %d)", (int)r);
    }

    const size_t smem_bytes = (size_t)p.stages * stage_bytes + slabs_bytes + 1024;
    const int grid = std::min(p.num_tiles, num_sms);
    if (getenv("DD_DEBUG_CONV"))
        fprintf(stderr, "[conv] B%d %dx%d %d->%d k%d g%d: box %dx%dx%d m_tiles %d n_tile %d tiles %d k_iters %d KC %d sub %d stages %d\n",
                B, H, W, Cin, Cout, p.kw, groups, p.wt, p.ht, p.bt, p.m_tiles, p.n_tile, p.num_tiles, p.k_iters, KC, p.sub,
                p.stages);
#define DD_LAUNCH_IGEMM(KC_, EW_)                                                                                  \
    do {                                                                                                           \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            DD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<KC_, EW_>,                                       \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));          \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        DD_CHECK_CUDA(launch_pdl(conv_igemm_kernel<KC_, EW_>, grid, 64 + 32 * EW_, smem_bytes, stream, tmA, tmA2, tmB, p)); \
    } while (0)
    if (KC == 64) { if (ew == 12) DD_LAUNCH_IGEMM(64, 12); else if (ew == 8) DD_LAUNCH_IGEMM(64, 8); else DD_LAUNCH_IGEMM(64, 4); }
    else          { if (ew == 12) DD_LAUNCH_IGEMM(32, 12); else if (ew == 8) DD_LAUNCH_IGEMM(32, 8); else DD_LAUNCH_IGEMM(32, 4); }
#undef DD_LAUNCH_IGEMM
    DD_CHECK_LAUNCH();
    return 0;
}


// role-timeline buffer of the diagnostic instantiations (allocated on first use, never freed; shared with conv3x3_dx.cu)
constexpr size_t kTraceBytes = (size_t)kTraceCtas * kTraceSlots * kTraceTiles * sizeof(unsigned long long);
unsigned long long* trace_buffer(bool allocate) { return dd_conv_trace_buffer(allocate); }
int* trace_meta() { return dd_conv_trace_meta(); }

// 3x3 halo variant: used when the image is at least 8 rows tall (levels 0-2 of the 45 s latent).
int launch_halo(ConvParams& p, const void* x, const void* w_prepped, int groups, cudaStream_t stream) {
    PFN_encodeTiled encode = get_encode_fn();
    DD_REQUIRE(encode != nullptr, "dd_mpconv_forward: cuTensorMapEncodeTiled unavailable (driver too old?)");
    const int B = p.B, H = p.H, W = p.W, Cin = p.Cin, Cout = p.Cout;
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    p.cin_g = cin_g; p.cout_g = cout_g;
    p.wt = kHaloW;
    p.ht = H >= 16 ? 16 : 8;                       // 16 groups of 8 pixels: 16 rows x 1 item, or 8 rows x 2 items
    p.bt = std::min(B, 16 / p.ht);
    p.tiles_w = ceil_div(W, kHaloW); p.tiles_h = ceil_div(H, p.ht); p.tiles_b = ceil_div(B, p.bt);
    p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    p.halo_bytes = (uint32_t)((p.ht + 2) * p.bt * kHaloPitch) * 128u;
    // a UMMA always reads 16 groups x 10-row pitch (+ the largest tap shift) from the stage, whether or not the
    // tile has 16 valid groups: size the stage for that so reads never leave the allocation
    const uint32_t read_span = (uint32_t)(15 * kHaloPitch + 2 * kHaloPitch * p.bt + 2 + 8) * 128u;
    p.halo_stride = (std::max(p.halo_bytes, read_span) + 1023u) / 1024u * 1024u;
    p.kchunks = ceil_div(cin_g, 64);
    p.ks_last = (cin_g - 64 * (p.kchunks - 1)) / 16;
    p.k_iters = p.kchunks;
    p.nacc = 1;
    p.prefetch_res = want_residual_prefetch(p);
    const bool staged = want_staged_epilogue(p);
    // direct epilogue: 214 KB of operands; staged: 224 KB shared between operands and the 4 KB-per-warp slabs (the output
    // tile width is chosen as if there were 4 epilogue warps, then the warp count is cut back to what still fits)
    const uint32_t total = staged ? 224u * 1024u : 214u * 1024u;
    const uint32_t budget = total - (staged ? 4u * kSlabBytes : 0u);
    // tuning experiments only (tools/exp_halo_stages.sh): cap the output-channel tile / ask for more halo stages than the
    // two the default insists on, to trade weight-panel width against activation boxes in flight
    static const int ntile_cap = getenv("DD_HALO_NTILE_MAX") ? std::max(16, atoi(getenv("DD_HALO_NTILE_MAX"))) : 256;
    static const uint32_t min_stages = getenv("DD_HALO_MIN_STAGES") ? (uint32_t)std::max(2, atoi(getenv("DD_HALO_MIN_STAGES"))) : 2u;
    int n_tile = 0;
    for (int n = 16; n <= std::min(cout_g, std::min(256, ntile_cap)); n += 16) {
        if (cout_g % n) continue;
        if ((uint32_t)p.kchunks * 9u * n * 128u + min_stages * p.halo_stride <= budget) n_tile = n;
    }
    if (n_tile == 0) return -1;     // weight panel does not fit: caller falls back to the per-tap kernel
    p.n_tile = n_tile;
    p.n_tiles_per_group = cout_g / n_tile;
    p.num_tiles = p.m_tiles * groups * p.n_tiles_per_group;
    p.fd_m_tiles = make_fastdiv(p.m_tiles); p.fd_npg = make_fastdiv(p.n_tiles_per_group);
    p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);
    p.b_block_bytes = (uint32_t)n_tile * 128u;
    const uint32_t b_total = (uint32_t)p.kchunks * 9u * p.b_block_bytes;
    p.nacc = 1;
    p.dbg_taps = 9;
    p.dbg_nostore = getenv("DD_DBG_NOSTORE") != nullptr;
    p.nbuf = 2;
    while (p.nbuf < kMaxAccBufs && 2 * p.nbuf * p.n_tile * p.nacc <= 512) p.nbuf *= 2;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.nbuf * p.n_tile * p.nacc)) cols <<= 1;
    p.tmem_cols = cols;
    int ew = choose_epi_warps(p);
    if (staged) {
        // keep at least three halo stages (or as many as four epilogue warps would leave) before adding warp groups
        const int want_stages = std::min(3, (int)((budget - b_total) / p.halo_stride));
        while (ew > 4 && b_total + (uint32_t)want_stages * p.halo_stride + (uint32_t)ew * kSlabBytes > total) ew -= 4;
    }
    const uint32_t slabs_bytes = staged ? (uint32_t)ew * kSlabBytes : 0u;
    p.a_stages = std::max(2, std::min<int>(kMaxAStages, (int)((total - slabs_bytes - b_total) / p.halo_stride)));
    p.stage_off = staged ? b_total + (uint32_t)p.a_stages * p.halo_stride : 0u;

    CUtensorMap tmA, tmB;
    {
        // dims ordered (C, W, B, H): the box lands in smem as [h][b][w][c], i.e. the image rows of the bt batch
        // items of a tile are interleaved, which keeps the stride between the tile's 8-pixel groups uniform
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)B, (cuuint64_t)H};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)H * W * Cin * 2, (cuuint64_t)W * Cin * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)kHaloPitch, (cuuint32_t)p.bt, (cuuint32_t)(p.ht + 2)};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: halo tensor map encode failed (CUresult %d)", (int)r);
    }
    {
        // weights [Cout][9][cin_g] viewed as (ci, co, tap): one box (64 channels, n_tile, 9 taps) is a whole chunk of the
        // panel, landing as [tap][co][ci] (one TMA instead of nine per chunk)
        cuuint64_t dims[3] = {(cuuint64_t)cin_g, (cuuint64_t)Cout, 9};
        cuuint64_t strides[2] = {(cuuint64_t)9 * cin_g * 2, (cuuint64_t)cin_g * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)p.n_tile, 9};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_prepped), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: weight tensor map encode failed (CUresult %d)", (int)r);
    }
    const size_t smem_bytes = (size_t)b_total + (size_t)p.a_stages * p.halo_stride + slabs_bytes + 1024;
    const int grid = std::min(p.num_tiles, dd_num_sms());
#define DD_LAUNCH_HALO(EW_)                                                                                        \
    do {                                                                                                           \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            DD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<EW_>,                                          \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));          \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        DD_CHECK_CUDA(launch_pdl(conv3x3_halo_kernel<EW_>, grid, 64 + 32 * EW_, smem_bytes, stream, tmA, tmB, p));   \
    } while (0)
    static const bool trace_on = getenv("DD_CONV_TRACE") != nullptr;       // diagnostic instantiation, tools/trace_halo.py
    if (trace_on) {
        DD_REQUIRE(trace_buffer(true) != nullptr, "dd_mpconv_forward: could not allocate the trace buffer");
        p.trace = trace_buffer(false);
        DD_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, kTraceBytes, stream));
#define DD_LAUNCH_HALO_TRACE(EW_)                                                                                  \
    do {                                                                                                           \
        DD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<EW_, true>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));              \
        DD_CHECK_CUDA(launch_pdl(conv3x3_halo_kernel<EW_, true>, grid, 64 + 32 * EW_, smem_bytes, stream, tmA, tmB, p)); \
    } while (0)
        if (ew == 12) DD_LAUNCH_HALO_TRACE(12); else if (ew == 8) DD_LAUNCH_HALO_TRACE(8); else DD_LAUNCH_HALO_TRACE(4);
#undef DD_LAUNCH_HALO_TRACE
        trace_meta()[0] = p.num_tiles; trace_meta()[1] = grid; trace_meta()[2] = p.n_tile; trace_meta()[3] = p.a_stages;
        trace_meta()[4] = p.nbuf; trace_meta()[5] = ew; trace_meta()[6] = p.kchunks; trace_meta()[7] = p.stage_off ? 1 : 0;
        DD_CHECK_LAUNCH();
        return 0;
    }
    if (ew == 12) DD_LAUNCH_HALO(12); else if (ew == 8) DD_LAUNCH_HALO(8); else DD_LAUNCH_HALO(4);
#undef DD_LAUNCH_HALO
    DD_CHECK_LAUNCH();
    return 0;
}

int launch_conv(ConvParams& p, const void* x, const void* w_prepped, int groups, cudaStream_t stream) {
    static const bool no_halo = getenv("DD_DISABLE_HALO") != nullptr;     // tuning experiments only
    if (p.taps == 9 && groups > 1 && p.epi != DD_EPI_HEAD) {
        // grouped 3x3 layers: tap-stacked kernel (conv3x3_dx.cu); -1 = shape not handled there
        DxConvArgs a{};
        a.x = x; a.w = w_prepped; a.out = p.out;
        a.B = p.B; a.H = p.H; a.W = p.W; a.Cin = p.Cin; a.Cout = p.Cout; a.groups = groups;
        a.epi = p.epi; a.epi2 = p.epi2; a.alpha = p.alpha; a.beta = p.beta; a.clip = p.clip;
        a.scale = p.scale; a.scale2 = p.scale2; a.residual = p.residual; a.out2 = p.out2;
        const int r = dd_launch_conv3x3_dx(a, stream);
        if (r >= 0) return r;
    }
    if (p.taps == 9 && p.H >= 8 && p.W >= kHaloW && (p.Cin / groups) % 16 == 0 && p.Cin >= 64 && !no_halo) {
        ConvParams q = p;
        const int r = launch_halo(q, x, w_prepped, groups, stream);
        if (r >= 0) return r;
    }
    return fill_and_launch(p, x, w_prepped, groups, stream);
}

}  // namespace

unsigned long long* dd_conv_trace_buffer(bool allocate) {
    static unsigned long long* buf = nullptr;
    if (buf == nullptr && allocate && cudaMalloc(&buf, kTraceBytes) != cudaSuccess) buf = nullptr;
    return buf;
}
int* dd_conv_trace_meta() {
    static int meta[8] = {0};
    return meta;
}

extern "C" int dd_conv_trace_read(unsigned long long* stamps_host, int n_stamps, int* meta_host) {
    DD_REQUIRE(stamps_host && meta_host, "dd_conv_trace_read: null pointer");
    DD_REQUIRE(n_stamps == kTraceCtas * kTraceSlots * kTraceTiles, "dd_conv_trace_read: expected %d stamps",
               kTraceCtas * kTraceSlots * kTraceTiles);
    DD_REQUIRE(trace_buffer(false) != nullptr, "dd_conv_trace_read: no traced launch yet (set DD_CONV_TRACE=1 before the "
                                               "library is loaded and run a 3x3 layer of >= 8 rows)");
    DD_CHECK_CUDA(cudaDeviceSynchronize());
    DD_CHECK_CUDA(cudaMemcpy(stamps_host, trace_buffer(false), kTraceBytes, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) meta_host[i] = trace_meta()[i];
    return 0;
}

extern "C" int dd_mpconv_forward(const void* x, const void* w_prepped, void* out, int B, int H, int W, int Cin,
                                 int Cout, int ksize, int groups, const dd_conv_epilogue* epi, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && w_prepped && out, "dd_mpconv_forward: null pointer");
    DD_REQUIRE(ksize == 1 || ksize == 3, "dd_mpconv_forward: kernel size %d unsupported (1 or 3)", ksize);
    DD_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "dd_mpconv_forward: bad groups");
    DD_REQUIRE((Cin / groups) % 32 == 0, "dd_mpconv_forward: Cin/groups=%d must be a multiple of 32", Cin / groups);
    DD_REQUIRE((Cout / groups) % 16 == 0, "dd_mpconv_forward: Cout/groups=%d must be a multiple of 16", Cout / groups);
    DD_REQUIRE(B > 0 && H > 0 && W > 0, "dd_mpconv_forward: empty input");

    ConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.kh = p.kw = ksize; p.taps = ksize * ksize;
    if (epi) {
        p.epi = epi->mode; p.epi2 = epi->mode2;
        p.alpha = epi->alpha; p.beta = epi->beta;
        p.clip = epi->clip > 0.f ? epi->clip : INFINITY;
        p.scale = static_cast<const float*>(epi->scale);
        p.scale2 = static_cast<const float*>(epi->scale2);
        p.residual = static_cast<const __nv_bfloat16*>(epi->residual);
        p.out2 = static_cast<__nv_bfloat16*>(epi->out2);
        DD_REQUIRE(p.epi >= DD_EPI_NONE && p.epi <= DD_EPI_RESIDUAL, "dd_mpconv_forward: bad epilogue mode %d", p.epi);
        DD_REQUIRE(p.epi != DD_EPI_SCALE_SILU || p.scale, "dd_mpconv_forward: epilogue scale missing");
        DD_REQUIRE(p.epi != DD_EPI_RESIDUAL || p.residual, "dd_mpconv_forward: epilogue residual missing");
        DD_REQUIRE(p.epi2 == DD_EPI2_NONE || p.out2, "dd_mpconv_forward: epilogue out2 missing");
        DD_REQUIRE(p.epi2 != DD_EPI2_SCALE || p.scale2, "dd_mpconv_forward: epilogue scale2 missing");
    } else {
        p.epi = DD_EPI_NONE; p.epi2 = DD_EPI2_NONE; p.clip = INFINITY;
    }
    p.out = static_cast<__nv_bfloat16*>(out);
    return launch_conv(p, x, w_prepped, groups, stream);
}

// 1x1 MPConv over the channel concatenation [x1 | x2] without materialising it (decoder conv_skip after mp_cat,
// unet_edm2_b4.py:295-296 + :129): the K loop switches tensor maps at channel C1.  The mp_cat weights wa / wb belong in the
// prepared weight's columns.
extern "C" int dd_mpconv_forward_cat(const void* x1, int C1, const void* x2, int C2, const void* w_prepped, void* out, int B,
                                     int H, int W, int Cout, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x1 && x2 && w_prepped && out, "dd_mpconv_forward_cat: null pointer");
    DD_REQUIRE(C1 > 0 && C2 > 0 && (C1 + C2) % 32 == 0 && Cout % 16 == 0, "dd_mpconv_forward_cat: bad channel counts");
    DD_REQUIRE(B > 0 && H > 0 && W > 0, "dd_mpconv_forward_cat: empty input");
    ConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = C1 + C2; p.Cout = Cout;
    p.kh = p.kw = 1; p.taps = 1;
    p.epi = DD_EPI_NONE; p.epi2 = DD_EPI2_NONE; p.clip = INFINITY;
    p.out = static_cast<__nv_bfloat16*>(out);
    return fill_and_launch(p, x1, w_prepped, 1, stream, x2, C1);
}

// UNet head on the tensor cores: 3x3 conv to `Cout` (<= 16) channels whose weights were padded to 16 output
// rows by dd_weight_prep(pad_rows), fused with the EDM output preconditioning (fp32 NCHW result).
extern "C" int dd_conv_out(const void* x, const void* w_prepped16, const float* x_in, const float* sigma,
                           float sigma_data, const float* x_ref, float* d_out, int B, int C, int H, int W, int Cout,
                           void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && w_prepped16 && x_in && sigma && d_out, "dd_conv_out: null pointer");
    DD_REQUIRE(Cout >= 1 && Cout <= 16, "dd_conv_out: out_channels=%d unsupported (1..16)", Cout);
    DD_REQUIRE(C % 32 == 0, "dd_conv_out: C=%d must be a multiple of 32", C);
    ConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = C; p.Cout = 16;
    p.kh = p.kw = 3; p.taps = 9;
    p.epi = DD_EPI_HEAD; p.epi2 = DD_EPI2_NONE; p.clip = INFINITY;
    p.head_cout = Cout; p.sigma_data = sigma_data; p.sigma = sigma; p.x_in = x_in; p.x_ref = x_ref; p.d_out = d_out;
    return launch_conv(p, x, w_prepped16, 1, stream);
}
