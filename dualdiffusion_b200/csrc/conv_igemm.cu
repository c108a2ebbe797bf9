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
//   warps 2..5 : epilogue       (tcgen05.ld -> registers -> fused elementwise -> global)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1; the kernel is persistent over a static round-robin tile schedule.
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <math.h>

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvParams {
    int B, H, W, Cin, Cout;
    int kh, kw, taps;
    int cin_g, cout_g;
    int wt, ht, bt;                 // pixel box of one M tile (wt*ht*bt <= 128)
    int tiles_w, tiles_h, tiles_b, m_tiles;
    int n_tile, n_tiles_per_group;
    int kchunks;                    // cin_g / KC
    int num_tiles;
    int stages;
    uint32_t a_bytes, b_bytes;      // bytes landed per stage by the two TMA boxes
    uint32_t tmem_cols;
    // epilogue
    int epi;                        // DD_EPI_*
    int epi2;                       // DD_EPI2_*
    float alpha, beta, clip;
    const float* scale;             // [B][Cout]   (DD_EPI_SCALE_SILU)
    const float* scale2;            // [B][Cout]   (DD_EPI2_SCALE)
    const __nv_bfloat16* residual;  // [B][H][W][Cout]
    __nv_bfloat16* out;
    __nv_bfloat16* out2;
};

struct TileCoord {
    int g, n_idx, b0, h0, w0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
    TileCoord t;
    const int m_idx = tile % p.m_tiles;
    const int rest = tile / p.m_tiles;
    t.n_idx = rest % p.n_tiles_per_group;
    t.g = rest / p.n_tiles_per_group;
    t.w0 = (m_idx % p.tiles_w) * p.wt;
    t.h0 = ((m_idx / p.tiles_w) % p.tiles_h) * p.ht;
    t.b0 = (m_idx / (p.tiles_w * p.tiles_h)) * p.bt;
    return t;
}

template <int KC>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvParams p) {
    constexpr uint32_t kRowBytes = KC * 2;           // one swizzle span per row
    constexpr uint32_t kABufBytes = kTileM * kRowBytes;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B swizzle atom
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_buf_bytes = (uint32_t)p.n_tile * kRowBytes;
    const uint32_t stage_bytes = kABufBytes + b_buf_bytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full_bar[a], 1);
            ptx::mbar_init(&tmem_empty_bar[a], 4);
        }
        ptx::mbar_fence_init();
        ptx::fence_proxy_async_smem();
    }
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_slot, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int k_iters = p.taps * p.kchunks;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile);
                const int a_c0 = t.g * p.cin_g;
                const int b_row = t.g * p.cout_g + t.n_idx * p.n_tile;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int dy = tap / p.kw - p.kh / 2;
                    const int dx = tap % p.kw - p.kw / 2;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = smem + stage * stage_bytes;
                        uint8_t* b_dst = a_dst + kABufBytes;
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
                        ptx::tma_load_4d(a_dst, &tmA, &full_bar[stage], a_c0 + kc * KC, t.w0 + dx, t.h0 + dy, t.b0);
                        ptx::tma_load_2d(b_dst, &tmB, &full_bar[stage], tap * p.cin_g + kc * KC, b_row);
                        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_bf16(kTileM, p.n_tile);
            uint32_t stage = 0, phase = 0;
            uint32_t local = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
                const uint32_t acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                ptx::tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.n_tile;
                for (int it = 0; it < k_iters; ++it) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tcgen05_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem + stage * stage_bytes);
                    const uint32_t b_addr = a_addr + kABufBytes;
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks) {
                        const uint64_t a_desc = ptx::make_kmajor_desc(a_addr + ks * 32, kRowBytes);
                        const uint64_t b_desc = ptx::make_kmajor_desc(b_addr + ks * 32, kRowBytes);
                        ptx::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);        // frees the smem stage once the MMAs retire
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full_bar[acc]);          // accumulator complete -> epilogue
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;           // row of the 128-row tile == box-linear pixel index
        const int ww = row % p.wt;
        const int hh = (row / p.wt) % p.ht;
        const int bb = row / (p.wt * p.ht);
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
            const TileCoord t = decode_tile(p, tile);
            const uint32_t acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            const int b = t.b0 + bb, h = t.h0 + hh, w = t.w0 + ww;
            const bool valid = (bb < p.bt) && (b < p.B) && (h < p.H) && (w < p.W);
            const size_t pix = ((size_t)b * p.H + h) * p.W + w;
            const int ch0 = t.g * p.cout_g + t.n_idx * p.n_tile;

            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)p.n_tile;

            for (int c = 0; c < p.n_tile; c += 32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(taddr + c, r);
                ptx::tmem_ld_wait();
                if (valid) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    const int ch = ch0 + c;
                    if (p.epi == DD_EPI_SCALE_SILU) {
                        const float4* sc = reinterpret_cast<const float4*>(p.scale + (size_t)b * p.Cout + ch);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 s = __ldg(sc + i);
                            v[4 * i + 0] = mp_silu_f(v[4 * i + 0] * s.x);
                            v[4 * i + 1] = mp_silu_f(v[4 * i + 1] * s.y);
                            v[4 * i + 2] = mp_silu_f(v[4 * i + 2] * s.z);
                            v[4 * i + 3] = mp_silu_f(v[4 * i + 3] * s.w);
                        }
                    } else if (p.epi == DD_EPI_RESIDUAL) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + ch);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 q = __ldg(rp + i);
                            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = unpack_bf16x2(u[j]);
                                float a0 = p.alpha * v[8 * i + 2 * j + 0] + p.beta * f.x;
                                float a1 = p.alpha * v[8 * i + 2 * j + 1] + p.beta * f.y;
                                v[8 * i + 2 * j + 0] = fminf(fmaxf(a0, -p.clip), p.clip);
                                v[8 * i + 2 * j + 1] = fminf(fmaxf(a1, -p.clip), p.clip);
                            }
                        }
                    }
                    uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.Cout + ch);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 q;
                        q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
                        q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                        q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
                        q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                        op[i] = q;
                    }
                    if (p.epi2 != DD_EPI2_NONE) {
                        if (p.epi2 == DD_EPI2_SILU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = mp_silu_f(v[i]);
                        } else {
                            const float4* sc = reinterpret_cast<const float4*>(p.scale2 + (size_t)b * p.Cout + ch);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 s = __ldg(sc + i);
                                v[4 * i + 0] *= s.x; v[4 * i + 1] *= s.y; v[4 * i + 2] *= s.z; v[4 * i + 3] *= s.w;
                            }
                        }
                        uint4* op2 = reinterpret_cast<uint4*>(p.out2 + pix * p.Cout + ch);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 q;
                            q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
                            q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                            q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
                            q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                            op2[i] = q;
                        }
                    }
                }
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

int choose_n_tile(int cout_g, long m_tiles, int groups, int num_sms) {
    // candidates: multiples of 32 that divide cout_g, at most 256 wide
    int best = 32;
    for (int n = 32; n <= std::min(cout_g, 256); n += 32) {
        if (cout_g % n) continue;
        best = n;
    }
    // small problems: prefer narrower tiles until the grid covers the SMs
    while (best > 32) {
        const long tiles = m_tiles * groups * (cout_g / best);
        if (tiles >= num_sms) break;
        int next = 0;
        for (int n = best - 32; n >= 32; n -= 32) if (cout_g % n == 0) { next = n; break; }
        if (!next) break;
        best = next;
    }
    return best;
}

}  // namespace

extern "C" int dd_mpconv_forward(const void* x, const void* w_prepped, void* out, int B, int H, int W, int Cin,
                                 int Cout, int ksize, int groups, const dd_conv_epilogue* epi, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && w_prepped && out, "dd_mpconv_forward: null pointer");
    DD_REQUIRE(ksize == 1 || ksize == 3, "dd_mpconv_forward: kernel size %d unsupported (1 or 3)", ksize);
    DD_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "dd_mpconv_forward: bad groups");
    const int cin_g = Cin / groups, cout_g = Cout / groups;
    DD_REQUIRE(cin_g % 32 == 0, "dd_mpconv_forward: Cin/groups=%d must be a multiple of 32", cin_g);
    DD_REQUIRE(cout_g % 32 == 0, "dd_mpconv_forward: Cout/groups=%d must be a multiple of 32", cout_g);
    DD_REQUIRE(B > 0 && H > 0 && W > 0, "dd_mpconv_forward: empty input");
    PFN_encodeTiled encode = get_encode_fn();
    DD_REQUIRE(encode != nullptr, "dd_mpconv_forward: cuTensorMapEncodeTiled unavailable (driver too old?)");

    ConvParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.kh = p.kw = ksize; p.taps = ksize * ksize;
    p.cin_g = cin_g; p.cout_g = cout_g;
    choose_box(B, H, W, p.wt, p.ht, p.bt);
    p.tiles_w = ceil_div(W, p.wt); p.tiles_h = ceil_div(H, p.ht); p.tiles_b = ceil_div(B, p.bt);
    p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int num_sms = dd_num_sms();
    p.n_tile = choose_n_tile(cout_g, p.m_tiles, groups, num_sms);
    p.n_tiles_per_group = cout_g / p.n_tile;
    const int KC = (cin_g % 64 == 0) ? 64 : 32;
    p.kchunks = cin_g / KC;
    p.num_tiles = p.m_tiles * groups * p.n_tiles_per_group;
    const uint32_t row_bytes = KC * 2;
    p.a_bytes = (uint32_t)(p.wt * p.ht * p.bt) * row_bytes;
    p.b_bytes = (uint32_t)p.n_tile * row_bytes;
    const uint32_t stage_bytes = kTileM * row_bytes + p.b_bytes;
    p.stages = std::max(2, std::min<int>(kMaxStages, (int)((200u * 1024u) / stage_bytes)));
    uint32_t cols = 32;
    while (cols < 2u * p.n_tile) cols <<= 1;
    p.tmem_cols = cols;

    if (epi) {
        p.epi = epi->mode; p.epi2 = epi->mode2;
        p.alpha = epi->alpha; p.beta = epi->beta;
        p.clip = epi->clip > 0.f ? epi->clip : INFINITY;
        p.scale = static_cast<const float*>(epi->scale);
        p.scale2 = static_cast<const float*>(epi->scale2);
        p.residual = static_cast<const __nv_bfloat16*>(epi->residual);
        p.out2 = static_cast<__nv_bfloat16*>(epi->out2);
        DD_REQUIRE(p.epi != DD_EPI_SCALE_SILU || p.scale, "dd_mpconv_forward: epilogue scale missing");
        DD_REQUIRE(p.epi != DD_EPI_RESIDUAL || p.residual, "dd_mpconv_forward: epilogue residual missing");
        DD_REQUIRE(p.epi2 == DD_EPI2_NONE || p.out2, "dd_mpconv_forward: epilogue out2 missing");
        DD_REQUIRE(p.epi2 != DD_EPI2_SCALE || p.scale2, "dd_mpconv_forward: epilogue scale2 missing");
    } else {
        p.epi = DD_EPI_NONE; p.epi2 = DD_EPI2_NONE; p.clip = INFINITY;
    }
    p.out = static_cast<__nv_bfloat16*>(out);

    // tensor maps
    CUtensorMap tmA, tmB;
    const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.wt, (cuuint32_t)p.ht, (cuuint32_t)p.bt};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_forward: activation tensor map encode failed (CUresult %d)", (int)r);
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

    const size_t smem_bytes = (size_t)p.stages * stage_bytes + 1024;
    const int grid = std::min(p.num_tiles, num_sms);
    if (KC == 64) {
        static bool attr_done = false;
        if (!attr_done) {
            DD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               220 * 1024));
            attr_done = true;
        }
        conv_igemm_kernel<64><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, p);
    } else {
        static bool attr_done = false;
        if (!attr_done) {
            DD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               220 * 1024));
            attr_done = true;
        }
        conv_igemm_kernel<32><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmB, p);
    }
    DD_CHECK_LAUNCH();
    return 0;
}
