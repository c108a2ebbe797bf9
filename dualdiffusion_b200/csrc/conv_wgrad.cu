// MPConv weight gradient on the Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the weight-gradient half of autograd's conv2d backward for MPConv.forward
// (/root/reference/src/modules/mp_tools.py:369; driven by loss.backward() in
// training/trainer.py:1022-1044) for the stride-1 1x1 and 3x3 (grouped) convolutions of the EDM2 UNet:
//
//   dW[co][tap][ci] = sum over pixels p of dY[p][co] * X[p + tap][g(co)*cin_g + ci]       (fp32)
//
// GEMM view: M = output channels, N = input channels, K = pixels.  Activations are NHWC, so the pixel (K)
// dimension is the slow one of both operands: the UMMA shared-memory descriptors are MN-major (SWIZZLE_128B:
// one 128-byte row = 64 channels of one pixel, 8 pixels per swizzle atom) and no tensor is transposed.
//   3x3: one (8 w) x (ht h) pixel tile per pipeline stage.  dY lands as [h][w][co]; X lands once, as a
//        (8+2) x (ht+2) halo tile, and the operand of filter tap (dy,dx) is that same tile read from row
//        dy*10+dx with a 10-row pitch between 8-pixel groups (zero padding = TMA out-of-bounds fill).
//        Each tap has its own 128 x 64 fp32 accumulator in TMEM; 9 x 64 columns do not fit the 512 TMEM
//        columns, so a CTA owns 5 or 4 taps.
//   1x1: 128 consecutive pixels per stage, up to 256 input channels (4 swizzle atoms) per UMMA.
// A CTA owns (128 output channels) x (64..256 input channels) x (its taps) x (a contiguous range of pixel
// tiles); partial sums of different pixel ranges meet in HBM through vector fp32 reductions
// (red.global.add.v4.f32).  Grouped convolutions compute 128 x 64 channel tiles that straddle the block
// diagonal and store only the in-group part.
//   warp 0     : TMA producer
//   warp 1     : TMEM allocator + tcgen05.mma issuer
//   warps 2..5 : epilogue (tcgen05.ld -> 128-byte contiguous fp32 rows of dW)
#include "common.cuh"
#include "dualdiffusion_b200.h"

#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int kWgStages = 4;
constexpr int kWgThreads = 192;

struct WgradParams {
    int B, H, W, Cin, Cout, taps, cin_g, cout_g;
    int halo;                 // 1: 3x3 halo mode, 0: 1x1 flat mode
    int ht;                   // halo: image rows per pixel tile
    int tiles_w, tiles_h;
    int pix_tiles;            // pixel tiles in the whole activation
    int ksteps;               // UMMA K=16 steps per pixel tile
    int m_tiles, J, nch, tap_parts, taps_per_part, splits;
    int stages;
    uint32_t a_box_bytes, x_box_bytes, x_box_stride, stage_bytes;
    uint32_t tmem_cols;
    uint32_t b_sbo;           // bytes between consecutive 8-pixel groups of the X operand
    int use_red;
    float scale;              // dW is multiplied by this (mp_sum coefficient folded into the gradient)
    float* dw;
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor: the MN extent advances by `lbo` bytes per 64-element
// atom, the K extent by `sbo` bytes per 8 rows (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_mn_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ WgradParams p) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kWgStages];
    __shared__ __align__(8) uint64_t empty_bar[kWgStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    // ---- work item of this CTA (uniform) ----
    int idx = blockIdx.x;
    const int split = idx % p.splits; idx /= p.splits;
    const int tp = idx % p.tap_parts; idx /= p.tap_parts;
    const int j = idx % p.J;
    const int m0 = (idx / p.J) * 128;
    const int g_lo = m0 / p.cout_g, g_hi = min(p.Cout - 1, m0 + 127) / p.cout_g;
    const int lo = (g_lo * p.cin_g) / 64 * 64;
    const int hi = min((p.Cin + 63) / 64 * 64, ((g_hi + 1) * p.cin_g + 63) / 64 * 64);
    const int cnt = min(p.nch, (hi - lo) / 64 - j * p.nch);       // 64-channel X chunks of this CTA
    if (cnt <= 0) return;
    const int n0 = lo + 64 * j * p.nch;
    const int t0 = tp * p.taps_per_part;
    const int nt = min(p.taps_per_part, p.taps - t0);
    const int pt_begin = (int)((long)split * p.pix_tiles / p.splits);
    const int pt_end = (int)((long)(split + 1) * p.pix_tiles / p.splits);
    if (pt_begin >= pt_end) return;
    const int a_boxes = (m0 + 64 < p.Cout) ? 2 : 1;
    const uint32_t acc_stride = 64u * (uint32_t)p.nch;            // TMEM columns between tap accumulators

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmY);
        ptx::prefetch_tensormap(&tmX);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(&tmem_full_bar, 1);
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

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // warp-uniform loop, one elected lane issues (a single-lane loop makes the compiler wrap every UTMALDG in a
        // lane-serialisation loop, ~250 cycles per box); the pixel-tile coordinates advance without divisions
        {
            uint32_t stage = 0, phase = 0;
            int tw = 0, th = 0, b = 0;
            if (p.halo) {
                tw = pt_begin % p.tiles_w;
                th = (pt_begin / p.tiles_w) % p.tiles_h;
                b = pt_begin / (p.tiles_w * p.tiles_h);
            }
            for (int pt = pt_begin; pt < pt_end; ++pt) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* a_dst = smem + (size_t)stage * p.stage_bytes;
                uint8_t* x_dst = a_dst + 2 * p.a_box_bytes;
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)a_boxes * p.a_box_bytes + (uint32_t)cnt * p.x_box_bytes);
                    if (p.halo) {
                        for (int a = 0; a < a_boxes; ++a)
                            ptx::tma_load_4d(a_dst + a * p.a_box_bytes, &tmY, &full_bar[stage], m0 + 64 * a, tw * 8, th * p.ht, b);
                        for (int c = 0; c < cnt; ++c)
                            ptx::tma_load_4d(x_dst + c * p.x_box_stride, &tmX, &full_bar[stage], n0 + 64 * c, tw * 8 - 1,
                                             th * p.ht - 1, b);
                    } else {
                        const int pix0 = pt * 128;
                        for (int a = 0; a < a_boxes; ++a)
                            ptx::tma_load_2d(a_dst + a * p.a_box_bytes, &tmY, &full_bar[stage], m0 + 64 * a, pix0);
                        for (int c = 0; c < cnt; ++c)
                            ptx::tma_load_2d(x_dst + c * p.x_box_stride, &tmX, &full_bar[stage], n0 + 64 * c, pix0);
                    }
                }
                __syncwarp();
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++b; } }
                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        // both operands MN-major (bits 15/16 of the instruction descriptor)
        const uint32_t idesc = ptx::make_idesc_bf16(128, 64 * cnt) | (1u << 15) | (1u << 16);
        const uint32_t smem_base = ptx::smem_u32(smem);
        uint32_t stage = 0, phase = 0;
        for (int pt = pt_begin; pt < pt_end; ++pt) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tcgen05_fence_after();
            if (ptx::elect_one()) {
                const uint32_t a_base = smem_base + stage * p.stage_bytes;
                const uint32_t x_base = a_base + 2 * p.a_box_bytes;
                const uint64_t a_desc0 = make_mn_desc_sw128(a_base, p.a_box_bytes, 1024);
                const uint32_t first = (pt == pt_begin) ? 1u : 0u;
                for (int t = 0; t < nt; ++t) {
                    const int tap = t0 + t;
                    const uint32_t shift = p.halo ? (uint32_t)((tap / 3) * 10 + (tap % 3)) * 128u : 0u;
                    const uint64_t b_desc0 = make_mn_desc_sw128(x_base + shift, p.x_box_stride, p.b_sbo);
                    const uint32_t d_tmem = tmem_base + (uint32_t)t * acc_stride;
                    for (int ks = 0; ks < p.ksteps; ++ks) {
                        // one K=16 step = two 8-pixel groups: +2 KB in the dY tile, +2*sbo in the X tile
                        ptx::umma_bf16_ss(d_tmem, a_desc0 + (uint64_t)ks * (2048u >> 4),
                                          b_desc0 + (uint64_t)ks * ((2u * p.b_sbo) >> 4), idesc,
                                          (first && ks == 0) ? 0u : 1u);
                    }
                }
                ptx::umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit(&tmem_full_bar);
        __syncwarp();
    } else {
        // ------------------------------ epilogue ------------------------------
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int co = m0 + quad * 32 + lane;
        const bool row_ok = co < p.Cout;
        const int g = row_ok ? co / p.cout_g : 0;
        const int c_lo = g * p.cin_g, c_hi = c_lo + p.cin_g;
        ptx::mbar_wait(&tmem_full_bar, 0);
        ptx::tcgen05_fence_after();
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int t = 0; t < nt; ++t) {
            const int tap = t0 + t;
            for (int c32 = 0; c32 < 2 * cnt; ++c32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(lane_addr + (uint32_t)t * acc_stride + (uint32_t)c32 * 32u, r);
                ptx::tmem_ld_wait();
                const int ci_abs = n0 + 32 * c32;
                if (row_ok && ci_abs >= c_lo && ci_abs + 32 <= c_hi) {
                    float* dst = p.dw + ((size_t)co * p.taps + tap) * p.cin_g + (ci_abs - c_lo);
                    if (p.use_red) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            red_add_v4(dst + 4 * q, __uint_as_float(r[4 * q]) * p.scale, __uint_as_float(r[4 * q + 1]) * p.scale,
                                       __uint_as_float(r[4 * q + 2]) * p.scale, __uint_as_float(r[4 * q + 3]) * p.scale);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            reinterpret_cast<float4*>(dst)[q] =
                                make_float4(__uint_as_float(r[4 * q]) * p.scale, __uint_as_float(r[4 * q + 1]) * p.scale,
                                            __uint_as_float(r[4 * q + 2]) * p.scale, __uint_as_float(r[4 * q + 3]) * p.scale);
                    }
                }
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiledW wg_encode_fn() {
    static PFN_encodeTiledW fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiledW>(sym);
    }
    return fn;
}

}  // namespace

extern "C" int dd_mpconv_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout,
                               int ksize, int groups, float scale, int accumulate, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && dy && dw, "dd_mpconv_wgrad: null pointer");
    DD_REQUIRE(ksize == 1 || ksize == 3, "dd_mpconv_wgrad: kernel size %d unsupported (1 or 3)", ksize);
    DD_REQUIRE(groups >= 1 && Cin % groups == 0 && Cout % groups == 0, "dd_mpconv_wgrad: bad groups");
    DD_REQUIRE((Cin / groups) % 32 == 0, "dd_mpconv_wgrad: Cin/groups=%d must be a multiple of 32", Cin / groups);
    DD_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "dd_mpconv_wgrad: channel counts must be multiples of 8");
    DD_REQUIRE(B > 0 && H > 0 && W > 0, "dd_mpconv_wgrad: empty input");
    PFN_encodeTiledW encode = wg_encode_fn();
    DD_REQUIRE(encode != nullptr, "dd_mpconv_wgrad: cuTensorMapEncodeTiled unavailable (driver too old?)");

    WgradParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.taps = ksize * ksize;
    p.cin_g = Cin / groups; p.cout_g = Cout / groups;
    p.halo = ksize == 3;
    p.scale = scale;
    p.dw = dw;
    const int chunks_total = (Cin + 63) / 64;
    if (p.halo) {
        p.ht = std::min(16, (H + 1) & ~1);
        p.tiles_w = ceil_div(W, 8);
        p.tiles_h = ceil_div(H, p.ht);
        p.pix_tiles = p.tiles_w * p.tiles_h * B;
        p.ksteps = p.ht / 2;
        p.a_box_bytes = (uint32_t)(8 * p.ht) * 128u;
        p.x_box_bytes = (uint32_t)(10 * (p.ht + 2)) * 128u;
        p.x_box_stride = (p.x_box_bytes + 1023u) / 1024u * 1024u;
        p.b_sbo = 1280;
        p.nch = 1;
        p.taps_per_part = 5;
        p.tap_parts = 2;
    } else {
        const long npix = (long)B * H * W;
        p.pix_tiles = (int)((npix + 127) / 128);
        p.ksteps = 8;
        p.a_box_bytes = 16384;
        p.x_box_bytes = 16384;
        p.x_box_stride = 16384;
        p.b_sbo = 1024;
        p.nch = std::min(4, chunks_total);
        p.taps_per_part = 1;
        p.tap_parts = 1;
    }
    p.stage_bytes = 2 * p.a_box_bytes + (uint32_t)p.nch * p.x_box_stride;
    p.stages = std::max(2, std::min<int>(kWgStages, (int)((200u * 1024u) / p.stage_bytes)));
    p.m_tiles = ceil_div(Cout, 128);
    p.J = 1;
    for (int i = 0; i < p.m_tiles; ++i) {
        const int m0 = i * 128;
        const int g_lo = m0 / p.cout_g, g_hi = std::min(Cout - 1, m0 + 127) / p.cout_g;
        const int lo = (g_lo * p.cin_g) / 64 * 64;
        const int hi = std::min(chunks_total * 64, ((g_hi + 1) * p.cin_g + 63) / 64 * 64);
        p.J = std::max(p.J, ceil_div((hi - lo) / 64, p.nch));
    }
    const int work = p.m_tiles * p.J * p.tap_parts;
    const int num_sms = dd_num_sms();
    int splits = std::max(1, (2 * num_sms + work - 1) / work);
    splits = std::min(splits, std::max(1, p.pix_tiles / 4));
    splits = std::min(splits, p.pix_tiles);
    p.splits = splits;
    p.use_red = (splits > 1 || accumulate) ? 1 : 0;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.taps_per_part * 64 * p.nch)) cols <<= 1;
    p.tmem_cols = cols;

    CUtensorMap tmY, tmX;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (p.halo) {
        cuuint64_t dimsY[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strY[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
        cuuint32_t boxY[4] = {64, 8, (cuuint32_t)p.ht, 1};
        CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dimsY, strY, boxY, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_wgrad: dY tensor map encode failed (CUresult %d)", (int)r);
        cuuint64_t dimsX[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strX[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t boxX[4] = {64, 10, (cuuint32_t)(p.ht + 2), 1};
        r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dimsX, strX, boxX, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_wgrad: X tensor map encode failed (CUresult %d)", (int)r);
    } else {
        const cuuint64_t npix = (cuuint64_t)B * H * W;
        cuuint64_t dimsY[2] = {(cuuint64_t)Cout, npix};
        cuuint64_t strY[1] = {(cuuint64_t)Cout * 2};
        cuuint32_t box[2] = {64, 128};
        CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), dimsY, strY, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_wgrad: dY tensor map encode failed (CUresult %d)", (int)r);
        cuuint64_t dimsX[2] = {(cuuint64_t)Cin, npix};
        cuuint64_t strX[1] = {(cuuint64_t)Cin * 2};
        r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dimsX, strX, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DD_REQUIRE(r == CUDA_SUCCESS, "dd_mpconv_wgrad: X tensor map encode failed (CUresult %d)", (int)r);
    }

    if (p.use_red && !accumulate)
        DD_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * p.taps * p.cin_g * sizeof(float), stream));
    static bool attr_done = false;
    if (!attr_done) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    // at least 120 KB so that one CTA owns an SM (each CTA allocates up to all 512 TMEM columns)
    const size_t smem_bytes = std::max<size_t>((size_t)p.stages * p.stage_bytes + 1024, 120 * 1024);
    const int grid = work * splits;
    if (getenv("DD_DEBUG_CONV"))
        fprintf(stderr, "[wgrad] B%d %dx%d %d->%d k%d g%d: halo %d ht %d pix_tiles %d ksteps %d m_tiles %d J %d nch %d splits %d "
                        "stages %d stage_bytes %u grid %d red %d\n",
                B, H, W, Cin, Cout, ksize, groups, p.halo, p.ht, p.pix_tiles, p.ksteps, p.m_tiles, p.J, p.nch, p.splits,
                p.stages, p.stage_bytes, grid, p.use_red);
    DD_CHECK_CUDA(dd_launch_pdl(conv_wgrad_kernel, dim3(grid), dim3(kWgThreads), smem_bytes, stream, tmY, tmX, p));
    DD_CHECK_LAUNCH();
    return 0;
}
