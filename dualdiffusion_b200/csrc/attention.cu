// Attention of the EDM2 UNet blocks (N = H*W tokens, head_dim 64; sized for the small sequences of the latent: K / V of
// up to 640 tokens stay resident in shared memory, longer sequences stream through in 512-key chunks):
// reference modules/unets/unet_edm2_b4.py:137-151.
//
// Per (batch, head, 128-query tile) CTA: K and V of the head are cosine-normalised (mp_tools.normalize over
// the head channels, eps 1e-4) while they are staged into shared memory (row-major, V^T fragments come from
// ldmatrix.trans), then a
// FlashAttention-style online-softmax loop runs QK^T and PV on the tensor cores (mma.sync m16n8k16 bf16,
// fp32 accumulate).  The block's `mp_silu(y * (emb_linear_v(emb) + 1))` (:150-151) is the epilogue.
// Attention is 0.8 % of the UNet FLOPs; the kernel is sized for latency, not peak.
#include "common.cuh"
#include <type_traits>
#include "dualdiffusion_b200.h"

namespace {

constexpr int kD = 64;            // head dim
constexpr int kQTile = 128;       // queries per CTA (8 warps x 16 rows)
constexpr int kAttThreads = 256;
constexpr int kKStride = kD + 8;  // bf16 elements per K/Q smem row (conflict-free fragment loads)
constexpr float kNormEps = 1e-4f;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// Addressing of one attention problem set.  A "sequence" is indexed by (outer, inner) = (blockIdx.z / n_inner,
// blockIdx.z % n_inner); token t of it sits at  base + outer*outer_stride + inner*inner_stride + t*tok_stride
// (element units), head h at + h*64.  Full-image attention: outer = batch item, n_inner = 1.  Axis attention of the
// legacy ddec UNets (unets/old/unet_edm2_ddec_mdct_b3.py:146-163) folds the reference's
// permute(0,2,4,1,3) -> reshape(b*z*w, ...) -> SDPA -> reshape -> permute(0,3,1,4,2) into these strides: nothing is
// physically transposed, tokens along H are simply W*C elements apart in the channels_last_3d tensor.
struct AttnLayout {
    long q_tok, q_outer, q_inner;      // q and k share strides (k = q pointer + k_offset)
    long v_tok, v_outer, v_inner;
    long o_tok, o_outer, o_inner;
    int n_inner;
    int scale_div;                     // sequences per row of scale_v (scale row = blockIdx.z / scale_div)
};

__global__ void __launch_bounds__(kAttThreads)
attention_kernel(const __nv_bfloat16* __restrict__ q_ptr, const __nv_bfloat16* __restrict__ k_ptr,
                 const __nv_bfloat16* __restrict__ v_ptr, const float* __restrict__ scale_v,
                 __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ raw_out, int N, int heads, int npad, int kcap,
                 const __grid_constant__ AttnLayout lay) {
    extern __shared__ __align__(16) uint8_t smem_att[];
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const int C = heads * kD;
    // K / V are resident `kcap` keys at a time (the whole sequence when it fits: N <= 640); longer sequences stream
    // through the same buffers chunk by chunk while the online softmax carries on
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_att);            // [kcap][kKStride]
    __nv_bfloat16* Vs = Ks + (size_t)kcap * kKStride;                          // [kcap][kKStride]
    __nv_bfloat16* Qs = Vs + (size_t)kcap * kKStride;                          // [kQTile][kKStride]

    const int q0 = blockIdx.x * kQTile, head = blockIdx.y;
    const int outer = blockIdx.z / lay.n_inner, inner = blockIdx.z - outer * lay.n_inner;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = tid & 7;          // which 8-channel slice of the head vector this lane loads
    const int tok_in_pass = tid >> 3; // 32 tokens per pass

    const size_t q_off = (size_t)outer * lay.q_outer + (size_t)inner * lay.q_inner + head * kD + sub * 8;
    const __nv_bfloat16* q_base = q_ptr + q_off;
    const __nv_bfloat16* k_base = k_ptr + q_off;
    const __nv_bfloat16* v_base = v_ptr + (size_t)outer * lay.v_outer + (size_t)inner * lay.v_inner + head * kD + sub * 8;

    // Staging is latency-bound: every load a thread needs for K, V and Q of up to 384 keys is issued before the first one is
    // consumed (one memory round trip instead of one per tensor and per 6 tokens), then normalised and stored.
    constexpr int kPass = 12;          // 32 tokens per pass: 384 keys per chunk
    auto load_rows = [&](auto& q, const __nv_bfloat16* base, size_t token_stride, int first_token, int r0, int n_rows,
                         int row_limit) {
#pragma unroll
        for (int u = 0; u < (int)(sizeof(q) / sizeof(uint4)); ++u) {
            const int r = r0 + u * 32 + tok_in_pass;
            q[u] = make_uint4(0, 0, 0, 0);
            if (r < n_rows && first_token + r < row_limit)
                q[u] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(first_token + r) * token_stride));
        }
    };
    auto norm_store_rows = [&](const auto& q, int r0, int n_rows, __nv_bfloat16* dst) {
#pragma unroll
        for (int u = 0; u < (int)(sizeof(q) / sizeof(uint4)); ++u) {
            const int r = r0 + u * 32 + tok_in_pass;
            if (r0 + u * 32 >= n_rows) break;           // warp-uniform: whole pass beyond the tile
            const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
            float f[8];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 t2 = unpack_bf16x2(w[j]);
                f[2 * j] = t2.x; f[2 * j + 1] = t2.y;
                ss += t2.x * t2.x + t2.y * t2.y;
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            const float inv = 1.f / (kNormEps + sqrtf(ss) * 0.125f);   // ||x|| / sqrt(64)
            if (r < n_rows)
                *reinterpret_cast<uint4*>(dst + (size_t)r * kKStride + sub * 8) =
                    make_uint4(pack_bf16x2(f[0] * inv, f[1] * inv), pack_bf16x2(f[2] * inv, f[3] * inv),
                               pack_bf16x2(f[4] * inv, f[5] * inv), pack_bf16x2(f[6] * inv, f[7] * inv));
        }
    };
    // One chunk of keys (and, with the first chunk, the query tile) into shared memory, PASS x 32 rows per batch of loads.
    auto stage_chunk = [&](auto pass_tag, int cbase, int cn, bool with_q) {
        constexpr int PASS = decltype(pass_tag)::value;
        for (int c0 = 0; c0 < cn; c0 += 32 * PASS) {
            uint4 kq[PASS], vq[PASS], qq[kQTile / 32];
            load_rows(kq, k_base, (size_t)lay.q_tok, cbase, c0, cn, N);
            load_rows(vq, v_base, (size_t)lay.v_tok, cbase, c0, cn, N);
            if (with_q && c0 == 0) load_rows(qq, q_base, (size_t)lay.q_tok, q0, 0, kQTile, N);
            norm_store_rows(kq, c0, cn, Ks);
            norm_store_rows(vq, c0, cn, Vs);
            if (with_q && c0 == 0) norm_store_rows(qq, 0, kQTile, Qs);
        }
    };
    // first (usually only) chunk: staged before the accumulators exist, so its 28 vectors per thread cost no spills
    stage_chunk(std::integral_constant<int, kPass>{}, 0, min(kcap, npad), true);
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16;
    uint32_t qa[4][4];                           // Q fragments for the 4 k-steps over head_dim
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const __nv_bfloat16* p0 = Qs + (size_t)(row0 + g) * kKStride + kk * 16 + 2 * t;
        const __nv_bfloat16* p1 = p0 + 8 * kKStride;
        qa[kk][0] = *reinterpret_cast<const uint32_t*>(p0);
        qa[kk][1] = *reinterpret_cast<const uint32_t*>(p1);
        qa[kk][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        qa[kk][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
    }
    // per-channel output scale of this thread's 16 head channels: requested now, consumed in the epilogue
    float scv[16];
    {
        const float* sc = scale_v ? scale_v + (size_t)(blockIdx.z / lay.scale_div) * C + head * kD : nullptr;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            scv[2 * n] = sc ? __ldg(sc + n * 8 + 2 * t) : 1.f;
            scv[2 * n + 1] = sc ? __ldg(sc + n * 8 + 2 * t + 1) : 1.f;
        }
    }
    const float sl2 = 0.125f * 1.44269504089f;   // 1/sqrt(64) * log2(e)
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }

    for (int cbase = 0; cbase < npad; cbase += kcap) {
    const int cn = min(kcap, npad - cbase);          // keys of this chunk (a multiple of 64)
    if (cbase > 0) {                                 // sequences beyond 640 tokens: next 512 keys through the same buffers
        __syncthreads();                             // every warp is done with the previous chunk
        stage_chunk(std::integral_constant<int, 2>{}, cbase, cn, false);
        __syncthreads();
    }

    for (int kb = 0; kb < cn; kb += 64) {
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            const __nv_bfloat16* kp = Ks + (size_t)(kb + n * 8 + g) * kKStride + 2 * t;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kp + kk * 16);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kp + kk * 16 + 8);
                mma_bf16_16816(s[n], qa[kk], b0, b1);
            }
        }
        // mask padded keys, block row max
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int key = cbase + kb + n * 8 + 2 * t;
            if (key >= N) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
            if (key + 1 >= N) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);     // finite: every 64-key block holds >= 1 real key
            corr[r] = fast_exp2((m_run[r] - m_new) * sl2);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) { o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1]; }
        // probabilities -> bf16 A fragments
        uint32_t pa[4][4];
        const float nm0 = -m_run[0] * sl2, nm1 = -m_run[1] * sl2;      // one FFMA + one MUFU per score
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = fast_exp2(fmaf(s[n][0], sl2, nm0)), p1 = fast_exp2(fmaf(s[n][1], sl2, nm0));
            const float p2 = fast_exp2(fmaf(s[n][2], sl2, nm1)), p3 = fast_exp2(fmaf(s[n][3], sl2, nm1));
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        // O += P V : V^T fragments via ldmatrix.trans on the row-major [key][d] tile.
        // lanes 0-7 / 8-15 address keys +0..7 / +8..15 of head-dim block n, lanes 16-31 the same keys of block n+1
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const __nv_bfloat16* vrow = Vs + (size_t)(kb + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kKStride +
                                        (lane >> 4) * 8;
#pragma unroll
            for (int n = 0; n < 8; n += 2) {
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, vrow + n * 8);
                mma_bf16_16816(o[n], pa[kk], vb[0], vb[1]);
                mma_bf16_16816(o[n + 1], pa[kk], vb[2], vb[3]);
            }
        }
    }

    }      // key chunks

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv_l[2] = {1.f / l_run[0], 1.f / l_run[1]};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + row0 + g + r * 8;
        if (q >= N) continue;
        const size_t o_off = (size_t)outer * lay.o_outer + (size_t)inner * lay.o_inner + (size_t)q * lay.o_tok + head * kD;
        __nv_bfloat16* op = out + o_off;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int d = n * 8 + 2 * t;
            const float a0 = o[n][2 * r + 0] * inv_l[r], a1 = o[n][2 * r + 1] * inv_l[r];
            if (raw_out) *reinterpret_cast<uint32_t*>(raw_out + o_off + d) = pack_bf16x2(a0, a1);   // train mode
            const float y0 = a0 * scv[2 * n];
            const float y1 = a1 * scv[2 * n + 1];
            if (out) *reinterpret_cast<uint32_t*>(op + d) = pack_bf16x2(mp_silu_fast(y0), mp_silu_fast(y1));
        }
    }
}

int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const float* scale_v,
                     __nv_bfloat16* out, __nv_bfloat16* raw_out, int n_seq, int N, int heads, const AttnLayout& lay,
                     cudaStream_t stream) {
    const int npad = ceil_div(N, 64) * 64;
    const int kcap = npad <= 640 ? npad : 512;           // resident keys: whole sequence, or 512-key chunks
    const size_t smem = ((size_t)2 * kcap * kKStride + (size_t)kQTile * kKStride) * 2;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        DD_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    DD_REQUIRE(n_seq <= 65535, "dd_attention: %d sequences exceed the grid limit", n_seq);
    const dim3 grid(ceil_div(N, kQTile), heads, n_seq);
    DD_CHECK_CUDA(dd_launch_pdl(attention_kernel, grid, dim3(kAttThreads), smem, stream, q, k, v, scale_v, out, raw_out, N,
                                heads, npad, kcap, lay));
    return 0;
}

}  // namespace

extern "C" int dd_attention(const void* qk, const void* v, const float* scale_v, void* out, int B, int N, int heads,
                            int head_dim, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(qk && v && scale_v && out, "dd_attention: null pointer");
    DD_REQUIRE(head_dim == kD, "dd_attention: head_dim=%d unsupported (64)", head_dim);
    DD_REQUIRE(N > 0 && N <= 65536, "dd_attention: N=%d unsupported (1..65536)", N);
    const long C = (long)heads * kD;
    AttnLayout lay{2 * C, (long)N * 2 * C, 0, C, (long)N * C, 0, C, (long)N * C, 0, 1, 1};
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qk);
    return launch_attention(q, q + C, static_cast<const __nv_bfloat16*>(v), scale_v, static_cast<__nv_bfloat16*>(out),
                            nullptr, B, N, heads, lay, stream);
}

extern "C" int dd_attention_train(const void* qk, const void* v, const float* scale_v, void* out, void* raw_out, int B,
                                  int N, int heads, int head_dim, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(qk && v && scale_v && out && raw_out, "dd_attention_train: null pointer");
    DD_REQUIRE(head_dim == kD, "dd_attention_train: head_dim=%d unsupported (64)", head_dim);
    DD_REQUIRE(N > 0 && N <= 640, "dd_attention_train: N=%d unsupported (1..640: the backward kernels keep the whole sequence resident)", N);
    const long C = (long)heads * kD;
    AttnLayout lay{2 * C, (long)N * 2 * C, 0, C, (long)N * C, 0, C, (long)N * C, 0, 1, 1};
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qk);
    return launch_attention(q, q + C, static_cast<const __nv_bfloat16*>(v), scale_v, static_cast<__nv_bfloat16*>(out),
                            static_cast<__nv_bfloat16*>(raw_out), B, N, heads, lay, stream);
}

extern "C" int dd_attention_qkv(const void* qkv, void* out_raw, int B, int N, int heads, int head_dim, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(qkv && out_raw, "dd_attention_qkv: null pointer");
    DD_REQUIRE(head_dim == kD, "dd_attention_qkv: head_dim=%d unsupported (64)", head_dim);
    DD_REQUIRE(N > 0 && N <= 65536, "dd_attention_qkv: N=%d unsupported (1..65536)", N);
    const long C = (long)heads * kD, C3 = 3 * C;
    AttnLayout lay{C3, (long)N * C3, 0, C3, (long)N * C3, 0, C, (long)N * C, 0, 1, 1};
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
    return launch_attention(q, q + C, q + 2 * C, nullptr, nullptr, static_cast<__nv_bfloat16*>(out_raw), B, N, heads, lay,
                            stream);
}

extern "C" int dd_attention_axis(const void* qkv, void* out, int B, int Z, int H, int W, int heads, int head_dim, int axis,
                                 void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(qkv && out, "dd_attention_axis: null pointer");
    DD_REQUIRE(head_dim == kD, "dd_attention_axis: head_dim=%d unsupported (64)", head_dim);
    DD_REQUIRE(axis == 0 || axis == 1, "dd_attention_axis: axis must be 0 (H) or 1 (W)");
    const long C = (long)heads * kD, C3 = 3 * C;
    const int N = axis == 0 ? H : W;
    DD_REQUIRE(N > 0 && N <= 65536, "dd_attention_axis: %d tokens unsupported (1..65536)", N);
    AttnLayout lay;
    int n_seq;
    if (axis == 0) {      // attend over H; (b, z) outer, w inner -- b3.py:146-148
        lay = AttnLayout{(long)W * C3, (long)H * W * C3, C3, (long)W * C3, (long)H * W * C3, C3,
                         (long)W * C, (long)H * W * C, C, W, 1};
        n_seq = B * Z * W;
    } else {              // attend over W; (b, z, h) outer
        lay = AttnLayout{C3, (long)W * C3, 0, C3, (long)W * C3, 0, C, (long)W * C, 0, 1, 1};
        n_seq = B * Z * H;
    }
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
    return launch_attention(q, q + C, q + 2 * C, nullptr, static_cast<__nv_bfloat16*>(out), nullptr, n_seq, N, heads, lay,
                            stream);
}
