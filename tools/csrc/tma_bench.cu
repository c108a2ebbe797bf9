// Microbenchmark: sustained TMA (cp.async.bulk.tensor) delivery rate into shared memory for the box shapes the conv
// kernels use, no consumer.  One elected thread per CTA keeps S boxes in flight over a [B][H][W][C] bf16 tensor.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_bench tools/csrc/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(void* dst, const void* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct P { int tiles_w, tiles_h, B, nchunk, bw, bh, ox, oy, stages, boxes_per_cta_total; uint32_t box_bytes, stride; };

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ P p, int total, int reps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[16];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t0 = (int)((long)blockIdx.x * total / gridDim.x), t1 = (int)((long)(blockIdx.x + 1) * total / gridDim.x);
        int n = 0;
        for (int rep = 0; rep < reps; ++rep)
        for (int t = t0; t < t1; ++t, ++n) {
            const int s = n % p.stages;
            if (n >= p.stages) mbar_wait(&full[s], ((n / p.stages) - 1) & 1);
            int r = t;
            const int kc = r % p.nchunk; r /= p.nchunk;
            const int tw = r % p.tiles_w; r /= p.tiles_w;
            const int th = r % p.tiles_h; const int b = r / p.tiles_h;
            mbar_expect(&full[s], p.box_bytes);
            tma4(smem + (size_t)s * p.stride, &tm, &full[s], kc * 64, tw * p.bw + p.ox, th * p.bh + p.oy, b);
        }
        for (int k = 0; k < p.stages && k < n; ++k) {       // drain
            const int idx = n - 1 - k, s = idx % p.stages;
            mbar_wait(&full[s], (idx / p.stages) & 1);
        }
    }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int B = 2, H = 32, W = 688, C = 512;
    void* x; CK(cudaMalloc(&x, (size_t)B * H * W * C * 2)); CK(cudaMemset(x, 0, (size_t)B * H * W * C * 2));
    void* flush; CK(cudaMalloc(&flush, 256u << 20));
    void* sym; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    PFN enc = (PFN)sym;
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Case { const char* name; int bw, bh, boxw, boxh, ox, oy; } cases[] = {
        {"halo 10x18 (3x3 conv A tile)", 8, 16, 10, 18, -1, -1},
        {"plain 8x16 (no halo)", 8, 16, 8, 16, 0, 0},
        {"row 128x1 (1x1 conv A tile)", 128, 1, 128, 1, 0, 0},
        {"row 16x8", 16, 8, 16, 8, 0, 0},
        {"halo 18x10 (wide)", 16, 8, 18, 10, -1, -1},
    };
    for (auto& cs : cases) {
        for (int stages : {2, 6}) {
            for (int grid : {8, 148}) {
                for (int cold = 0; cold < 2; ++cold) {
                    CUtensorMap tm;
                    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
                    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
                    cuuint32_t box[4] = {64, (cuuint32_t)cs.boxw, (cuuint32_t)cs.boxh, 1};
                    cuuint32_t es[4] = {1, 1, 1, 1};
                    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
                    P p{};
                    p.tiles_w = (W + cs.bw - 1) / cs.bw; p.tiles_h = (H + cs.bh - 1) / cs.bh; p.B = B; p.nchunk = C / 64;
                    p.bw = cs.bw; p.bh = cs.bh; p.ox = cs.ox; p.oy = cs.oy; p.stages = stages;
                    p.box_bytes = (uint32_t)cs.boxw * cs.boxh * 128u;
                    p.stride = (p.box_bytes + 1023u) / 1024u * 1024u;
                    const int total = p.tiles_w * p.tiles_h * B * p.nchunk;
                    const int total_run = grid == 148 ? total : total * 8 / 148;       // same work per CTA
                    const size_t smem = (size_t)stages * p.stride + 1024;
                    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    const int reps = 16;
                    float best = 1e9f;
                    for (int it = 0; it < 4; ++it) {
                        if (cold) CK(cudaMemsetAsync(flush, it, 256u << 20));
                        CK(cudaEventRecord(e0));
                        tma_kernel<<<grid, 128, smem>>>(tm, p, total_run, reps);
                        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (it > 0 && ms < best) best = ms;
                    }
                    CK(cudaGetLastError());
                    const double bytes = (double)total_run * reps * p.box_bytes;
                    const double boxes_per_cta = (double)total_run * reps / grid;
                    printf("%-30s stages %d grid %3d %s: %7.1f us  %7.1f GB/s  %6.1f B/clk/SM  %6.0f cyc/box  %5.2f cyc/row\n", cs.name, stages,
                           grid, cold ? "cold" : "warm", best * 1e3, bytes / best / 1e6, bytes / grid / (best * 1e-3 * 1.965e9),
                           best * 1e-3 * 1.965e9 / boxes_per_cta, best * 1e-3 * 1.965e9 / boxes_per_cta / (cs.boxw * cs.boxh));
                }
            }
        }
    }
    return 0;
}
