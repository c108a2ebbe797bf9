// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32) on B200 as a function of the tile width N,
// where the A operand comes from (dense smem tile / 3x3 halo view with a 10-row group pitch / TMEM) and the CTA
// group (1 SM, M = 128; 2 SMs, M = 256).  No loads, no epilogue: one thread issues `reps` x 36 UMMAs
// (9 "taps" x 4 k-steps) into alternating accumulators, commits, waits, and reports SM cycles per UMMA.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dualdiffusion_b200/csrc -I include \
//               -o tools/umma_bench tools/csrc/umma_bench.cu
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

void dd_set_error(const char*, ...) {}
int dd_num_sms() { return 148; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum Mode { kDense = 0, kHalo = 1, kTmemA = 2 };

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// 1-SM variants.  ksteps: k-steps per tap (4 = 64-channel chunk, 2 = 32-channel chunk).
template <int MODE, int KSTEPS>
__global__ void __launch_bounds__(128, 1) umma_rate_1sm(int n, int reps, unsigned long long* out) {
    constexpr int mode = MODE, ksteps = KSTEPS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_fence_init(); }
    ptx::fence_proxy_async_smem();
    if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t a_base = ptx::smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t idesc = ptx::make_idesc_bf16(128, n);
        const uint64_t a0 = ptx::make_kmajor_desc_sw128(a_base, mode == kHalo ? 1280 : 1024);
        const uint64_t b0 = ptx::make_kmajor_desc_sw128(b_base, 1024);
        constexpr uint32_t bblk = (uint32_t)(256 * 128) >> 4;       // B blocks of 256 rows whatever N is: immediates
        const int nacc = n <= 128 ? 2 : 1;
        long long t0 = 0, t1 = 0;
        if (ptx::elect_one()) {
            t0 = clock64();
#pragma unroll 1
            for (int r = 0; r < reps; ++r) {
                const uint32_t d = tmem + (uint32_t)((r % nacc) * n);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint64_t at = a0 + (mode == kHalo ? (uint64_t)(((tap / 3) * 10 + (tap % 3)) * 8) : 0ull);
                    const uint64_t bt = b0 + (uint64_t)((tap & 1) * bblk);
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) {
                        if (mode == kTmemA) umma_ts(d, tmem + 448 + 8 * ks, bt + 2 * ks, idesc, 1u);
                        else ptx::umma_bf16_ss_acc(d, at + 2 * ks, bt + 2 * ks, idesc);
                    }
                }
            }
            ptx::umma_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        t1 = clock64();
        t0 = __shfl_sync(0xffffffffu, t0, __ffs(__ballot_sync(0xffffffffu, t0 != 0)) - 1);
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

// 2-SM variant: M = 256 (128 rows of A per CTA), each CTA holds n/2 rows of B; rank 0 issues.
template <int MODE, int KSTEPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_rate_2sm(int n, int reps, unsigned long long* out) {
    constexpr int mode = MODE, ksteps = KSTEPS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
    const int warp = threadIdx.x >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_fence_init(); }
    ptx::fence_proxy_async_smem();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
    ptx::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && rank == 0) {
        const uint32_t a_base = ptx::smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t idesc = ptx::make_idesc_bf16(256, n);
        const uint64_t a0 = ptx::make_kmajor_desc_sw128(a_base, mode == kHalo ? 1280 : 1024);
        const uint64_t b0 = ptx::make_kmajor_desc_sw128(b_base, 1024);
        constexpr uint32_t bblk = (uint32_t)(128 * 128) >> 4;
        const int nacc = n <= 128 ? 2 : 1;
        long long t0 = 0, t1 = 0;
        if (ptx::elect_one()) {
            t0 = clock64();
#pragma unroll 1
            for (int r = 0; r < reps; ++r) {
                const uint32_t d = tmem + (uint32_t)((r % nacc) * n);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint64_t at = a0 + (mode == kHalo ? (uint64_t)(((tap / 3) * 10 + (tap % 3)) * 8) : 0ull);
                    const uint64_t bt = b0 + (uint64_t)((tap & 1) * bblk);
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) umma_ss2(d, at + 2 * ks, bt + 2 * ks, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        t1 = clock64();
        t0 = __shfl_sync(0xffffffffu, t0, __ffs(__ballot_sync(0xffffffffu, t0 != 0)) - 1);
        if ((threadIdx.x & 31) == 0) out[blockIdx.x / 2] = (unsigned long long)(t1 - t0);
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        ptx::tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main() {
    const int grid = 148, reps = 200;
    unsigned long long* d_out;
    CK(cudaMalloc(&d_out, grid * sizeof(unsigned long long)));
    typedef void (*KFn)(int, int, unsigned long long*);
    KFn k1[3][2] = {{umma_rate_1sm<0, 4>, umma_rate_1sm<0, 2>}, {umma_rate_1sm<1, 4>, umma_rate_1sm<1, 2>}, {umma_rate_1sm<2, 4>, umma_rate_1sm<2, 2>}};
    KFn k2[2][2] = {{umma_rate_2sm<0, 4>, umma_rate_2sm<0, 2>}, {umma_rate_2sm<1, 4>, umma_rate_2sm<1, 2>}};
    for (auto& r : k1) for (auto f : r) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (auto& r : k2) for (auto f : r) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const char* names[] = {"SS dense (SBO 1024)", "SS halo  (SBO 1280)", "TS (A in TMEM)     "};
    const int ns[] = {32, 64, 96, 128, 192, 256};
    std::vector<unsigned long long> h(grid);
    printf("cycles per UMMA (K = 16), median over CTAs; floor at full rate = N/2 per SM\n");
    for (int two = 0; two < 2; ++two)
        for (int mode = 0; mode < (two ? 2 : 3); ++mode)
            for (int ksteps : {4, 2})
                for (int n : ns) {
                    for (int g : {8, 148}) {
                        const int ctas = two ? (g / 2) * 2 : g;
                        for (int it = 0; it < 2; ++it) {
                            if (two) k2[mode][ksteps == 2]<<<ctas, 128, 100 * 1024>>>(n, reps, d_out);
                            else k1[mode][ksteps == 2]<<<ctas, 128, 100 * 1024>>>(n, reps, d_out);
                            CK(cudaGetLastError());
                            CK(cudaDeviceSynchronize());
                        }
                        const int nout = two ? ctas / 2 : ctas;
                        CK(cudaMemcpy(h.data(), d_out, nout * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                        std::sort(h.begin(), h.begin() + nout);
                        const double cyc = (double)h[nout / 2] / (reps * 9.0 * ksteps);
                        const double macs = (two ? 256.0 : 128.0) * n * 16.0 / cyc / (two ? 2.0 : 1.0);
                        printf("%s %s ksteps %d N %3d grid %3d: %7.1f cyc/UMMA  %7.0f MAC/clk/SM  (max %.1f)\n", two ? "2SM M=256" : "1SM M=128",
                               names[mode], ksteps, n, ctas, cyc, macs, (double)h[nout - 1] / (reps * 9.0 * ksteps));
                    }
                }
    return 0;
}
