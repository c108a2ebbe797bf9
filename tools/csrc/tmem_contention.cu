// Microbenchmark: tcgen05.ld throughput of W epilogue-like warps WHILE one thread keeps the tensor pipe busy with
// 128 x N x 16 SS-mode UMMAs into a different TMEM column range (B200).  Reports cycles per UMMA and LDTM bytes/clk/SM.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dualdiffusion_b200/csrc -I include \
//               -o tools/tmem_contention tools/csrc/tmem_contention.cu
#include "common.cuh"
#include <cstdlib>
void dd_set_error(const char*, ...) {}
int dd_num_sms() { return 148; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(64 + 512, 1) k(int n, int mma_reps, int ld_reps, int ld_warps, int do_mma, unsigned long long* out, float* sink) {
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
        if (do_mma) {
            const uint32_t a_base = ptx::smem_u32(smem), b_base = a_base + 32 * 1024;
            const uint32_t idesc = ptx::make_idesc_bf16(128, n);
            const uint64_t a0 = ptx::make_kmajor_desc_sw128(a_base, 1024), b0 = ptx::make_kmajor_desc_sw128(b_base, 1024);
            long long t0 = 0;
            if (ptx::elect_one()) {
                t0 = clock64();
#pragma unroll 1
                for (int r = 0; r < mma_reps; ++r) {
#pragma unroll
                    for (int i = 0; i < 12; ++i) ptx::umma_bf16_ss_acc(tmem, a0 + 2 * (i & 3) + 256 * (i >> 2), b0 + 2 * (i & 3), idesc);
                }
                ptx::umma_commit(&bar);
            }
            __syncwarp();
            ptx::mbar_wait(&bar, 0);
            const long long t1 = clock64();
            t0 = __shfl_sync(0xffffffffu, t0, __ffs(__ballot_sync(0xffffffffu, t0 != 0)) - 1);
            if ((threadIdx.x & 31) == 0) out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
        }
    } else if (warp >= 2 && warp < 2 + ld_warps) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        float acc = 0.f;
        const long long t0 = clock64();
        for (int r = 0; r < ld_reps; ++r) {
            uint32_t a[16], b[16], c[16];
            ptx::tmem_ld_32x16(taddr + (r & 3) * 16, a);
            ptx::tmem_ld_32x16(taddr + 64 + (r & 3) * 16, b);
            ptx::tmem_ld_32x16(taddr + 128 + (r & 3) * 16, c);
            ptx::tmem_ld_wait();
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) x ^= a[j] ^ b[j] ^ c[j];
            acc += __uint_as_float(x);
        }
        const long long t1 = clock64();
        if (warp == 2 && (threadIdx.x & 31) == 0) out[blockIdx.x * 2 + 1] = (unsigned long long)(t1 - t0);
        if (acc == 12345.678f) sink[0] = acc;
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    unsigned long long* d_out; float* sink;
    CK(cudaMalloc(&d_out, 148 * 16)); CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int n : {96, 192})
        for (int do_mma : {0, 1})
            for (int w : {4, 8, 16}) {
                const int mma_reps = 3000, ld_reps = n == 96 ? (48000 / w) : (32000 / w);
                unsigned long long h[296];
                for (int it = 0; it < 2; ++it) { k<<<148, 64 + 512, 100 * 1024>>>(n, mma_reps, ld_reps, w, do_mma, d_out, sink); CK(cudaDeviceSynchronize()); }
                CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
                const double ld_cyc = (double)h[1], bytes = (double)ld_reps * w * 3 * 32 * 16 * 4;
                printf("N %3d  MMA %s  ld warps %2d: %6.1f cyc/UMMA   LDTM %6.1f B/clk/SM (%6.1f cycles per 3 x16 loads per warp)\n", n,
                       do_mma ? "on " : "off", w, do_mma ? (double)h[0] / (mma_reps * 12.0) : 0.0, bytes / ld_cyc, ld_cyc / ld_reps);
            }
    return 0;
}
