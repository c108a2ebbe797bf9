// Microbenchmark: tcgen05.ld (TMEM -> registers) and SHFL throughput per SM on B200, with 4 / 8 / 12 warps.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dualdiffusion_b200/csrc -I include \
//               -o tools/tmem_bench tools/csrc/tmem_bench.cu
#include "common.cuh"
#include <cstdlib>
void dd_set_error(const char*, ...) {}
int dd_num_sms() { return 148; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// mode 0: tcgen05.ld 32x32b.x32 back to back (one wait per load); 1: two loads in flight per wait; 2: SHFL; 3: x16 loads
__global__ void __launch_bounds__(384, 1) k(int mode, int reps, unsigned long long* out, float* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t taddr = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int r = 0; r < reps; ++r) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(taddr + (r & 15) * 32, v);
            ptx::tmem_ld_wait();
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) x ^= v[j];
            acc += __uint_as_float(x);
        }
    } else if (mode == 1) {
        for (int r = 0; r < reps; r += 2) {
            uint32_t v[32], w[32];
            ptx::tmem_ld_32x32(taddr + (r & 14) * 32, v);
            ptx::tmem_ld_32x32(taddr + ((r & 14) + 1) * 32, w);
            ptx::tmem_ld_wait();
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) x ^= v[j] ^ w[j];
            acc += __uint_as_float(x);
        }
    } else if (mode == 3) {
        for (int r = 0; r < reps; ++r) {
            uint32_t v[16];
            ptx::tmem_ld_32x16(taddr + (r & 31) * 16, v);
            ptx::tmem_ld_wait();
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) x ^= v[j];
            acc += __uint_as_float(x);
        }
    } else {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (float)(threadIdx.x + j);
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j & 7] += __shfl_up_sync(0xffffffffu, x[j & 7], 1);
        }
        acc = x[0] + x[1] + x[2] + x[3] + x[4] + x[5] + x[6] + x[7];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 12345.678f) sink[0] = acc;
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tmem_slot, 512); }
}

int main() {
    unsigned long long* d_out; float* sink;
    CK(cudaMalloc(&d_out, 148 * 8)); CK(cudaMalloc(&sink, 4));
    const int reps = 4096;
    const char* names[] = {"tcgen05.ld 32x32b.x32, 1 in flight", "tcgen05.ld 32x32b.x32, 2 in flight", "shfl.up x32 per rep", "tcgen05.ld 32x32b.x16, 1 in flight"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 12}) {
            unsigned long long h[148];
            for (int it = 0; it < 2; ++it) { k<<<148, warps * 32, 0>>>(mode, reps, d_out, sink); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
            const double cyc = (double)h[0];
            if (mode == 2) printf("%-40s warps %2d: %.2f cycles per warp-SHFL per SM (%.1f per warp)\n", names[mode], warps, cyc / (reps * 32.0 * warps), cyc / (reps * 32.0));
            else {
                const double bytes = (double)reps * warps * 32 * (mode == 3 ? 16 : 32) * 4;
                printf("%-40s warps %2d: %.1f B/clk/SM, %.1f cycles per load per warp\n", names[mode], warps, bytes / cyc, cyc / reps);
            }
        }
    return 0;
}
