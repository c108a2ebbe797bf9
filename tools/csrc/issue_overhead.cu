// Microbenchmark: what the UMMA-issuing thread pays per tile for its control path on B200.  One warp issues groups of
// 12 SS-mode 128 x 96 x 16 UMMAs; between groups it optionally (1) commits to an mbarrier, (2) waits on two mbarriers that
// are already complete and fences, (3) uses the elect_one / __syncwarp structure of the conv kernels.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dualdiffusion_b200/csrc -I include \
//               -o tools/issue_overhead tools/csrc/issue_overhead.cu
#include "common.cuh"
#include <cstdlib>
void dd_set_error(const char*, ...) {}
int dd_num_sms() { return 148; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n, int reps_total, unsigned long long* out) {
    int reps = reps_total;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar, dummy[4], ready[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&dummy[i], 1);
        ptx::mbar_init(&ready[0], 1); ptx::mbar_init(&ready[1], 1);
        ptx::mbar_fence_init();
    }
    ptx::fence_proxy_async_smem();
    if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1 || (MODE == 6 && warp == 2)) {
        if (MODE == 6) reps /= 2;
        const uint32_t a_base = ptx::smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t idesc = ptx::make_idesc_bf16(128, n);
        const uint64_t a0 = ptx::make_kmajor_desc_sw128(a_base, 1024), b0 = ptx::make_kmajor_desc_sw128(b_base, 1024);
        const long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
            if (MODE == 2 || MODE == 3 || MODE == 6) {            // two waits on barriers whose awaited phase is already complete + fences
                ptx::mbar_wait(&ready[0], 1);
                ptx::tcgen05_fence_after();
                ptx::mbar_wait(&ready[1], 1);
                ptx::tcgen05_fence_after();
            }
            if (MODE == 4) { ptx::mbar_wait(&ready[0], 1); ptx::mbar_wait(&ready[1], 1); }      // waits only
            if (MODE == 5) { ptx::tcgen05_fence_after(); ptx::tcgen05_fence_after(); }          // fences only
            bool peek0 = true, peek1 = true;
            if (MODE == 7) { peek0 = ptx::mbar_test_wait(&ready[0], 1); peek1 = ptx::mbar_test_wait(&ready[1], 1); }   // probe early, consume late
            if (ptx::elect_one()) {
#pragma unroll
                for (int i = 0; i < 12; ++i) ptx::umma_bf16_ss_acc(tmem + (r & 1) * 96 + (warp - 1) * 192, a0 + 2 * (i & 3) + 256 * (i >> 2), b0 + 2 * (i & 3), idesc);
                if (MODE >= 1) { ptx::umma_commit(&dummy[r & 1]); if (MODE == 3 || MODE == 6 || MODE == 7) ptx::umma_commit(&dummy[2]); }
            }
            __syncwarp();
            if (MODE == 7) {
                if (!peek0) ptx::mbar_wait(&ready[0], 1);
                if (!peek1) ptx::mbar_wait(&ready[1], 1);
                ptx::tcgen05_fence_after();
            }
        }
        if (ptx::elect_one()) ptx::umma_commit(warp == 1 ? &bar : &dummy[3]);
        __syncwarp();
        ptx::mbar_wait(warp == 1 ? &bar : &dummy[3], 0);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && warp == 1) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
    unsigned long long* d_out;
    CK(cudaMalloc(&d_out, 148 * 8));
    const int reps = 2000;
    typedef void (*KFn)(int, int, unsigned long long*);
    KFn ks[8] = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>, k<6>, k<7>};
    const char* names[8] = {"issue only", "+ 1 commit per group", "+ 2 ready waits + fences", "+ second commit", "1 commit + 2 waits (no fence)",
                            "1 commit + 2 fences (no wait)", "as row 4, two issuing warps", "as row 4, waits probed before the UMMAs"};
    for (int n : {96, 192})
        for (int m = 0; m < 8; ++m) {
            CK(cudaFuncSetAttribute(ks[m], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            unsigned long long h[148];
            for (int it = 0; it < 2; ++it) { ks[m]<<<148, 128, 100 * 1024>>>(n, reps, d_out); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
            printf("N %3d %-28s: %7.1f cycles per group of 12 UMMAs (pipe floor %d)\n", n, names[m], (double)h[0] / reps, 12 * (n == 96 ? 56 : 96));
        }
    return 0;
}
