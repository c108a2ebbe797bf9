// Sampler-loop options of DualDiffusionPipeline.diffusion_decode that are pure index maps / elementwise fp32 work on the
// latent (pipelines/dual_diffusion_pipeline.py:637-640 `stereo_fix`, :651-656 and :729-732 `seamless_loop`).  The latent
// is tiny (B*C*H*W ~ 10^5 floats): these kernels exist so that the loop stays free of host round trips and of
// PyTorch compute on the product path, not for bandwidth.  Index maps are bit-exact by construction.
#include "common.cuh"
#include "dualdiffusion_b200.h"
#include "sampler_math.cuh"

namespace {

// out[r][j] = x[r][(j - pad - shift) mod W], j in [0, W + 2 pad):  torch.roll(x, shift, -1) followed by the circular
// padding cat(x[..., -pad:], x, x[..., :pad])  (:653-654).  copies > 1 repeats the result along the leading dimension
// (`.repeat(2,1,1,1)` of :661 / :628).
__global__ void roll_pad_w_kernel(const float* __restrict__ x, float* __restrict__ out, long rows, int W, int shift,
                                  int pad, int copies) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Wp = W + 2 * pad;
    const long total = rows * Wp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / Wp;
        const int j = (int)(i - r * Wp);
        const float v = x[r * W + roll_pad_src(j, W, shift, pad)];
        for (int c = 0; c < copies; ++c) out[(long)c * total + i] = v;
    }
}

// out[r][i] = xp[r][pad + (i + shift) mod W]:  torch.roll(xp[..., pad:-pad], -shift, -1)  (:730-732)
__global__ void crop_unroll_w_kernel(const float* __restrict__ xp, float* __restrict__ out, long rows, int W, int shift,
                                     int pad) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    const int Wp = W + 2 * pad;
    const long total = rows * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / W;
        const int j = (int)(i - r * W);
        out[i] = xp[r * Wp + crop_unroll_src(j, W, shift, pad)];
    }
}

// :638-640: noise[:, ::2] = noise[:, 1::2]; noise = mp_sum(fresh, noise, t) = lerp(fresh, noise, t) / sqrt((1-t)^2 + t^2)
// (mp_tools.py:273-279, torch.lerp's two-sided formula).  noise, fresh, out: fp32 (B, C, hw), C even.
__global__ void stereo_fix_noise_kernel(const float* __restrict__ noise, const float* __restrict__ fresh, float t,
                                        float inv_norm, float* __restrict__ out, int C, long hw, long total) {
    ptx::grid_launch_dependents();      // PDL: the next kernel's launch / prologue overlaps this one
    ptx::grid_dependency_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long bc = i / hw;
        const long pos = i - bc * hw;
        const int c = (int)(bc % C);
        const float b = noise[(bc - c + stereo_src_channel(c)) * hw + pos];
        out[i] = mp_sum_elem(fresh[i], b, t, inv_norm);
    }
}

inline int grid_for_s(long total, int block) {
    const long blocks = (total + block - 1) / block;
    return (int)std::max<long>(1, std::min<long>(blocks, (long)dd_num_sms() * 8));
}

}  // namespace

extern "C" int dd_roll_pad_w(const float* x, float* out, long rows, int W, int shift, int pad, int copies, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(x && out, "dd_roll_pad_w: null pointer");
    DD_REQUIRE(rows > 0 && W > 0 && pad >= 0 && pad <= W && copies >= 1, "dd_roll_pad_w: bad sizes (rows=%ld W=%d pad=%d)",
               rows, W, pad);
    DD_REQUIRE(shift >= 0 && shift < W, "dd_roll_pad_w: shift=%d must lie in [0, W)", shift);
    DD_CHECK_CUDA(dd_launch_pdl(roll_pad_w_kernel, dim3(grid_for_s(rows * (W + 2 * pad), 256)), dim3(256), 0, stream, x, out, rows, W, shift, pad, copies));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_crop_unroll_w(const float* xp, float* out, long rows, int W, int shift, int pad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(xp && out, "dd_crop_unroll_w: null pointer");
    DD_REQUIRE(rows > 0 && W > 0 && pad >= 0 && pad <= W, "dd_crop_unroll_w: bad sizes (rows=%ld W=%d pad=%d)", rows, W, pad);
    DD_REQUIRE(shift >= 0 && shift < W, "dd_crop_unroll_w: shift=%d must lie in [0, W)", shift);
    DD_CHECK_CUDA(dd_launch_pdl(crop_unroll_w_kernel, dim3(grid_for_s(rows * W, 256)), dim3(256), 0, stream, xp, out, rows, W, shift, pad));
    DD_CHECK_LAUNCH();
    return 0;
}

extern "C" int dd_stereo_fix_noise(const float* noise, const float* fresh, float t, float* out, int B, int C, long hw,
                                   void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DD_REQUIRE(noise && fresh && out, "dd_stereo_fix_noise: null pointer");
    DD_REQUIRE(B > 0 && C > 0 && hw > 0, "dd_stereo_fix_noise: bad sizes");
    DD_REQUIRE(C % 2 == 0, "dd_stereo_fix_noise: the channel count must be even (noise[:, ::2] = noise[:, 1::2])");
    const double norm = sqrt((1.0 - (double)t) * (1.0 - (double)t) + (double)t * (double)t);
    const long total = (long)B * C * hw;
    DD_CHECK_CUDA(dd_launch_pdl(stereo_fix_noise_kernel, dim3(grid_for_s(total, 256)), dim3(256), 0, stream, noise, fresh, t, (float)(1.0 / norm), out, C, hw,
                                                                        total));
    DD_CHECK_LAUNCH();
    return 0;
}
