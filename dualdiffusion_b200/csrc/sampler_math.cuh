// Index maps / per-element arithmetic of csrc/sampler.cu, free of CUDA-only constructs so that the CPU test harness
// (tests/csrc/sampler_host_check.cpp, built with g++) checks the very same functions against torch.roll / torch.cat /
// the reference's mp_sum.  Reference: pipelines/dual_diffusion_pipeline.py:638-640, :651-656, :729-732.
#pragma once

#ifdef __CUDACC__
#define DD_SHD __host__ __device__ __forceinline__
#else
#define DD_SHD inline
#endif

// column of x that lands in column j of cat(roll(x, shift)[..., -pad:], roll(x, shift), roll(x, shift)[..., :pad])
DD_SHD int roll_pad_src(int j, int W, int shift, int pad) {
    int src = (j - pad - shift) % W;
    return src < 0 ? src + W : src;
}

// column of the padded tensor that lands in column i of roll(xp[..., pad:-pad], -shift)
DD_SHD int crop_unroll_src(int i, int W, int shift, int pad) { return pad + (i + shift) % W; }

// noise[:, ::2] = noise[:, 1::2]: even channels read their odd neighbour
DD_SHD int stereo_src_channel(int c) { return c | 1; }

// mp_sum(a, b, t) with a float t (modules/mp_tools.py:273-279): torch.lerp's two-sided formula, times 1/sqrt((1-t)^2+t^2)
DD_SHD float mp_sum_elem(float a, float b, float t, float inv_norm) {
    const float l = (t < 0.5f) ? a + t * (b - a) : b - (b - a) * (1.f - t);
    return l * inv_norm;
}
