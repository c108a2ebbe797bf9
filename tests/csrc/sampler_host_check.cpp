// TEST INFRASTRUCTURE ONLY (built with g++ by tests/test_sampler_options.py; never linked into the product library).
// Sequential host loops over the same index / element functions csrc/sampler.cu compiles (csrc/sampler_math.cuh) and the
// same flat indexing, so the maps can be compared bit for bit with torch.roll / torch.cat on the CPU.
#include <math.h>
#include "sampler_math.cuh"

extern "C" void host_roll_pad_w(const float* x, float* out, long rows, int W, int shift, int pad, int copies) {
    const int Wp = W + 2 * pad;
    const long total = rows * Wp;
    for (long i = 0; i < total; ++i) {
        const long r = i / Wp;
        const int j = (int)(i - r * Wp);
        const float v = x[r * W + roll_pad_src(j, W, shift, pad)];
        for (int c = 0; c < copies; ++c) out[(long)c * total + i] = v;
    }
}

extern "C" void host_crop_unroll_w(const float* xp, float* out, long rows, int W, int shift, int pad) {
    const int Wp = W + 2 * pad;
    for (long i = 0; i < rows * W; ++i) {
        const long r = i / W;
        const int j = (int)(i - r * W);
        out[i] = xp[r * Wp + crop_unroll_src(j, W, shift, pad)];
    }
}

extern "C" void host_stereo_fix_noise(const float* noise, const float* fresh, float t, float* out, int B, int C, long hw) {
    const float inv_norm = (float)(1.0 / sqrt((1.0 - (double)t) * (1.0 - (double)t) + (double)t * (double)t));
    const long total = (long)B * C * hw;
    for (long i = 0; i < total; ++i) {
        const long bc = i / hw;
        const long pos = i - bc * hw;
        const int c = (int)(bc % C);
        out[i] = mp_sum_elem(fresh[i], noise[(bc - c + stereo_src_channel(c)) * hw + pos], t, inv_norm);
    }
}
