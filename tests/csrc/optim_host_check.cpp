// TEST INFRASTRUCTURE ONLY (built with g++ by tests/test_oracle_optim.py; never linked into the product library).
// Runs the optimizer-side sweep of dualdiffusion_b200/csrc/optim.cu on HOST memory, sequentially, through the very same
// per-element functions (csrc/optim_math.cuh) and the same descriptor walk (row r of a [rows][row_len] view, last row
// bounded by numel), so that the arithmetic, the C struct layouts and ops.pack_optim_descs can be checked against the
// reference golden without a GPU.  The thread/block decomposition of the CUDA kernels is what the -m gpu tests cover.
#include "optim_math.cuh"

extern "C" void optim_host_grad_norm(const dd_gnorm_desc* descs, int n_descs, float max_norm, float* out2) {
    double total = 0.0;
    for (int d = 0; d < n_descs; ++d) {
        const long long chunks = (descs[d].numel + DD_GNORM_CHUNK - 1) / DD_GNORM_CHUNK;
        for (long long c = 0; c < chunks; ++c) {            // one fp32 partial per chunk, summed in fp64 (as the kernels do)
            const long long lo = c * DD_GNORM_CHUNK;
            const long long hi = lo + DD_GNORM_CHUNK < descs[d].numel ? lo + DD_GNORM_CHUNK : descs[d].numel;
            float ss = 0.f;
            for (long long i = lo; i < hi; ++i) ss += descs[d].g[i] * descs[d].g[i];
            total += (double)ss;
        }
    }
    out2[0] = (float)sqrt(total);
    out2[1] = clip_coef_from_norm(out2[0], max_norm);
}

extern "C" int optim_host_step(const dd_optim_desc* descs, int n_descs, int total_rows, const dd_optim_hyper* hy,
                               float coef) {
    const OptimHyperDev h = make_hyper_dev(*hy);
    int rows_seen = 0;
    for (int di = 0; di < n_descs; ++di) {
        const dd_optim_desc d = descs[di];
        if (d.row_begin != rows_seen) return 1;              // prefix sums must be exclusive and gap-free
        for (int r = 0; r < d.rows; ++r) {
            const size_t base = (size_t)r * d.row_len;
            const long long tail = d.numel - (long long)base;
            const int f = (int)(tail < (long long)d.row_len ? tail : (long long)d.row_len);
            if (f <= 0) return 2;
            float ss = 0.f;
            for (int i = 0; i < f; ++i) {
                float pi = d.p[base + i];
                if (d.g != nullptr) {
                    float mi = d.m[base + i], vi = d.v[base + i];
                    pi = adamw_elem(pi, d.g[base + i] * coef, mi, vi, h);
                    d.m[base + i] = mi;
                    d.v[base + i] = vi;
                }
                for (int k = 0; k < DD_OPTIM_MAX_EMA; ++k) {
                    if (k < h.n_ema && d.ema[k] != nullptr) {
                        if (h.ema_is_f64[k]) {
                            double* e = static_cast<double*>(d.ema[k]) + base + i;
                            pi = ema_elem(pi, *e, h.ema_w64[k], h.fb_w[k]);
                        } else {
                            float* e = static_cast<float*>(d.ema[k]) + base + i;
                            pi = ema_elem(pi, *e, h.ema_w[k], h.fb_w[k]);
                        }
                    }
                }
                d.p[base + i] = pi;
                ss += pi * pi;
            }
            if (d.normalize) {
                const float inv = row_inv_norm(ss, f);
                for (int i = 0; i < f; ++i) d.p[base + i] *= inv;
            }
        }
        rows_seen += d.rows;
    }
    return rows_seen == total_rows ? 0 : 3;
}
