// Body of one row of dd_optim_step_batched (csrc/optim.cu), included textually by the one-CTA-per-row kernels
// (DD_ROW_EXIT = return) and by the persistent experiment (DD_ROW_EXIT = continue) so that both compile the same source.
// Expects in scope: descs, lo, unit, h, clip_coef; macros DD_ROW_TID / DD_ROW_NT (the thread's index among the DD_ROW_NT
// threads that share the row) and DD_ROW_SUM(x) (sum over them).
    const dd_optim_desc d = descs[lo];
    const int r = unit - d.row_begin;
    if (r >= d.rows) DD_ROW_EXIT;
    const size_t base = (size_t)r * d.row_len;
    const long long tail = d.numel - (long long)base;
    const int f = (int)(tail < (long long)d.row_len ? tail : (long long)d.row_len);
    // g == NULL: a parameter that received no gradient this step -- torch's AdamW skips it, EMA_Manager.update() and
    // normalize_weights() still cover it (ema.py:292, mp_tools.py:375-378)
    const bool has_g = d.g != nullptr;
    float* p = d.p + base;
    const float* g = has_g ? d.g + base : nullptr;
    float* m = has_g ? d.m + base : nullptr;
    float* v = has_g ? d.v + base : nullptr;
    const float coef = clip_coef ? clip_coef[1] : 1.f;

    // 128-bit path: every stream of this row 16-byte aligned, all EMA copies fp32
    uintptr_t align = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v) | (uintptr_t)(f & 3) * 4u;
    bool any_f64 = false;
#pragma unroll
    for (int k = 0; k < DD_OPTIM_MAX_EMA; ++k)
        if (k < h.n_ema && d.ema[k] != nullptr) {
            any_f64 |= h.ema_is_f64[k] != 0;
            align |= reinterpret_cast<uintptr_t>(static_cast<float*>(d.ema[k]) + base);
        }
    const bool vec = !any_f64 && (align & 15u) == 0;

    float ss = 0.f;
    if (vec) {
        const int f4 = f >> 2;
        for (int i = DD_ROW_TID; i < f4; i += DD_ROW_NT) {
            float4 p4 = reinterpret_cast<float4*>(p)[i];
            if (has_g) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
                float4 m4 = reinterpret_cast<float4*>(m)[i];
                float4 v4 = reinterpret_cast<float4*>(v)[i];
                p4.x = adamw_elem(p4.x, g4.x * coef, m4.x, v4.x, h);
                p4.y = adamw_elem(p4.y, g4.y * coef, m4.y, v4.y, h);
                p4.z = adamw_elem(p4.z, g4.z * coef, m4.z, v4.z, h);
                p4.w = adamw_elem(p4.w, g4.w * coef, m4.w, v4.w, h);
                reinterpret_cast<float4*>(m)[i] = m4;
                reinterpret_cast<float4*>(v)[i] = v4;
            }
#pragma unroll
            for (int k = 0; k < DD_OPTIM_MAX_EMA; ++k) {
                if (k < h.n_ema && d.ema[k] != nullptr) {
                    float4* e = reinterpret_cast<float4*>(static_cast<float*>(d.ema[k]) + base) + i;
                    float4 e4 = *e;
                    p4.x = ema_elem(p4.x, e4.x, h.ema_w[k], h.fb_w[k]);
                    p4.y = ema_elem(p4.y, e4.y, h.ema_w[k], h.fb_w[k]);
                    p4.z = ema_elem(p4.z, e4.z, h.ema_w[k], h.fb_w[k]);
                    p4.w = ema_elem(p4.w, e4.w, h.ema_w[k], h.fb_w[k]);
                    *e = e4;
                }
            }
            reinterpret_cast<float4*>(p)[i] = p4;
            ss += p4.x * p4.x + p4.y * p4.y + p4.z * p4.z + p4.w * p4.w;
        }
    } else {
        for (int i = DD_ROW_TID; i < f; i += DD_ROW_NT) {
            float pi = p[i];
            if (has_g) {
                float mi = m[i], vi = v[i];
                pi = adamw_elem(pi, g[i] * coef, mi, vi, h);
                m[i] = mi;
                v[i] = vi;
            }
#pragma unroll
            for (int k = 0; k < DD_OPTIM_MAX_EMA; ++k) {
                if (k < h.n_ema && d.ema[k] != nullptr) {
                    if (h.ema_is_f64[k]) {
                        double* e = static_cast<double*>(d.ema[k]) + base + i;
                        double ei = *e;
                        pi = ema_elem(pi, ei, h.ema_w64[k], h.fb_w[k]);
                        *e = ei;
                    } else {
                        float* e = static_cast<float*>(d.ema[k]) + base + i;
                        float ei = *e;
                        pi = ema_elem(pi, ei, h.ema_w[k], h.fb_w[k]);
                        *e = ei;
                    }
                }
            }
            p[i] = pi;
            ss += pi * pi;
        }
    }
    if (!d.normalize) DD_ROW_EXIT;                // uniform over the CTA (descriptor field)
    ss = DD_ROW_SUM(ss);
    const float inv = row_inv_norm(ss, f);
    if (vec) {
        const int f4 = f >> 2;
        for (int i = DD_ROW_TID; i < f4; i += DD_ROW_NT) {
            float4 p4 = reinterpret_cast<float4*>(p)[i];
            p4.x *= inv; p4.y *= inv; p4.z *= inv; p4.w *= inv;
            reinterpret_cast<float4*>(p)[i] = p4;
        }
    } else {
        for (int i = DD_ROW_TID; i < f; i += DD_ROW_NT) p[i] *= inv;
    }
