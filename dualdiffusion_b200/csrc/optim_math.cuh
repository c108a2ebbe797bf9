// Per-element arithmetic of the optimizer-side sweep (csrc/optim.cu), kept free of CUDA-only constructs so that the
// CPU test harness (tests/csrc/optim_host_check.cpp, built with g++) exercises the very same functions against the
// reference golden.  Reference: torch/optim/adam.py (_single_tensor_adam, decoupled decay) as called from
// training/trainer.py:461-473,1062; training/ema.py:300-313; modules/mp_tools.py:42-49.
#pragma once
#include <math.h>
#include "dualdiffusion_b200.h"

#ifdef __CUDACC__
#define DD_HD __host__ __device__ __forceinline__
#else
#define DD_HD inline
#endif

constexpr float kNormEps = 1e-4f;   // modules/mp_tools.py:43

// torch.lerp's two-sided formula (ATen/native/Lerp.h): exact at both ends of the weight range
template <typename T>
DD_HD T lerp_t(T a, T b, T w) {
    return (w < T(0.5)) ? a + w * (b - a) : b - (b - a) * (T(1) - w);
}

struct OptimHyperDev {
    float decay, w_m, beta2, w_v, eps, step_size, bc2_sqrt;
    int use_decay, n_ema;
    float ema_w[DD_OPTIM_MAX_EMA];        // 1 - beta_k
    double ema_w64[DD_OPTIM_MAX_EMA];     // the same weight for fp64 EMA copies (EMA_Config.use_float64, ema.py:198)
    float fb_w[DD_OPTIM_MAX_EMA];         // 1 - feedback_beta_k, < 0: no feedback
    int ema_is_f64[DD_OPTIM_MAX_EMA];
};

// torch.optim.AdamW, single-tensor formulation (torch/optim/adam.py _single_tensor_adam with decoupled decay):
//   p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
DD_HD float adamw_elem(float pi, float gi, float& mi, float& vi, const OptimHyperDev& h) {
    if (h.use_decay) pi *= h.decay;
    mi = lerp_t(mi, gi, h.w_m);
    vi = h.beta2 * vi + h.w_v * gi * gi;
    const float denom = sqrtf(vi) / h.bc2_sqrt + h.eps;
    return pi - h.step_size * (mi / denom);
}

// ema.py:307-313: ema_k = lerp(ema_k, p, 1-beta_k); if feedback: p = lerp(p, ema_k, 1-feedback_beta_k), in config order
template <typename E>
DD_HD float ema_elem(float pi, E& ei, E w, float fb_w) {
    ei = lerp_t(ei, (E)pi, w);
    if (fb_w >= 0.f) pi = lerp_t(pi, (float)ei, fb_w);
    return pi;
}


// torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1 (NaN propagates, as torch.clamp does);
// max_norm <= 0 only reports the norm
DD_HD float clip_coef_from_norm(float norm, float max_norm) {
    if (!(max_norm > 0.f)) return 1.f;
    const float coef = max_norm / (norm + 1e-6f);
    return coef > 1.f ? 1.f : coef;           // the comparison is false for NaN: stays NaN
}

// mp_tools.py:42-49 for one weight row: 1 / (eps + ||row|| * sqrt(1 / fan_in))
DD_HD float row_inv_norm(float sum_sq, int fan_in) { return 1.f / (kNormEps + sqrtf(sum_sq) / sqrtf((float)fan_in)); }

// host: dd_optim_hyper (fp64 hyper-parameters of one launch) -> the fp32 constants the kernel uses
inline OptimHyperDev make_hyper_dev(const dd_optim_hyper& hy) {
    OptimHyperDev h{};
    h.use_decay = hy.weight_decay != 0.0;
    h.decay = (float)(1.0 - hy.lr * hy.weight_decay);
    h.w_m = (float)(1.0 - hy.beta1);
    h.beta2 = (float)hy.beta2;
    h.w_v = (float)(1.0 - hy.beta2);
    h.eps = (float)hy.eps;
    h.step_size = (float)(hy.lr / hy.bias_correction1);
    h.bc2_sqrt = (float)sqrt(hy.bias_correction2);
    h.n_ema = hy.n_ema;
    for (int k = 0; k < DD_OPTIM_MAX_EMA; ++k) {
        h.ema_w[k] = (float)(1.0 - hy.ema_beta[k]);
        h.ema_w64[k] = 1.0 - hy.ema_beta[k];
        h.fb_w[k] = hy.feedback_beta[k] >= 0.0 ? (float)(1.0 - hy.feedback_beta[k]) : -1.f;
        h.ema_is_f64[k] = hy.ema_is_f64[k];
    }
    return h;
}
