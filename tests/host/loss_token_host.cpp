// Host build of the per-token loss arithmetic (skyrl_b200/csrc/loss_token.cuh) for the CPU test suite.
// TEST INFRASTRUCTURE ONLY: compiled by tests/test_loss_token_host_cpu.py with g++, never linked into libskyb200.so.
//
// The driver below walks a [B, A] batch the way loss_main_kernel / loss_flat_kernel do (normalisers from the mask,
// one TokenCtx per row) and calls the SAME loss_token<> source the kernels compile; it exists so that a slip in
// that arithmetic shows up in `pytest -m "not gpu"`, before any GPU time is spent.
#include <cmath>
#include <cstdint>

#include "../../skyrl_b200/csrc/loss_token.cuh"

using namespace skyb200;

// kType: the template argument of loss_token (-1 run-time typed, -2 clipped family, >= 0 fixed type)
template <int kType>
static void run(const skyb200_loss_params& p, int B, int A, const float* lp, const float* olp, const float* adv,
                const float* mask, const float* base, const float* tis, float* d_lp, double* out) {
    double total_m = 0.0;
    for (int i = 0; i < B * A; ++i) total_m += mask ? mask[i] : 1.f;
    const float inv_total = 1.f / fmaxf(static_cast<float>(total_m), 1.f);
    const float Bf = static_cast<float>(B);
    TokenCtx cx;
    cx.gspo = (p.loss_type == SKYB200_LOSS_GSPO);
    cx.dual = (p.loss_type == SKYB200_LOSS_DUAL_CLIP);
    cx.lo = 1.f - p.eps_clip_low;
    cx.hi = 1.f + p.eps_clip_high;
    cx.seq_ratio = 1.f;
    cx.seq_pass = 1.f;
    cx.clip_c = p.clip_ratio_c;
    cx.kl_type = p.kl_type;
    cx.loss_type = p.loss_type;
    cx.tau_pos = p.sapo_tau_pos;
    cx.tau_neg = p.sapo_tau_neg;
    cx.has_mask = (mask != nullptr);
    cx.has_base = (base != nullptr);
    cx.d_ent_val = 0.f;
    double s_loss = 0.0, s_clip = 0.0, s_kl = 0.0, s_ratio = 0.0;
    float r_max = -INFINITY, r_min = INFINITY;
    for (int b = 0; b < B; ++b) {
        double row_md = 0.0, lw = 0.0;
        for (int t = 0; t < A; ++t) {
            const float m = mask ? mask[b * A + t] : 1.f;
            row_md += m;
            lw += static_cast<double>(lp[b * A + t] - olp[b * A + t]) * m;
        }
        const float row_m = static_cast<float>(row_md);
        const float inv_row = 1.f / fmaxf(row_m, 1.f);
        if (p.reduction == SKYB200_REDUCE_TOKEN_MEAN)
            cx.w_scale = inv_total;
        else if (p.reduction == SKYB200_REDUCE_SEQUENCE_MEAN)
            cx.w_scale = inv_row / Bf;
        else if (p.reduction == SKYB200_REDUCE_SUM)
            cx.w_scale = 1.f;
        else
            cx.w_scale = 1.f / (p.max_seq_len * Bf);
        cx.kl_scale = inv_row / Bf;
        cx.kl_grad_scale = p.kl_has_grad ? p.kl_loss_coef * cx.kl_scale : 0.f;
        if (cx.gspo) {
            const float row_lw = static_cast<float>(lw) / fmaxf(row_m, 1.f);
            cx.seq_ratio = expf(fminf(row_lw, 10.f));
            cx.seq_pass = (row_lw <= 10.f) ? 1.f : 0.f;
        }
        LossAcc acc;  // per row, folded in double below (the kernels fold per-thread partials)
        for (int t = 0; t < A; ++t) {
            const int i = b * A + t;
            d_lp[i] = loss_token<kType>(cx, acc, lp[i], olp[i], adv[i], mask ? mask[i] : 1.f, base ? base[i] : 0.f,
                                        tis ? tis[i] : 1.f);
        }
        s_loss += acc.s_loss;
        s_clip += acc.s_clip;
        s_kl += acc.s_kl;
        s_ratio += acc.s_ratio;
        r_max = fmaxf(r_max, acc.r_max);
        r_min = fminf(r_min, acc.r_min);
    }
    out[0] = s_loss;                         // policy loss (already weighted by the reduction)
    out[1] = s_clip * inv_total;             // clip ratio
    out[2] = s_kl;                           // mean_b masked_mean_t kl
    out[3] = s_ratio * inv_total;            // mean ratio
    out[4] = r_max;
    out[5] = r_min;
    out[6] = total_m;
}

extern "C" int skyb200_host_loss_tokens(const skyb200_loss_params* p, int ktype, int B, int A, const float* lp,
                                        const float* olp, const float* adv, const float* mask, const float* base,
                                        const float* tis, float* d_lp, double* out) {
    switch (ktype) {
        case -1: run<-1>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case -2: run<-2>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_REGULAR: run<SKYB200_LOSS_REGULAR>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_DUAL_CLIP: run<SKYB200_LOSS_DUAL_CLIP>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_GSPO: run<SKYB200_LOSS_GSPO>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_SAPO: run<SKYB200_LOSS_SAPO>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_CISPO: run<SKYB200_LOSS_CISPO>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_CROSS_ENTROPY: run<SKYB200_LOSS_CROSS_ENTROPY>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_IMPORTANCE_SAMPLING: run<SKYB200_LOSS_IMPORTANCE_SAMPLING>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
        case SKYB200_LOSS_KL_COV: run<SKYB200_LOSS_KL_COV>(*p, B, A, lp, olp, adv, mask, base, tis, d_lp, out); return 0;
    }
    return 1;
}
