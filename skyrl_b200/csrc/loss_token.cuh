// Per-token arithmetic of the policy losses (value, d loss / d log-prob, the reductions the worker reports), shared by
// loss_main_kernel and loss_flat_kernel (loss.cu).  Plain scalar fp32 math with no CUDA intrinsic, so the same source
// also compiles for the host: tests/test_loss_token_host_cpu.py builds it with g++ and checks every loss type, KL
// estimator and reduction against the oracle in the CPU suite (test infrastructure - nothing in the product calls the
// host build).
//
// Reference arithmetic: utils/ppo_utils.py:548-586 (regular / dual_clip), :661-725 (gspo), :589-658 (sapo),
// :728-761 (cispo), :892-981 (cross_entropy, importance_sampling), :833-888 (kl_cov), :88-124 (KL estimators).
#pragma once
#include <cmath>

#include "../../include/skyb200.h"

#ifdef __CUDACC__
#define SKYB200_TOKEN_FN __device__ __forceinline__
#else
#define SKYB200_TOKEN_FN inline
#endif

// exp() of a clamped argument (|x| <= 20, or <= 10 for the GSPO sequence ratio).  On the device this is
// ex2.approx(x * log2 e): 2 instructions instead of ~10, relative error <= 2^-22 (ex2.approx) + 2^-24 * |x| * log2 e
// (rounding of the scaled argument) <= 1.3e-6 at |x| = 20 - inside the 1e-5 tolerance of the loss, its gradient and
// the ratio / KL metrics, and exact where it matters structurally: exp(0) == 1 (on-policy ratio), and the clip tie
// s1 == s2 only depends on the ratio lying inside the clip range, not on its last bits.  The token pass is
// instruction-bound (profiles/README.md r01g), so the two exponentials per token are worth removing.  The host build
// of this header (tests/test_loss_token_host_cpu.py) keeps libm's expf.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float skyb200_ex2_exp(float x) {  // no denormal pre-scaling sequence around the MUFU (.ftz)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
#define SKYB200_EXPF(x) skyb200_ex2_exp(x)
#else
#define SKYB200_EXPF(x) expf(x)
#endif

namespace skyb200 {

SKYB200_TOKEN_FN float kl_estimate(float lp, float base, int type, float* dkl_dlp) {
    const float d = lp - base;
    switch (type) {
        case SKYB200_KL_K1:
            *dkl_dlp = 1.f;
            return d;
        case SKYB200_KL_ABS:
            *dkl_dlp = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
            return fabsf(d);
        case SKYB200_KL_K2:
            *dkl_dlp = d;
            return 0.5f * d * d;
        default: {  // k3
            const float xr = base - lp;
            const float x = fminf(fmaxf(xr, -20.f), 20.f);
            const float r = SKYB200_EXPF(x);
            const float k = r - x - 1.f;
            const float kc = fminf(fmaxf(k, -10.f), 10.f);
            const bool pass = (xr >= -20.f && xr <= 20.f) && (k >= -10.f && k <= 10.f);
            *dkl_dlp = pass ? (1.f - r) : 0.f;
            return kc;
        }
    }
}

struct LossAcc {
    float s_loss = 0.f, s_clip = 0.f, s_kl = 0.f, s_ent = 0.f, s_ratio = 0.f;
    float r_max = -INFINITY, r_min = INFINITY;
};
struct TokenCtx {
    bool gspo, dual, has_mask, has_base;
    float lo, hi, seq_ratio, seq_pass, w_scale, kl_scale, clip_c, kl_grad_scale, d_ent_val;
    float tau_pos = 1.f, tau_neg = 1.f;
    int kl_type;
    int loss_type = SKYB200_LOSS_REGULAR;
};

// One response token: accumulates the reductions, returns d final_loss / d log_prob.
// kType / kBase >= 0 fix the loss type / the presence of base log-probs at compile time (the [B, A] token pass is
// instantiated per loss type: it is instruction-bound, not only memory-bound); -1 = decide at run time;
// kType == -2 = decide at run time among the clipped-surrogate family only (regular / dual_clip / gspo).
// kKl == SKYB200_KL_K3 additionally fixes "k3 estimator, KL term without gradient" (the reference's KL loss).
template <int kType = -1, int kBase = -1, int kKl = -1>
SKYB200_TOKEN_FN float loss_token(const TokenCtx& cx, LossAcc& acc, float lp, float olp, float adv, float m,
                                            float base, float tis = 1.f, float cm = -1.f) {
    const int loss_type = (kType >= 0) ? kType : (kType == -2 ? static_cast<int>(SKYB200_LOSS_REGULAR) : cx.loss_type);
    const bool is_gspo = (kType >= 0) ? (kType == SKYB200_LOSS_GSPO) : cx.gspo;
    const bool is_dual = (kType >= 0) ? (kType == SKYB200_LOSS_DUAL_CLIP) : cx.dual;
    const bool has_base = (kBase >= 0) ? (kBase != 0) : cx.has_base;
    float ratio, loss, dl_dlp, clipped = 0.f;
    if (loss_type <= SKYB200_LOSS_GSPO) {
        float dr_dlp;
        if (is_gspo) {
            ratio = cx.seq_ratio;
            dr_dlp = cx.seq_ratio * cx.seq_pass;
        } else {
            const float delta = lp - olp;
            ratio = SKYB200_EXPF(fminf(fmaxf(delta, -20.f), 20.f));
            dr_dlp = (delta >= -20.f && delta <= 20.f) ? ratio : 0.f;
        }
        const float rc = fminf(fmaxf(ratio, cx.lo), cx.hi);
        const float in_rng = (ratio >= cx.lo && ratio <= cx.hi) ? 1.f : 0.f;
        const float s1 = ratio * adv, s2 = rc * adv;
        loss = -fminf(s1, s2);
        // d loss / d ratio: torch.min splits ties 50/50 between its two arguments
        float dl_dr = (s1 < s2) ? -adv : ((s1 > s2) ? -adv * in_rng : -adv * (0.5f + 0.5f * in_rng));
        clipped = (s2 < s1) ? 1.f : 0.f;  // (-surr2 > -surr1), before dual clip
        if (is_dual && adv < 0.f) {
            const float l3 = -adv * cx.clip_c;
            if (l3 < loss) {
                loss = l3;
                dl_dr = 0.f;
            } else if (l3 == loss) {
                dl_dr *= 0.5f;
            }
        }
        dl_dlp = dl_dr * dr_dlp;
    } else if (loss_type == SKYB200_LOSS_SAPO) {
        const float delta = lp - olp;
        ratio = SKYB200_EXPF(fminf(fmaxf(delta, -20.f), 20.f));
        const float dr_dlp = (delta >= -20.f && delta <= 20.f) ? ratio : 0.f;
        const float tau = (adv > 0.f) ? cx.tau_pos : cx.tau_neg;
        const float sg = 1.f / (1.f + SKYB200_EXPF(-tau * (ratio - 1.f)));
        loss = -(sg * (4.f / tau)) * adv;
        dl_dlp = -adv * 4.f * sg * (1.f - sg) * dr_dlp;
    } else if (loss_type == SKYB200_LOSS_CISPO) {
        ratio = SKYB200_EXPF(fminf(fmaxf(lp - olp, -20.f), 20.f));
        const float rc = fminf(fmaxf(ratio, cx.lo), cx.hi);  // detached: acts as a per-token weight
        loss = -adv * rc * lp;
        dl_dlp = -adv * rc;
        clipped = (ratio < cx.lo || ratio > cx.hi) ? 1.f : 0.f;
    } else if (loss_type == SKYB200_LOSS_CROSS_ENTROPY) {
        ratio = 1.f;
        loss = -lp;
        dl_dlp = -1.f;
    } else if (loss_type == SKYB200_LOSS_KL_COV) {  // `tis` carries the per-token KL weight of the selection
        const float d = lp - olp;
        ratio = SKYB200_EXPF(d);
        loss = fmaf(tis, fabsf(d), -ratio * adv);
        dl_dlp = fmaf(tis, (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f), -ratio * adv);
        tis = 1.f;
    } else {  // importance sampling: un-clamped ratio, Tinker semantics
        ratio = SKYB200_EXPF(lp - olp);
        loss = -ratio * adv;
        dl_dlp = loss;
    }
    loss *= tis;    // off-policy correction weight (detached)
    dl_dlp *= tis;
    const float w = m * cx.w_scale;
    acc.s_loss = fmaf(loss, w, acc.s_loss);
    acc.s_clip = fmaf(clipped, cm >= 0.f ? cm : m, acc.s_clip);
    acc.s_ratio = fmaf(ratio, m, acc.s_ratio);
    if (m != 0.f) {
        acc.r_max = fmaxf(acc.r_max, ratio);
        acc.r_min = fminf(acc.r_min, ratio);
    }
    float g = w * dl_dlp;
    if (has_base) {
        if constexpr (kKl == SKYB200_KL_K3) {  // value only: same arithmetic as kl_estimate's k3 branch
            const float x = fminf(fmaxf(base - lp, -20.f), 20.f);
            const float kl = fminf(fmaxf(SKYB200_EXPF(x) - x - 1.f, -10.f), 10.f);
            const float km = cx.has_mask ? kl * m : kl;
            acc.s_kl = fmaf(km * m, cx.kl_scale, acc.s_kl);
        } else {
            float dk;
            const float kl = kl_estimate(lp, base, cx.kl_type, &dk);
            const float km = cx.has_mask ? kl * m : kl;  // compute_approx_kl multiplies by the mask ...
            acc.s_kl = fmaf(km * m, cx.kl_scale, acc.s_kl);  // ... and masked_mean multiplies again
            g = fmaf(cx.kl_grad_scale * m * (cx.has_mask ? m : 1.f), dk, g);
        }
    }
    return g;
}

}  // namespace skyb200
