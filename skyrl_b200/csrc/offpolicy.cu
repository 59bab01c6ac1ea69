// Off-policy correction of the policy loss: truncated importance-sampling (TIS) weights, sequence masks and the
// outlier-token mask.  Everything depends only on (old_log_probs, rollout_logprobs, loss_mask) - no gradient.
//
// Reference arithmetic (skyrl-train/skyrl_train/utils/off_policy_correction_utils.py):
//   compute_tis_ratio               :7-60
//   compute_outlier_token_mask      :63-112
//   compute_sequence_mask           :115-170
//   compute_off_policy_correction   :173-262   (order: metrics on the ORIGINAL mask -> outlier mask -> TIS ->
//                                               sequence mask, each later step seeing the mask of the earlier one)
//
// One block per sequence: pass 1 finds outlier tokens, pass 2 accumulates the sums under the updated mask,
// pass 3 writes tis[b, t] and mask_out[b, t]; a single-block kernel folds the per-row results into the metrics.
#include "common.cuh"

namespace skyb200 {

struct RowOut {      // per-row results, folded by opc_metrics_kernel
    float over_high;     // any token above the outlier threshold
    float under_low;     // any token below it
    float keep_outlier;  // 1 - (over_high | under_low)
    float tok_capped;    // # tokens with ratio > token cap under the updated mask
    float tok_total;     // # tokens with updated mask > 0
    float seq_capped;    // sequence TIS ratio above its cap
    float sm_over, sm_under, sm_keep;  // sequence mask decision
    float s_rm, s_m, s_x, s_xx, x_max, x_min;  // is_ratio statistics on the original mask, x = ratio * mask
};

__device__ __forceinline__ float safe_exp20(float d) { return expf(fminf(fmaxf(d, -20.f), 20.f)); }

__global__ void __launch_bounds__(256)
opc_row_kernel(const float* __restrict__ olp, const float* __restrict__ rlp, const float* __restrict__ mask,
               float* __restrict__ tis, float* __restrict__ mask_out, RowOut* __restrict__ rows, int64_t A,
               skyb200_opc_params p) {
    __shared__ float sh[34];
    const int row = blockIdx.x;
    const int64_t base = static_cast<int64_t>(row) * A;
    // ---- pass 1: outlier tokens + is_ratio statistics (original mask)
    float over = 0.f, under = 0.f, s_rm = 0.f, s_m = 0.f, s_x = 0.f, s_xx = 0.f, x_max = -INFINITY, x_min = INFINITY;
    for (int64_t t = threadIdx.x; t < A; t += blockDim.x) {
        const float m = mask[base + t];
        const float r = safe_exp20(olp[base + t] - rlp[base + t]);
        if (m > 0.f) {
            if (p.has_outlier_high && r > p.outlier_high) over = 1.f;
            if (p.has_outlier_low && r < p.outlier_low) under = 1.f;
        }
        const float x = r * m;
        s_rm += x;
        s_m += m;
        s_x += x;
        s_xx = fmaf(x, x, s_xx);
        x_max = fmaxf(x_max, x);
        x_min = fminf(x_min, x);
    }
    over = block_max(over, sh);
    under = block_max(under, sh);
    s_rm = block_sum(s_rm, sh);
    s_m = block_sum(s_m, sh);
    s_x = block_sum(s_x, sh);
    s_xx = block_sum(s_xx, sh);
    x_max = block_max(x_max, sh);
    x_min = block_min(x_min, sh);
    const float keep_outlier = (over > 0.f || under > 0.f) ? 0.f : 1.f;
    // ---- pass 2: sums under the outlier-updated mask
    float s_lr = 0.f, s_m1 = 0.f, capped = 0.f, total = 0.f;
    for (int64_t t = threadIdx.x; t < A; t += blockDim.x) {
        const float m1 = mask[base + t] * keep_outlier;
        const float d = olp[base + t] - rlp[base + t];
        s_lr = fmaf(d, m1, s_lr);  // un-clamped token log-ratios, as the reference sums them
        s_m1 += m1;
        if (m1 > 0.f) {
            total += 1.f;
            if (p.tis_type == 1 && safe_exp20(d) > p.token_tis_cap) capped += 1.f;
        }
    }
    s_lr = block_sum(s_lr, sh);
    s_m1 = block_sum(s_m1, sh);
    capped = block_sum(capped, sh);
    total = block_sum(total, sh);
    // ---- per-row decisions
    float seq_tis = 1.f, seq_capped = 0.f;
    if (p.tis_type == 2) {
        const float sr = safe_exp20(s_lr);
        seq_capped = (sr > p.seq_tis_cap) ? 1.f : 0.f;
        seq_tis = fminf(sr, p.seq_tis_cap);
    }
    float sm_over = 0.f, sm_under = 0.f, sm_keep = 1.f;
    if (p.seq_mask_type == 1) {  // geometric mean of the token ratios
        const float gm = safe_exp20(s_lr / fmaxf(s_m1, 1.f));
        sm_over = (gm > p.geo_high) ? 1.f : 0.f;
        sm_under = (gm < p.geo_low) ? 1.f : 0.f;
    } else if (p.seq_mask_type == 2) {  // product of the token ratios
        const float pr = safe_exp20(s_lr);
        sm_over = (pr > p.prod_high) ? 1.f : 0.f;
        sm_under = (pr < p.prod_low) ? 1.f : 0.f;
    }
    if (sm_over > 0.f || sm_under > 0.f) sm_keep = 0.f;
    // ---- pass 3: outputs
    const float keep = keep_outlier * sm_keep;
    for (int64_t t = threadIdx.x; t < A; t += blockDim.x) {
        mask_out[base + t] = mask[base + t] * keep;
        float w = 1.f;
        if (p.tis_type == 1)
            w = fminf(safe_exp20(olp[base + t] - rlp[base + t]), p.token_tis_cap);
        else if (p.tis_type == 2)
            w = seq_tis;
        tis[base + t] = w;
    }
    if (threadIdx.x == 0) {
        RowOut o;
        o.over_high = over;
        o.under_low = under;
        o.keep_outlier = keep_outlier;
        o.tok_capped = capped;
        o.tok_total = total;
        o.seq_capped = seq_capped;
        o.sm_over = sm_over;
        o.sm_under = sm_under;
        o.sm_keep = sm_keep;
        o.s_rm = s_rm;
        o.s_m = s_m;
        o.s_x = s_x;
        o.s_xx = s_xx;
        o.x_max = x_max;
        o.x_min = x_min;
        rows[row] = o;
    }
}

__global__ void __launch_bounds__(256) opc_metrics_kernel(const RowOut* __restrict__ rows, int B, int64_t A,
                                                          float* __restrict__ out, skyb200_opc_params p) {
    __shared__ float sh[34];
    __shared__ double shd[34];
    float over = 0.f, under = 0.f, masked = 0.f, cap = 0.f, tot = 0.f, seqcap = 0.f, smo = 0.f, smu = 0.f, smm = 0.f;
    float s_rm = 0.f, s_m = 0.f, mx = -INFINITY, mn = INFINITY;
    double s_x = 0.0, s_xx = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const RowOut r = rows[i];
        over += r.over_high;
        under += r.under_low;
        masked += 1.f - r.keep_outlier;
        cap += r.tok_capped;
        tot += r.tok_total;
        seqcap += r.seq_capped;
        smo += r.sm_over;
        smu += r.sm_under;
        smm += 1.f - r.sm_keep;
        s_rm += r.s_rm;
        s_m += r.s_m;
        s_x += r.s_x;
        s_xx += r.s_xx;
        mx = fmaxf(mx, r.x_max);
        mn = fminf(mn, r.x_min);
    }
    over = block_sum(over, sh);
    under = block_sum(under, sh);
    masked = block_sum(masked, sh);
    cap = block_sum(cap, sh);
    tot = block_sum(tot, sh);
    seqcap = block_sum(seqcap, sh);
    smo = block_sum(smo, sh);
    smu = block_sum(smu, sh);
    smm = block_sum(smm, sh);
    s_rm = block_sum(s_rm, sh);
    s_m = block_sum(s_m, sh);
    s_x = block_sum(s_x, shd);
    s_xx = block_sum(s_xx, shd);
    mx = block_max(mx, sh);
    mn = block_min(mn, sh);
    if (threadIdx.x == 0) {
        const float Bf = static_cast<float>(B);
        const double n = static_cast<double>(B) * static_cast<double>(A);
        const double mean = s_x / n;
        // torch.std (unbiased): a single element has zero degrees of freedom -> NaN, as the reference reports it
        const double var = n > 1.0 ? fmax(s_xx - n * mean * mean, 0.0) / (n - 1.0) : nan("");
        out[SKYB200_OPC_IS_RATIO_MEAN] = s_rm / fmaxf(s_m, 1.f);
        out[SKYB200_OPC_IS_RATIO_STD] = static_cast<float>(sqrt(var));
        out[SKYB200_OPC_IS_RATIO_MAX] = mx;
        out[SKYB200_OPC_IS_RATIO_MIN] = mn;
        out[SKYB200_OPC_OUTLIER_MASKED] = masked / Bf;
        out[SKYB200_OPC_OUTLIER_OVER_HIGH] = over / Bf;
        out[SKYB200_OPC_OUTLIER_UNDER_LOW] = under / Bf;
        out[SKYB200_OPC_TIS_CLIP_HIGH] = (p.tis_type == 1) ? cap / fmaxf(tot, 1.f) : seqcap / Bf;
        out[SKYB200_OPC_SEQ_MASKED] = smm / Bf;
        out[SKYB200_OPC_SEQ_OVER_HIGH] = smo / Bf;
        out[SKYB200_OPC_SEQ_UNDER_LOW] = smu / Bf;
        for (int i = SKYB200_OPC_SEQ_UNDER_LOW + 1; i < SKYB200_OPC_NMETRICS; ++i) out[i] = 0.f;
    }
}

}  // namespace skyb200

using namespace skyb200;

extern "C" size_t skyb200_off_policy_workspace(int64_t B) { return (size_t)(B * sizeof(RowOut) + 512); }

extern "C" int skyb200_off_policy_correction(const float* old_log_probs, const float* rollout_logprobs,
                                             const float* loss_mask, float* tis_out, float* mask_out,
                                             float* metrics_out, void* workspace, size_t workspace_bytes, int64_t B,
                                             int64_t A, const skyb200_opc_params* params, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(old_log_probs && rollout_logprobs && loss_mask && tis_out && mask_out && metrics_out &&
                        workspace && params,
                    "null pointer");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < (1ll << 31), "bad shape");
    SKYB200_REQUIRE(params->tis_type >= 0 && params->tis_type <= 2, "Unknown tis_ratio_type");
    SKYB200_REQUIRE(params->seq_mask_type >= 0 && params->seq_mask_type <= 2, "Unknown sequence_mask_metric");
    if (workspace_bytes < skyb200_off_policy_workspace(B)) return fail(SKYB200_EWORKSPACE, "workspace too small");
    RowOut* rows = reinterpret_cast<RowOut*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    opc_row_kernel<<<(unsigned)B, 256, 0, stream>>>(old_log_probs, rollout_logprobs, loss_mask, tis_out, mask_out,
                                                    rows, A, *params);
    int rc = check_launch("opc_row_kernel");
    if (rc) return rc;
    opc_metrics_kernel<<<1, 256, 0, stream>>>(rows, (int)B, A, metrics_out, *params);
    return check_launch("opc_metrics_kernel");
}
