/*
 * skyb200 — C ABI of the B200-native policy-update hot path (libskyb200.so).
 *
 * The reference (SkyRL / skyrl_train) has no FFI for this path: its boundary is the Python plugin API
 * (AdvantageEstimatorRegistry / PolicyLossRegistry, logprobs_from_logits, chunked_entropy_from_logits).
 * The Python shims in skyrl_b200/ present exactly that API and call the entry points below; each entry
 * point names the reference function (file:line under skyrl-train/skyrl_train/) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; the caller owns and sizes every buffer;
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises the device;
 *   - return value 0 = success, otherwise a SKYB200_E* code; skyb200_last_error() gives the text
 *     (thread-local);
 *   - matrices are row-major; [B, A] = sequences x response positions; N = tokens, H = hidden, V = vocab.
 */
#ifndef SKYB200_H
#define SKYB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKYB200_OK 0
#define SKYB200_EINVAL 1 /* bad argument (null pointer, bad enum, misaligned leading dimension) */
#define SKYB200_ECUDA 2  /* a CUDA runtime / driver call failed */
#define SKYB200_EARCH 3  /* device is not compute capability 10.x */
#define SKYB200_EWORKSPACE 4

/* mask element types accepted by the advantage kernels (reference masks are int64 or fp32) */
#define SKYB200_MASK_F32 0
#define SKYB200_MASK_I64 1
#define SKYB200_MASK_NONE 2

/* policy_loss_type (ppo_utils.py:463-508) */
#define SKYB200_LOSS_REGULAR 0
#define SKYB200_LOSS_DUAL_CLIP 1
#define SKYB200_LOSS_GSPO 2
#define SKYB200_LOSS_SAPO 3                /* sapo_policy_loss, ppo_utils.py:589-658 */
#define SKYB200_LOSS_CISPO 4               /* compute_policy_loss_cispo, ppo_utils.py:728-761 (eps_clip_* = cispo eps) */
#define SKYB200_LOSS_CROSS_ENTROPY 5       /* cross_entropy_loss, ppo_utils.py:892-930 (use SKYB200_REDUCE_SUM) */
#define SKYB200_LOSS_IMPORTANCE_SAMPLING 6 /* importance_sampling_loss, ppo_utils.py:933-981 (SKYB200_REDUCE_SUM) */
#define SKYB200_LOSS_KL_COV 7 /* compute_policy_loss_kl_cov, ppo_utils.py:833-888: -A*exp(lp-olp) + w*|lp-olp| where the
                                 per-token weight w (ppo_kl_coef on the selected tokens, 0 elsewhere; from
                                 skyb200_cov_select) is passed as `tis`.  skyb200_policy_loss only. */
#define SKYB200_LOSS_LAST 7

/* loss_reduction (ppo_utils.py:984-1009) */
#define SKYB200_REDUCE_TOKEN_MEAN 0
#define SKYB200_REDUCE_SEQUENCE_MEAN 1
#define SKYB200_REDUCE_SEQ_MEAN_TOKEN_SUM_NORM 2
#define SKYB200_REDUCE_SUM 3 /* plain masked sum (Tinker semantics of cross_entropy / importance_sampling) */

/* kl_estimator_type (ppo_utils.py:88-124) */
#define SKYB200_KL_K1 0
#define SKYB200_KL_ABS 1
#define SKYB200_KL_K2 2
#define SKYB200_KL_K3 3

int skyb200_abi_version(void);
const char* skyb200_last_error(void);
/* 0 if the current device is sm_10x and the library can run, else SKYB200_EARCH / SKYB200_ECUDA */
int skyb200_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Policy loss, fused forward + backward (+ KL-loss and entropy terms of the worker's loss assembly).
 * Replaces ppo_policy_loss (ppo_utils.py:548-586), gspo_policy_loss (:661-725), reduce_loss (:984-1009),
 * compute_approx_kl (:88-124) as used at worker.py:864-874, the entropy masked_mean (worker.py:852-861)
 * and loss = policy + kl*coef - entropy*coef (worker.py:876).
 * ------------------------------------------------------------------------------------------------ */
typedef struct skyb200_loss_params {
    int32_t loss_type;   /* SKYB200_LOSS_* */
    int32_t reduction;   /* SKYB200_REDUCE_* */
    int32_t kl_type;     /* SKYB200_KL_*; used iff base_log_probs != NULL */
    int32_t kl_has_grad; /* 0 = reference behaviour (compute_approx_kl is @torch.no_grad) */
    int32_t use_entropy_loss; /* 1: final loss -= entropy*coef and d_entropy carries its gradient */
    int32_t reserved0;
    float eps_clip_low;
    float eps_clip_high;
    float clip_ratio_c;
    float max_seq_len; /* for SEQ_MEAN_TOKEN_SUM_NORM */
    float kl_loss_coef;
    float entropy_loss_coef;
    float sapo_tau_pos; /* SAPO gate temperatures (config.sapo.tau_pos / tau_neg) */
    float sapo_tau_neg;
} skyb200_loss_params;

/* indices into the fp32 stats vector written by skyb200_policy_loss (SKYB200_LOSS_NSTATS floats) */
#define SKYB200_STAT_POLICY_LOSS 0
#define SKYB200_STAT_CLIP_RATIO 1
#define SKYB200_STAT_KL_LOSS 2      /* mean_b(masked_mean_t(kl)) ; 0 if no base_log_probs */
#define SKYB200_STAT_ENTROPY 3      /* masked_mean(entropy, loss_mask); 0 if no entropy */
#define SKYB200_STAT_FINAL_LOSS 4   /* policy + kl*kl_coef - entropy*ent_coef (ent term iff use_entropy_loss) */
#define SKYB200_STAT_MASK_SUM 5     /* sum(loss_mask) (B*A if mask is NULL) */
#define SKYB200_STAT_RATIO_MEAN 6   /* masked mean of the importance ratio (diagnostic) */
#define SKYB200_STAT_RATIO_MAX 7
#define SKYB200_STAT_RATIO_MIN 8
#define SKYB200_LOSS_NSTATS 16

/* bytes of scratch the call needs */
size_t skyb200_policy_loss_workspace(int64_t B, int64_t A);

/*
 * log_probs, old_log_probs, advantages: [B, A] fp32.  loss_mask: [B, A] fp32 or NULL (= all ones).
 * base_log_probs: [B, A] fp32 or NULL (no KL term).  entropy: [B, A] fp32 or NULL.
 * d_log_probs: [B, A] fp32 out, = d final_loss / d log_probs (NULL = forward only).
 * d_entropy:   [B, A] fp32 out or NULL; zeros unless params->use_entropy_loss.
 * stats: SKYB200_LOSS_NSTATS floats out.
 * tis [B, A] fp32 or NULL: detached per-token weight multiplied into the loss (off-policy correction);
 * clip_mask [B, A] fp32 or NULL: mask for the clip_ratio metric when it differs from loss_mask (the reference
 * computes clip_ratio with the mask it had BEFORE the off-policy correction shrank it, ppo_utils.py:570-583).
 */
int skyb200_policy_loss(const float* log_probs, const float* old_log_probs, const float* advantages,
                        const float* loss_mask, const float* base_log_probs, const float* entropy,
                        float* d_log_probs, float* d_entropy, float* stats, void* workspace, size_t workspace_bytes,
                        int64_t B, int64_t A, const skyb200_loss_params* params, const float* tis,
                        const float* clip_mask, void* stream);

/*
 * Critic (value) loss, forward + backward (ppo_critic_loss, ppo_utils.py:175-193):
 *   loss = 0.5 * mean_b masked_mean_t max((old + clamp(v - old, +-clip) - R)^2, (v - R)^2)   (use_clip != 0)
 *        = 0.5 * mean_b masked_mean_t (v - R)^2                                               (use_clip == 0)
 * stats[0] = loss, stats[1] = clipfrac = masked_mean(1[surr1 > surr2]), stats[2] = sum(mask); 4 floats.
 * d_values [B, A] out (NULL = forward only).  workspace: skyb200_policy_loss_workspace(B, A) bytes.
 */
int skyb200_critic_loss(const float* values, const float* old_values, const float* returns, const float* loss_mask,
                        float* d_values, float* stats, void* workspace, size_t workspace_bytes, int64_t B, int64_t A,
                        float value_clip, int32_t use_clip, void* stream);

/*
 * Off-policy correction (compute_off_policy_correction, off_policy_correction_utils.py:173-262): TIS weights,
 * outlier-token mask and sequence mask from (old_log_probs, rollout_logprobs, loss_mask).  No gradient.
 *   tis_out  [B, A]: weight the per-token loss is multiplied with (1 when tis_type == 0)
 *   mask_out [B, A]: loss_mask * outlier_mask * sequence_mask
 *   metrics_out: SKYB200_OPC_NMETRICS floats (indices below)
 */
typedef struct skyb200_opc_params {
    int32_t tis_type;      /* 0 none, 1 "token", 2 "sequence" */
    int32_t seq_mask_type; /* 0 none, 1 "geometric", 2 "product" */
    int32_t has_outlier_low;
    int32_t has_outlier_high;
    float token_tis_cap; /* token_tis_ratio_clip_high */
    float seq_tis_cap;   /* sequence_tis_ratio_clip_high */
    float outlier_low;   /* outlier_token_is_threshold_low / _high */
    float outlier_high;
    float geo_low; /* geo_mask_low / geo_mask_high */
    float geo_high;
    float prod_low; /* product_mask_low / product_mask_high */
    float prod_high;
} skyb200_opc_params;
#define SKYB200_OPC_IS_RATIO_MEAN 0
#define SKYB200_OPC_IS_RATIO_STD 1
#define SKYB200_OPC_IS_RATIO_MAX 2
#define SKYB200_OPC_IS_RATIO_MIN 3
#define SKYB200_OPC_OUTLIER_MASKED 4
#define SKYB200_OPC_OUTLIER_OVER_HIGH 5
#define SKYB200_OPC_OUTLIER_UNDER_LOW 6
#define SKYB200_OPC_TIS_CLIP_HIGH 7 /* tis_token_clip_high_ratio or tis_seq_clip_high_ratio */
#define SKYB200_OPC_SEQ_MASKED 8    /* geo_/product_sequence_mask_masked_ratio */
#define SKYB200_OPC_SEQ_OVER_HIGH 9
#define SKYB200_OPC_SEQ_UNDER_LOW 10
#define SKYB200_OPC_NMETRICS 16
size_t skyb200_off_policy_workspace(int64_t B);
int skyb200_off_policy_correction(const float* old_log_probs, const float* rollout_logprobs, const float* loss_mask,
                                  float* tis_out, float* mask_out, float* metrics_out, void* workspace,
                                  size_t workspace_bytes, int64_t B, int64_t A, const skyb200_opc_params* params,
                                  void* stream);

/*
 * Critic value head (model_wrapper.py:472-489: value_head = nn.Linear(H, 1)): values[r] = <hidden[r, :], w> + bias.
 * hidden [N, H] bf16, weight [H] fp32, values / g_values [N] fp32, d_hidden [N, H] bf16 (or NULL),
 * d_weight [H] fp32 (or NULL; += if accumulate).  The gradient of the bias is sum(g_values) (caller's one-liner).
 */
int skyb200_value_head_forward(const void* hidden, const float* weight, float bias, float* values, int64_t N,
                               int64_t H, void* stream);
size_t skyb200_value_head_backward_workspace(int64_t N, int64_t H);
int skyb200_value_head_backward(const void* hidden, const float* weight, const float* g_values, void* d_hidden,
                                float* d_weight, int32_t accumulate, void* workspace, size_t workspace_bytes,
                                int64_t N, int64_t H, void* stream);

/* x[i] *= *scale_dev  (applies an upstream autograd gradient that lives on the device) */
int skyb200_scale_inplace(float* x, const float* scale_dev, int64_t n, void* stream);

/* masked_mean(x, mask) over everything (torch_utils.py:180-184); out: 1 float.  mask may be NULL. */
int skyb200_masked_mean(const float* x, const float* mask, float* out, void* workspace, size_t workspace_bytes,
                        int64_t n, void* stream);

/* compute_approx_kl (ppo_utils.py:88-124): out = estimator(lp, base) [* mask] */
int skyb200_approx_kl(const float* log_probs, const float* base_log_probs, const float* loss_mask, float* out,
                      int64_t n, int32_t kl_type, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Advantage estimation.
 * ------------------------------------------------------------------------------------------------ */

/*
 * GRPO (compute_grpo_outcome_advantage, ppo_utils.py:1132-1182) and RLOO
 * (compute_rloo_outcome_advantage, :1048-1098).  The uid -> dense group id map is built on the host
 * (bit-exact group assignment); groups are passed in CSR form:
 *   group_of_row[B]  int32  dense group id of each row
 *   group_start[G+1] int32  offsets into group_rows
 *   group_rows[B]    int32  row indices sorted by group (original order inside a group)
 * mode: 0 = GRPO with std normalisation, 1 = GRPO mean only (grpo_norm_by_std=False), 2 = RLOO.
 * rewards [B, A] fp32; mask [B, A] of mask_type; out [B, A] fp32 (advantages == returns, one buffer).
 * scores_ws: B floats of scratch.
 */
int skyb200_group_advantage(const float* rewards, const void* response_mask, int32_t mask_type,
                            const int32_t* group_of_row, const int32_t* group_start, const int32_t* group_rows,
                            float* out, float* scores_ws, int64_t B, int64_t A, int32_t num_groups, int32_t mode,
                            float epsilon, void* stream);

/*
 * The same estimators in ONE launch for batches whose prompt groups all have `group_size` rows (the trainer's
 * n_samples_per_prompt layout): a thread-block cluster per group exchanges the row sums through distributed shared
 * memory; rewards and mask cross HBM exactly once.  group g owns rows group_rows[g * group_size .. (g + 1) *
 * group_size).  Shapes outside skyb200_group_advantage_uniform_supported(A, group_size) (A % 4, A <= 32768,
 * group_size <= 8, or even and <= 16 with A <= 16384) take skyb200_group_advantage.  Pointers 16-byte aligned.
 */
int skyb200_group_advantage_uniform_supported(int64_t A, int64_t group_size);
int skyb200_group_advantage_uniform(const float* rewards, const void* response_mask, int32_t mask_type,
                                    const int32_t* group_rows, float* out, float* scores_ws, int64_t B, int64_t A,
                                    int32_t group_size, int32_t mode, float epsilon, void* stream);

/*
 * GAE (compute_gae_advantage_return, ppo_utils.py:1101-1129) + masked_whiten (:163-172).
 * One segment per ROW, no reset on mask (the reference does not reset).
 * rewards, values [B, A] fp32; mask of mask_type; advantages, returns [B, A] fp32 out.
 * stats_out: 4 floats {mask_sum, mean, var (Bessel), rstd}; the caller must raise if mask_sum is 0 or 1
 * (masked_var, ppo_utils.py:148-160).  workspace: skyb200_gae_workspace(B) bytes.
 */
size_t skyb200_gae_workspace(int64_t B);
int skyb200_gae_advantage(const float* rewards, const float* values, const void* response_mask, int32_t mask_type,
                          float* advantages, float* returns, float* stats_out, void* workspace,
                          size_t workspace_bytes, int64_t B, int64_t A, float gamma, float lambd, void* stream);

/*
 * The two halves of skyb200_gae_advantage, for batches sharded over several GPUs (whitening needs GLOBAL
 * moments, trainer.py:268-276): scan locally, all-reduce `moments` (3 doubles: sum m, sum m*x, sum m*x^2),
 * then whiten.  moments is a device pointer, 8-byte aligned.
 */
int skyb200_gae_scan(const float* rewards, const float* values, const void* response_mask, int32_t mask_type,
                     float* adv_raw, float* returns, double* moments, void* workspace, size_t workspace_bytes,
                     int64_t B, int64_t A, float gamma, float lambd, void* stream);
int skyb200_whiten_from_moments(float* x, const double* moments, float* stats_out, int64_t n, void* stream);

/*
 * REINFORCE++ (compute_reinforce_plus_plus_outcome_advantage, ppo_utils.py:1013-1045):
 *   returns_t = r_t + gamma * returns_{t+1} * mask_{t+1} (reset after EOS); advantages = masked_whiten(returns) * mask.
 * stats_out: 4 floats {mask_sum, mean, var, rstd} (raise for mask_sum in {0, 1}); workspace: skyb200_gae_workspace(B).
 */
int skyb200_reinforce_pp_advantage(const float* rewards, const void* response_mask, int32_t mask_type,
                                   float* advantages, float* returns, float* stats_out, void* workspace,
                                   size_t workspace_bytes, int64_t B, int64_t A, float gamma, void* stream);

/*
 * normalize_advantages_dict (ppo_utils.py:127-145): mean over ALL n positions (unmasked), variance over the mask:
 *   out = (adv - mean) * rsqrt(max(sum(m * (adv - mean)^2) / sum(m), 1e-8)).   stats_out: {sum m, mean, var, rstd}.
 */
size_t skyb200_normalize_advantages_workspace(void);
int skyb200_normalize_advantages(const float* advantages, const void* response_mask, int32_t mask_type, float* out,
                                 float* stats_out, void* workspace, size_t workspace_bytes, int64_t n, void* stream);

/* out[b, :] = src[row_index[b], :] (fp32 [*, A]); the step-wise trajectory broadcast of trainer.py:795-804 */
int skyb200_gather_rows(const float* src, const int32_t* row_index, float* out, int64_t num_src_rows, int64_t B,
                        int64_t A, void* stream);

/*
 * KL-in-reward (RayPPOTrainer.apply_reward_kl_penalty, trainer.py:981-1035):
 *   kl = compute_approx_kl(lp, base, loss_mask); rewards_out = rewards - kl * max(0, kl_coef)
 *   metrics_out[0] = mean_b(masked_mean_t(kl)), metrics_out[1] = mean_b(max_t |kl|)
 * row_ws: 2*B floats scratch.
 */
int skyb200_reward_kl_penalty(const float* log_probs, const float* base_log_probs, const float* loss_mask,
                              const float* rewards, float* rewards_out, float* metrics_out, float* row_ws, int64_t B,
                              int64_t A, int32_t kl_type, float kl_coef, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused LM head: projection + online log-softmax + label gather + entropy, never materialising
 * [tokens x vocab].  Replaces lm_head GEMM + logits.div_(T) + logprobs_from_logits + chunked entropy
 * (model_wrapper.py:309-363, torch_utils.py:59-177).
 * ------------------------------------------------------------------------------------------------ */

/* scratch bytes for forward / backward with token chunks of `chunk` rows (0 = library default) */
size_t skyb200_head_forward_workspace(int64_t N, int64_t H, int64_t V, int64_t chunk);
size_t skyb200_head_backward_workspace(int64_t N, int64_t H, int64_t V, int64_t chunk);

/*
 * hidden [N, H] bf16, weight [V, H] bf16 (nn.Linear layout), labels [N] int64.
 * logprobs [N] fp32 out; entropy [N] fp32 out (NULL to skip); lse2 [N] fp32 out = log2-sum-exp of
 * logits*log2(e)/T, saved for backward.  entropy_mask [N] fp32 or NULL multiplies the entropy
 * (attention_mask semantics of chunked_entropy_from_logits).
 */
int skyb200_head_forward(const void* hidden, const void* weight, const int64_t* labels, const float* entropy_mask,
                         float* logprobs, float* entropy, float* lse2, void* workspace, size_t workspace_bytes,
                         int64_t N, int64_t H, int64_t V, float temperature, int64_t chunk, void* stream);

/*
 * Token selection of the covariance losses (detached, no gradient): clip_cov (ppo_utils.py:785-814) and kl_cov
 * (:857-871).  cov = (adv - mean_adv) * (lp - mean_lp) per valid token.
 *   SKYB200_COV_KL  : the k = max(1, int(n_valid * frac)) valid tokens with the largest cov are selected
 *                     (ties: lowest flat index first).  Means are plain means over loss_mask > 0.
 *   SKYB200_COV_CLIP: candidates = valid tokens the PPO clip does not already clip, with cov_lb < cov < cov_ub;
 *                     k = min(max(int(frac * sum(mask)), 1), #candidates) of them are drawn uniformly at random
 *                     (counter-based generator keyed by `seed`: same seed, same draw).  Means are masked means.
 * out [B, A] fp32 = value_selected on the selected tokens, value_other elsewhere.  counts (device, 2 floats, may be
 * NULL) = {number selected, number of candidates}.  No host synchronisation: k is computed on the device.
 */
#define SKYB200_COV_KL 0
#define SKYB200_COV_CLIP 1
typedef struct {
    int32_t mode; /* SKYB200_COV_* */
    int32_t reserved0;
    float eps_clip_low, eps_clip_high; /* PPO clip range (clip_cov only) */
    float cov_lb, cov_ub;              /* clip_cov only */
    double frac;                       /* kl_cov_frac or clip_cov.clip_ratio */
    float value_selected, value_other;
    uint64_t seed;
} skyb200_cov_params;
size_t skyb200_cov_select_workspace(int64_t B, int64_t A);
int skyb200_cov_select(const float* log_probs, const float* old_log_probs, const float* advantages,
                       const float* loss_mask, float* out, float* counts, void* workspace, size_t workspace_bytes,
                       int64_t B, int64_t A, const skyb200_cov_params* params, void* stream);

/*
 * Vocab-parallel (tensor-parallel) building block — the B200 counterpart of Megatron's DistributedLogprob /
 * _VocabParallelEntropy (distributed/megatron/model_utils.py:64-247, 548-590).  `weight` is ONE vocab shard
 * [V_shard, H]; `labels` are shard-local (global label - shard start; labels outside [0, V_shard) never match).
 * stats [N, 4] fp32 out = per token {max_j y_j, sum_j 2^(y_j - max), sum_j 2^(y_j - max) * logit_j, logit[label] or 0}
 * with y = logit * log2(e) / T.  The ranks all-gather these 16 bytes per token and combine them
 * (skyrl_b200/tp_head.py); the backward is skyb200_head_backward on the shard with the GLOBAL lse2 / entropy.
 */
int skyb200_head_forward_stats(const void* hidden, const void* weight, const int64_t* labels, float* stats,
                               void* workspace, size_t workspace_bytes, int64_t N, int64_t H, int64_t V,
                               float temperature, int64_t chunk, void* stream);
/* stats_all [tp, N, 4] (the all-gathered per-shard stats) -> logprobs / entropy / lse2 [N] (lse2 = log2-sum-exp2 of
 * the scaled logits over the WHOLE vocabulary, what skyb200_head_backward expects). */
int skyb200_head_combine_stats(const float* stats_all, int32_t tp, int64_t N, float temperature, float* logprobs,
                               float* entropy, float* lse2, void* stream);

/*
 * Backward: given g_logprobs [N] and optional g_entropy [N] (upstream grads w.r.t. the forward outputs,
 * g_entropy already multiplied by entropy_mask by the caller if one was used) produce
 *   d_hidden [N, H] bf16   and   d_weight [V, H] fp32 (+= if accumulate != 0, else overwritten).
 * `entropy` is the UNMASKED per-token entropy from forward (needed iff g_entropy != NULL).
 */
int skyb200_head_backward(const void* hidden, const void* weight, const int64_t* labels, const float* lse2,
                          const float* entropy, const float* g_logprobs, const float* g_entropy, void* d_hidden,
                          float* d_weight, int32_t accumulate, void* workspace, size_t workspace_bytes, int64_t N,
                          int64_t H, int64_t V, float temperature, int64_t chunk, void* stream);

/*
 * The whole step in one call: head forward -> loss assembly -> head backward, chunk by chunk over the tokens
 * (replaces worker.py:787-877 after the trunk).  Because the loss gradient of a chunk is available as soon as its
 * log-probs are (normalisers come from a mask-only pre-pass), the forward GEMM keeps a bf16 copy of its logits
 * tiles in a bounded [chunk, V] store and the backward reads it instead of recomputing the projection:
 * 6*H*V executed flop per token instead of 8*H*V.  GSPO chunks are whole sequences (k * A tokens; A need not be
 * a multiple of the 256-row tile).
 *
 * old_log_probs / advantages / loss_mask / base_log_probs: [B, A] fp32 (mask / base may be NULL);
 * entropy_mask [B*A] fp32 or NULL; log_probs / entropy: [B*A] fp32 out (entropy already * entropy_mask);
 * d_hidden [B*A, H] bf16 and d_weight [V, H] fp32 may be NULL (both NULL = evaluate the loss only);
 * gradients are those of stats[SKYB200_STAT_FINAL_LOSS] with upstream gradient 1.
 *
 * Packed (ragged) layout: pass cu_seqlens (device int32 [B + 1], cu[0] = 0, cu[B] = n_tokens).  Every per-token
 * array is then flat [n_tokens]; sequence b owns tokens [cu[b], cu[b + 1]); A is ignored; positions that a padded
 * batch would mask out simply do not exist.  Results equal the padded layout with loss_mask = 0 on the padding.
 * (GSPO is not available in the packed layout.)  Padded layout: cu_seqlens = NULL, n_tokens ignored.
 */
size_t skyb200_policy_step_workspace(int64_t B, int64_t A, int64_t n_tokens, int64_t H, int64_t V, int32_t loss_type,
                                     int32_t need_grad);
/*
 * How skyb200_policy_step will cut that batch into token chunks (host-only, no CUDA call): plan_out[0] = row capacity
 * of the per-chunk buffers (a multiple of 256), [1] = tokens one chunk covers (== capacity, except GSPO, whose chunks
 * are whole sequences: k * A tokens for any padded length A, tile-aligned or not), [2] = number of chunks.
 */
int skyb200_policy_step_plan(int64_t B, int64_t A, int64_t n_tokens, int64_t H, int64_t V, int32_t loss_type,
                             int32_t need_grad, int64_t* plan_out);
/*
 * Optional inputs of the step (pass NULL for none).
 *   tis / clip_mask : the off-policy correction of the reference (off_policy_correction_utils.py:173-262) as
 *     skyb200_off_policy_correction produces it BEFORE the step (it depends only on old / rollout log-probs): `tis`
 *     [B, A] is the detached per-token importance weight that multiplies the loss; `loss_mask` is then the CORRECTED
 *     mask and `clip_mask` [B, A] the mask before the correction, which the clip_ratio metric keeps
 *     (ppo_utils.py:570 vs :579-583).
 */
/*
 *   dw_scatter : fuse the data-parallel reduce-scatter of the lm_head weight gradient into the d_weight GEMM (what
 *     the reference gets from FSDP after the backward: fsdp_strategy.py:151-153, 214-226 - fp32 reduce, gradients
 *     averaged over the ranks, every rank keeps its own parameter shard).  shard[r] is rank r's [V / world, H] fp32
 *     gradient shard AS MAPPED INTO THIS PROCESS (shard[rank] = the local buffer, the others peer memory opened with
 *     skyb200_ipc_open); every accumulator piece is multiplied by `scale` (1 / world) and added into its owner's shard
 *     by the GEMM epilogue - TMA reduce-add on a tensor map over the owner's shard, local or peer (NVLink carries the
 *     reduction to the owner's L2).  The
 *     `d_weight` argument of the step is then ignored (no [V, H] buffer exists) and nothing is ever overwritten: the
 *     caller zeroes its shard, synchronises the ranks (so nobody adds into a shard before it is zeroed), runs the step
 *     on every rank and synchronises again before it reads the shard.  V must be a multiple of 32 * world.
 */
typedef struct {
    int32_t world; /* 2 .. 8 */
    int32_t rank;
    float* shard[8];
    float scale;
    int32_t remote_red; /* 0: peers' pieces through TMA reduce-add on the peer-mapped shard; 1: fp32 vector reductions */
} skyb200_dw_scatter;
typedef struct {
    const float* tis;
    const float* clip_mask;
    int64_t timing;      /* != 0: bracket every GEMM launch with CUDA timing events (see skyb200_policy_step_timing) */
    const skyb200_dw_scatter* dw_scatter; /* NULL: plain local d_weight */
    int64_t reserved[4]; /* must be zero */
} skyb200_step_opts;
int skyb200_policy_step(const void* hidden, const void* weight, const int64_t* labels, const float* old_log_probs,
                        const float* advantages, const float* loss_mask, const float* base_log_probs,
                        const float* entropy_mask, float* log_probs, float* entropy, void* d_hidden, float* d_weight,
                        int32_t accumulate, float* stats, void* workspace, size_t workspace_bytes, int64_t B, int64_t A,
                        int64_t H, int64_t V, const int32_t* cu_seqlens, int64_t n_tokens, float temperature,
                        const skyb200_loss_params* params, const skyb200_step_opts* opts, void* stream);

/*
 * Per-kernel durations of the steps run with opts->timing since the last call, measured LIVE (CUDA events on the
 * caller's stream around each GEMM launch, inside the running step - the power-capped regime the step really sees):
 * out[0..2] = summed ms of the forward / d_hidden / d_weight GEMM launches, out[3..5] = their launch counts.
 * Synchronises on the last recorded event and clears the record.  Per device (call it on the step's device).
 */
int skyb200_policy_step_timing(double* out);

/*
 * Peer mapping of a device buffer between the processes of one node (one process per GPU): the owner exports a handle
 * for a buffer that lies inside a cudaMalloc allocation (a torch tensor does), the other ranks open it.
 * skyb200_ipc_export: handle_out = 64 bytes (cudaIpcMemHandle_t of the allocation), *offset_out = byte offset of `ptr`
 * inside it.  skyb200_ipc_open: maps the allocation into this process with peer access enabled and returns
 * base + offset; skyb200_ipc_close takes the pointer skyb200_ipc_open returned.
 */
int skyb200_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out);
int skyb200_ipc_open(const void* handle, int64_t offset, void** ptr_out);
int skyb200_ipc_close(void* ptr);

/*
 * Same step with the hidden states resident in HOST memory (pinned for an asynchronous copy): hidden_host [N, H]
 * bf16 is streamed into hidden_staging (device, [N, H] bf16, uninitialised on entry, holds the hidden states on
 * return) one token chunk at a time on an internal copy stream, ahead of the GEMMs that read it, so only the first
 * chunk's transfer is exposed.  Everything else as skyb200_policy_step.
 */
int skyb200_policy_step_from_host(const void* hidden_host, void* hidden_staging, const void* weight,
                                  const int64_t* labels, const float* old_log_probs, const float* advantages,
                                  const float* loss_mask, const float* base_log_probs, const float* entropy_mask,
                                  float* log_probs, float* entropy, void* d_hidden, float* d_weight,
                                  int32_t accumulate, float* stats, void* workspace, size_t workspace_bytes, int64_t B,
                                  int64_t A, int64_t H, int64_t V, const int32_t* cu_seqlens, int64_t n_tokens,
                                  float temperature, const skyb200_loss_params* params,
                                  const skyb200_step_opts* opts, void* stream);

/*
 * logprobs_from_logits (torch_utils.py:115-177) and chunked_entropy_from_logits (:59-111) on materialised
 * logits [N, V] (bf16 or fp32), for callers that still hold logits.  logits_dtype: 0 = fp32, 1 = bf16.
 * Any output pointer may be NULL.  entropy_mask [N] fp32 or NULL.  lse [N] = natural-log sum-exp (saved for
 * backward).
 */
int skyb200_logits_logprob_entropy(const void* logits, int32_t logits_dtype, const int64_t* labels,
                                   const float* entropy_mask, float* logprobs, float* entropy, float* lse, int64_t N,
                                   int64_t V, void* stream);

/*
 * Backward of the materialised-logits path.  d_logits (same dtype as logits; may alias `logits`, which is the
 * inplace_backward=True behaviour of the reference's flash-attn cross-entropy, model_wrapper.py:317-321):
 *   d_j = g_lp * (1[j == label] - p_j) - g_ent * p_j * (log p_j + H)
 * g_lp / g_ent [N] fp32, either may be NULL (treated as 0); entropy [N] = UNMASKED forward entropy (iff g_ent).
 */
int skyb200_logits_backward(const void* logits, int32_t logits_dtype, const int64_t* labels, const float* lse,
                            const float* entropy, const float* g_lp, const float* g_ent, void* d_logits, int64_t N,
                            int64_t V, void* stream);

/*
 * Self-test hook: C[M,N] (fp32) = A * B^T through the same tcgen05 mainloop the head uses.
 * a_mn / b_mn: 0 = K-major ([M,K] / [N,K] row-major), 1 = MN-major ([K,M] / [K,N] row-major).
 * ctas: 1 or 2 (cta_group).  bn: 128 or 256.
 */
int skyb200_gemm_selftest(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int32_t a_mn,
                          int32_t b_mn, int32_t ctas, int32_t bn, int32_t accumulate, void* stream);

/*
 * Tuning hook (scripts/gemm_tune.py): the production GEMM configuration with a chosen epilogue and tile order,
 * K-major operands A [M, K], B [N, K].  epi: 0 = fp32 C [M, N], 1 = bf16 C [M, N], 2 = discard (C may be NULL),
 * 3 = head-forward LSE partials (scratch = (3*ceil(N/256)+1)*M floats followed by M int64 labels; C = optional
 * bf16 logits store with pitch round_up(N, 64)).  order / group: see gemm_sm100.cuh decode_tile.
 */
int skyb200_gemm_bench(const void* A, const void* B, void* C, void* scratch, int64_t M, int64_t N, int64_t K,
                       int32_t epi, int32_t group, int32_t order, void* stream);

/*
 * Host-only: the N-group size of the d_weight GEMM's tile order for a chunk of `rows` tokens (the hidden-chunk panels of
 * one group must stay L2-resident while the d(logits) panels stream past; minimises a DRAM-traffic model, see
 * csrc/head.cu dw_group_n).  SKYB200_DW_GROUP_N overrides.
 */
int skyb200_dw_group_plan(int64_t rows, int64_t H, int64_t V);

/*
 * Host-only: the tile widths (columns of one CTA pair's 256-row tile) the three GEMMs of a `rows`-token chunk run with:
 * bn_out[0] = forward (always 256: two accumulators, the epilogue of a tile runs under the next tile's mainloop),
 * bn_out[1] = d_hidden, bn_out[2] = d_weight - 512 (one accumulator in all 512 tensor-memory columns, 25 % fewer
 * operand bytes per flop) unless tile quantisation on the 74 CTA pairs costs more than that gains, or SKYB200_WIDE
 * disables it; bn_out[3] = the N-group size of the d_weight tile order in tiles of that width.
 */
int skyb200_head_tile_plan(int64_t rows, int64_t H, int64_t V, int32_t* bn_out);

/*
 * Tuning hook (scripts/bwd_gemm_probe.py): ONE launch of a backward GEMM of the fused step exactly as the step issues
 * it (operand majors, tile order, L2 hints, lockstep) on caller-provided buffers.  which = 1: d_hidden [rows, H] bf16 =
 * S [rows, V] * weight [V, H]; which = 2: d_weight [V, H] fp32 (+)= S^T * hidden [rows, H].  S = bf16 d(logits) with
 * pitch round_up(V, 64).
 */
int skyb200_head_gemm_bench(int32_t which, const void* hidden, const void* weight, const void* S, void* d_hidden,
                            float* d_weight, int32_t accumulate, int64_t rows, int64_t H, int64_t V, void* stream);

/* host pointer to the last GEMM watchdog record (8 uint32: flag, role, which, stage, parity, tile, block, pad) */
const uint32_t* skyb200_gemm_debug_record(void);

#ifdef __cplusplus
}
#endif
#endif /* SKYB200_H */
