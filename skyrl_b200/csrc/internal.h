// Host-side entry points shared between translation units of libskyb200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/skyb200.h"

namespace skyb200 {

// ---- loss.cu: the loss of a batch evaluated chunk by chunk over flat token ranges
struct LossChunkWs {
    float* row_msum;         // [B]   sum_t mask[b, t]
    float* row_lw;           // [B]   GSPO: masked mean of (lp - olp) per sequence
    float* row_csum;         // [B]   sum_t clip_mask[b, t] (only with an off-policy correction)
    float* total_msum;       // [0] sum(mask), [1] sum(clip_mask)
    unsigned int* counters;  // [4]
    float* partials;         // [max_blocks][8]
};
size_t loss_chunk_ws_bytes(int64_t B, int64_t n_tokens);
void loss_chunk_ws_carve(void* base, int64_t B, int64_t A, LossChunkWs* ws);
int64_t loss_chunk_blocks(int64_t n_tokens);  // blocks (= partial slots) a flat range of n tokens uses

// row mask sums + global mask sum for the whole batch (needs only the mask)
// cu (device, [B + 1], nullable): packed layout - row b owns tokens [cu[b], cu[b + 1]) instead of [b*A, (b+1)*A)
// clip_mask (nullable): the mask the clip_ratio metric is normalised with when an off-policy correction changed
// loss_mask; its sum goes to total_msum[1]
int loss_prepass(const float* loss_mask, const float* clip_mask, const LossChunkWs& ws, int64_t B, int64_t A,
                 const int32_t* cu, cudaStream_t stream);
// GSPO sequence log-ratios for rows [row0, row0 + rows)
int loss_row_logratio(const float* lp, const float* olp, const float* loss_mask, const LossChunkWs& ws, int64_t row0,
                      int64_t rows, int64_t A, cudaStream_t stream);
// per-token pass over flat tokens [t0, t1) of the [B, A] batch; partials go to slots [slot0, slot0 + blocks).
// lp / olp / adv / mask / base / ent are whole-batch arrays; d_lp / d_ent are CHUNK-LOCAL (element t0 at index 0).
// tis / clip_mask (whole-batch, nullable): the detached off-policy weight of each token and the clip_ratio mask.
int loss_flat_range(const float* lp, const float* olp, const float* adv, const float* loss_mask, const float* base,
                    const float* ent, const float* tis, const float* clip_mask, float* d_lp, float* d_ent,
                    const LossChunkWs& ws, int64_t B, int64_t A, int64_t t0, int64_t t1, int64_t slot0,
                    const skyb200_loss_params& p, const int32_t* cu, cudaStream_t stream);
// fold partial slots [0, nslots) into the packed stats vector
int loss_finalize(const LossChunkWs& ws, int64_t nslots, bool has_entropy, bool has_clip_mask, float* stats,
                  const skyb200_loss_params& p, cudaStream_t stream);

}  // namespace skyb200
