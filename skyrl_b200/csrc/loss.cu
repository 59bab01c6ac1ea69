// Policy loss of the PPO/GRPO/GSPO update, forward and backward fused in one memory-bound pass.
//
// Reference arithmetic (skyrl-train/skyrl_train/):
//   ppo_policy_loss            utils/ppo_utils.py:548-586   ("regular", "dual_clip")
//   gspo_policy_loss           utils/ppo_utils.py:661-725
//   reduce_loss                utils/ppo_utils.py:984-1009
//   compute_approx_kl          utils/ppo_utils.py:88-124    (under @torch.no_grad -> no gradient by default)
//   masked_mean                utils/torch_utils.py:180-184 (denominator clamp(min=1))
//   loss assembly              workers/worker.py:852-876
//
// Two launches: (1) per-row mask sums (+ the GSPO sequence log-ratio) and the global mask sum;
// (2) the per-token pass that writes d_log_probs / d_entropy and reduces every scalar the worker reports.
// All reductions are order-deterministic (per-block partials, last block folds them in index order).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "loss_token.cuh"

namespace skyb200 {

constexpr int kLossThreads = 256;
constexpr int kLossVec = 4;
constexpr int kLossIters = 4;  // elements per block = 256 * 4 * 4 = 4096
constexpr int kLossTile = kLossThreads * kLossVec * kLossIters;
constexpr int kMaxMainGrid = 1024;
constexpr int kNumPartials = 8;

struct LossWs {
    float* row_msum;          // [B]
    float* row_lw;            // [B]  GSPO: masked mean of (lp - olp)
    float* row_scale;         // [B]  1 / max(row mask sum, 1) / B: the per-row normaliser of the token pass
    float* total_msum;        // [1]
    unsigned int* counters;   // [2]
    float* partials;          // [num_blocks][kNumPartials]
    float* tmp;               // 148*4*2 floats of scratch for small side reductions
};

static size_t loss_ws_layout(int64_t B, int64_t A, LossWs* ws, void* base) {
    const int64_t chunks = (A + kLossTile - 1) / kLossTile;
    const int64_t blocks = B * chunks + kMaxMainGrid;  // the token pass runs on at most kMaxMainGrid blocks
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    size_t o_rm = take(B * 4), o_lw = take(B * 4), o_rs = take(B * 4), o_tot = take(16), o_cnt = take(16),
           o_part = take(blocks * kNumPartials * 4), o_tmp = take(148 * 4 * 2 * 4 + 64);
    if (ws != nullptr) {
        char* b = static_cast<char*>(base);
        ws->row_msum = reinterpret_cast<float*>(b + o_rm);
        ws->row_lw = reinterpret_cast<float*>(b + o_lw);
        ws->row_scale = reinterpret_cast<float*>(b + o_rs);
        ws->total_msum = reinterpret_cast<float*>(b + o_tot);
        ws->counters = reinterpret_cast<unsigned int*>(b + o_cnt);
        ws->partials = reinterpret_cast<float*>(b + o_part);
        ws->tmp = reinterpret_cast<float*>(b + o_tmp);
    }
    return off;
}

// ---------------------------------------------------------------------------- (1) row statistics
__global__ void __launch_bounds__(kLossThreads)
loss_rowstats_kernel(const float* __restrict__ lp, const float* __restrict__ olp, const float* __restrict__ mask,
                     float* __restrict__ row_msum, float* __restrict__ row_lw, float* __restrict__ total_msum,
                     unsigned int* __restrict__ counter, int B, int64_t A, int need_lw,
                     const int32_t* __restrict__ cu = nullptr, float* __restrict__ row_scale = nullptr,
                     unsigned int* __restrict__ clear = nullptr) {
    __shared__ float sh[34];
    __shared__ bool is_last;
    pdl_launch_dependents();  // the token pass may be scheduled while this pre-pass drains (it waits before reading)
    const int row = blockIdx.x;
    const int64_t base = cu ? static_cast<int64_t>(cu[row]) : static_cast<int64_t>(row) * A;
    if (cu) A = static_cast<int64_t>(cu[row + 1]) - base;  // packed layout: this row's own length
    float ms = 0.f, lw = 0.f;
    const bool al = ((base & 3) == 0) && (A % 4 == 0) && (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 15u) == 0) &&
                    (!need_lw || (((reinterpret_cast<uintptr_t>(lp) | reinterpret_cast<uintptr_t>(olp)) & 15u) == 0));
    if (al && mask != nullptr) {
        // 128-bit loads, four of them in flight per thread (plain loads: the mask stays in L2 for the token pass)
        const float4* m4 = reinterpret_cast<const float4*>(mask + base);
        const float4* l4 = reinterpret_cast<const float4*>(lp + base);
        const float4* o4 = reinterpret_cast<const float4*>(olp + base);
        const int64_t n4 = A / 4;
        if (!need_lw) {
            for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 8 * blockDim.x) {
                float4 m[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int64_t i = i0 + u * blockDim.x;
                    m[u] = (i < n4) ? __ldg(m4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) ms += (m[u].x + m[u].y) + (m[u].z + m[u].w);
            }
        } else {
            for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 2 * blockDim.x) {
                float4 m[2], a[2], b[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int64_t i = i0 + u * blockDim.x;
                    const bool ok = i < n4;
                    m[u] = ok ? __ldg(m4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    a[u] = ok ? __ldg(l4 + i) : m[u];
                    b[u] = ok ? __ldg(o4 + i) : m[u];
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    lw += (a[u].x - b[u].x) * m[u].x + (a[u].y - b[u].y) * m[u].y + (a[u].z - b[u].z) * m[u].z +
                          (a[u].w - b[u].w) * m[u].w;
                    ms += (m[u].x + m[u].y) + (m[u].z + m[u].w);
                }
            }
        }
    } else {
        for (int64_t t = threadIdx.x; t < A; t += blockDim.x) {
            const float m = mask ? mask[base + t] : 1.f;
            ms += m;
            if (need_lw) lw += (lp[base + t] - olp[base + t]) * m;
        }
    }
    ms = block_sum(ms, sh);
    if (need_lw) lw = block_sum(lw, sh);
    if (threadIdx.x == 0) {
        row_msum[row] = ms;
        row_lw[row] = need_lw ? lw / fmaxf(ms, 1.f) : 0.f;
        if (row_scale != nullptr) row_scale[row] = (1.f / fmaxf(ms, 1.f)) / static_cast<float>(B);
        if (clear != nullptr && row == 0) *clear = 0u;  // the token pass's "last block" ticket (no memset node)
        is_last = false;
        if (total_msum != nullptr) {
            __threadfence();
            const unsigned int prev = atomicAdd(counter, 1u);
            is_last = (prev == static_cast<unsigned int>(B) - 1u);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float tot = 0.f;
        for (int i = threadIdx.x; i < B; i += blockDim.x) tot += __ldcg(&row_msum[i]);
        tot = block_sum(tot, sh);
        if (threadIdx.x == 0) {
            *total_msum = tot;
            *counter = 0u;
        }
    }
}

// ---------------------------------------------------------------------------- (2) per-token pass
struct LossArgs {
    const float* lp;
    const float* olp;
    const float* adv;
    const float* mask;
    const float* base;
    const float* ent;
    const float* tis;        // detached per-token weight (off-policy correction) or nullptr
    const float* clip_mask;  // mask of the clip_ratio metric when it differs from `mask`, or nullptr
    float* d_lp;
    float* d_ent;
    float* stats;
    LossWs ws;
    int B;
    int vec4;
    int64_t A;
    skyb200_loss_params p;
};

// One 1024-position tile of one row as a thread sees it: 4 positions (consecutive in the 128-bit path, strided by the
// block size otherwise) of every input, plus the row's normalisers.
constexpr int kMainTile = kLossThreads * 4;
struct TileIn {
    float lp[4], olp[4], adv[4], m[4], bs[4], en[4], ti[4], cm[4];
    int64_t i0;        // flat index of component 0; component u sits at i0 + u * stride
    unsigned ok;       // bit u set = component u is a real position
    int row;
};

__device__ __forceinline__ void ld4(const float* p, int64_t i, float (&dst)[4]) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p + i));
    dst[0] = v.x;
    dst[1] = v.y;
    dst[2] = v.z;
    dst[3] = v.w;
}

// Issues the loads of one tile (inputs only: nothing here depends on the pre-pass).  kExtra = the off-policy
// correction arrays (tis, clip_mask) may be present; without it they cost neither registers nor instructions.
template <bool kVec, bool kExtra>
__device__ __forceinline__ void tile_load(const LossArgs& a, int row, int tile_in_row, bool valid, TileIn& d) {
    d.ok = 0u;
    if (!valid) return;
    const int64_t t0 = static_cast<int64_t>(tile_in_row) * kMainTile;
    const int64_t rbase = static_cast<int64_t>(row) * a.A;
    d.row = row;
    if constexpr (kVec) {
        const int64_t t = t0 + static_cast<int64_t>(threadIdx.x) * 4;
        if (t >= a.A) return;  // consume() skips tiles with ok == 0: nothing else needs a value
        const int64_t i = rbase + t;
        ld4(a.lp, i, d.lp);
        ld4(a.olp, i, d.olp);
        ld4(a.adv, i, d.adv);
        if (a.mask) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(a.mask + i));  // L2 hit after the row-stats pass
            d.m[0] = v.x, d.m[1] = v.y, d.m[2] = v.z, d.m[3] = v.w;
        } else {
            d.m[0] = d.m[1] = d.m[2] = d.m[3] = 1.f;
        }
        if (a.base) ld4(a.base, i, d.bs);
        if (a.ent) ld4(a.ent, i, d.en);
        if constexpr (kExtra) {
            if (a.tis) {
                ld4(a.tis, i, d.ti);
            } else {
                d.ti[0] = d.ti[1] = d.ti[2] = d.ti[3] = 1.f;
            }
            if (a.clip_mask) {
                ld4(a.clip_mask, i, d.cm);
            } else {
                d.cm[0] = d.cm[1] = d.cm[2] = d.cm[3] = -1.f;
            }
        }
        d.ok = 0xFu;
        d.i0 = i;
    } else {
        d.i0 = rbase + t0 + threadIdx.x;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t t = t0 + static_cast<int64_t>(u) * kLossThreads + threadIdx.x;
            d.lp[u] = d.olp[u] = d.adv[u] = d.m[u] = d.bs[u] = d.en[u] = 0.f;
            d.ti[u] = 1.f;
            d.cm[u] = -1.f;
            if (t < a.A) {
                const int64_t i = rbase + t;
                d.ok |= 1u << u;
                d.lp[u] = a.lp[i];
                d.olp[u] = a.olp[i];
                d.adv[u] = a.adv[i];
                d.m[u] = a.mask ? a.mask[i] : 1.f;
                if (a.base) d.bs[u] = a.base[i];
                if (a.ent) d.en[u] = a.ent[i];
                if (a.tis) d.ti[u] = a.tis[i];
                if (a.clip_mask) d.cm[u] = a.clip_mask[i];
            }
        }
    }
}

// Persistent grid (a few blocks per SM): every block walks 1024-position tiles with a grid stride, kInFlight tiles in
// flight per thread (all loads of all of them are issued before the first is consumed), one block reduction at the
// end.  The first round of loads is issued BEFORE the kernel waits for the mask pre-pass (programmatic dependent
// launch): the pre-pass's tail and this kernel's DRAM latency overlap.
template <int kInFlight, int kMinBlocks, int kType, int kBase, int kKl, bool kVec, bool kExtra>
__global__ void __launch_bounds__(kLossThreads, kMinBlocks) loss_main_kernel(LossArgs a) {
    __shared__ float sh[34];
    __shared__ bool is_last;
    const int64_t A = a.A;
    const float Bf = static_cast<float>(a.B);
    const bool gspo = (a.p.loss_type == SKYB200_LOSS_GSPO);
    const bool use_ent_loss = (a.p.use_entropy_loss != 0);
    const int reduction = a.p.reduction;

    LossAcc acc;
    TokenCtx cx;
    cx.gspo = gspo;
    cx.dual = (a.p.loss_type == SKYB200_LOSS_DUAL_CLIP);
    cx.lo = 1.f - a.p.eps_clip_low;
    cx.hi = 1.f + a.p.eps_clip_high;
    cx.seq_ratio = 1.f;
    cx.seq_pass = 1.f;
    cx.clip_c = a.p.clip_ratio_c;
    cx.kl_type = a.p.kl_type;
    cx.loss_type = a.p.loss_type;
    cx.tau_pos = a.p.sapo_tau_pos;
    cx.tau_neg = a.p.sapo_tau_neg;
    cx.has_mask = (a.mask != nullptr);
    cx.has_base = (a.base != nullptr);
    cx.w_scale = (reduction == SKYB200_REDUCE_SUM) ? 1.f : 1.f / (a.p.max_seq_len * Bf);  // the row-independent forms
    float total_m = 0.f, inv_total = 0.f;

    // after the pre-pass: global mask sum = fixed-order fold of its row sums (the same in every block)
    auto after_prepass = [&]() {
        pdl_wait();
        float tot = 0.f;
        for (int i = threadIdx.x; i < a.B; i += blockDim.x) tot += __ldcg(&a.ws.row_msum[i]);
        total_m = block_sum(tot, sh);
        inv_total = 1.f / fmaxf(total_m, 1.f);
        cx.d_ent_val = use_ent_loss ? -a.p.entropy_loss_coef * inv_total : 0.f;
        if (reduction == SKYB200_REDUCE_TOKEN_MEAN) cx.w_scale = inv_total;
    };

    auto consume = [&](const TileIn& d) {
        if (!d.ok) return;
        const float rs = __ldcg(&a.ws.row_scale[d.row]);  // 1 / max(row mask sum, 1) / B
        if (reduction == SKYB200_REDUCE_SEQUENCE_MEAN) cx.w_scale = rs;
        cx.kl_scale = rs;
        cx.kl_grad_scale = a.p.kl_has_grad ? a.p.kl_loss_coef * rs : 0.f;
        if (gspo) {
            const float lw = __ldcg(&a.ws.row_lw[d.row]);
            cx.seq_ratio = expf(fminf(lw, 10.f));
            cx.seq_pass = (lw <= 10.f) ? 1.f : 0.f;
        }
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            g[u] = 0.f;
            if (kVec || (d.ok >> u & 1u)) {  // the 128-bit path handles whole float4s: ok is all or nothing
                g[u] = loss_token<kType, kBase, kKl>(cx, acc, d.lp[u], d.olp[u], d.adv[u], d.m[u], d.bs[u],
                                                     kExtra ? d.ti[u] : 1.f, kExtra ? d.cm[u] : -1.f);
                if (a.ent) acc.s_ent = fmaf(d.en[u], d.m[u], acc.s_ent);
            }
        }
        if constexpr (kVec) {
            if (a.d_lp) __stcs(reinterpret_cast<float4*>(a.d_lp + d.i0), make_float4(g[0], g[1], g[2], g[3]));
            if (a.ent && a.d_ent)
                __stcs(reinterpret_cast<float4*>(a.d_ent + d.i0),
                       make_float4(cx.d_ent_val * d.m[0], cx.d_ent_val * d.m[1], cx.d_ent_val * d.m[2],
                                   cx.d_ent_val * d.m[3]));
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (d.ok >> u & 1u) {
                    const int64_t i = d.i0 + static_cast<int64_t>(u) * kLossThreads;
                    if (a.d_lp) a.d_lp[i] = g[u];
                    if (a.ent && a.d_ent) a.d_ent[i] = cx.d_ent_val * d.m[u];
                }
            }
        }
    };

    // (row, tile-in-row) advance incrementally by the grid stride: no 64-bit divisions inside the loop
    const unsigned tiles_per_row = static_cast<unsigned>((A + kMainTile - 1) / kMainTile);
    const unsigned step_rows = gridDim.x / tiles_per_row, step_tiles = gridDim.x % tiles_per_row;
    unsigned row = blockIdx.x / tiles_per_row, tir = blockIdx.x % tiles_per_row;
    auto advance = [&]() {
        row += step_rows;
        tir += step_tiles;
        if (tir >= tiles_per_row) {
            tir -= tiles_per_row;
            ++row;
        }
    };
    if constexpr (kInFlight == 2) {
        // Rolling two-buffer pipeline: while one tile is consumed the other one's loads (and, right after the consume,
        // the reload of the buffer just freed) are in flight - the DRAM latency of a round of loads is hidden behind
        // the arithmetic of the other tile instead of being paid once per round.
        TileIn d[2];
        bool live[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            live[u] = row < static_cast<unsigned>(a.B);  // block-uniform
            tile_load<kVec, kExtra>(a, static_cast<int>(row), static_cast<int>(tir), live[u], d[u]);
            advance();
        }
        after_prepass();
        while (live[0] || live[1]) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (live[u]) consume(d[u]);
                live[u] = row < static_cast<unsigned>(a.B);
                tile_load<kVec, kExtra>(a, static_cast<int>(row), static_cast<int>(tir), live[u], d[u]);
                advance();
            }
        }
    } else {
        bool waited = false;
        while (row < static_cast<unsigned>(a.B)) {
            TileIn d[kInFlight];
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) {
                tile_load<kVec, kExtra>(a, static_cast<int>(row), static_cast<int>(tir), row < static_cast<unsigned>(a.B), d[u]);
                advance();
            }
            if (!waited) {  // block-uniform
                after_prepass();
                waited = true;
            }
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) consume(d[u]);
        }
        if (!waited) after_prepass();
    }
    float s_loss = acc.s_loss, s_clip = acc.s_clip, s_kl = acc.s_kl, s_ent = acc.s_ent, s_ratio = acc.s_ratio;
    float r_max = acc.r_max, r_min = acc.r_min;

    // block partials -> workspace, last block folds them in a fixed order
    const int nblk = gridDim.x * gridDim.y;
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    s_loss = block_sum(s_loss, sh);
    s_clip = block_sum(s_clip, sh);
    s_kl = block_sum(s_kl, sh);
    s_ent = block_sum(s_ent, sh);
    s_ratio = block_sum(s_ratio, sh);
    r_max = block_max(r_max, sh);
    r_min = block_min(r_min, sh);
    if (threadIdx.x == 0) {
        float* p = a.ws.partials + static_cast<int64_t>(blk) * kNumPartials;
        p[0] = s_loss;
        p[1] = s_clip;
        p[2] = s_kl;
        p[3] = s_ent;
        p[4] = s_ratio;
        p[5] = r_max;
        p[6] = r_min;
        __threadfence();
        const unsigned int prev = atomicAdd(a.ws.counters + 1, 1u);
        is_last = (prev == static_cast<unsigned int>(nblk) - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float f_loss = 0.f, f_clip = 0.f, f_kl = 0.f, f_ent = 0.f, f_ratio = 0.f, f_max = -INFINITY, f_min = INFINITY;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        const float* p = a.ws.partials + static_cast<int64_t>(b) * kNumPartials;
        f_loss += __ldcg(p + 0);
        f_clip += __ldcg(p + 1);
        f_kl += __ldcg(p + 2);
        f_ent += __ldcg(p + 3);
        f_ratio += __ldcg(p + 4);
        f_max = fmaxf(f_max, __ldcg(p + 5));
        f_min = fminf(f_min, __ldcg(p + 6));
    }
    f_loss = block_sum(f_loss, sh);
    f_clip = block_sum(f_clip, sh);
    f_kl = block_sum(f_kl, sh);
    f_ent = block_sum(f_ent, sh);
    f_ratio = block_sum(f_ratio, sh);
    f_max = block_max(f_max, sh);
    f_min = block_min(f_min, sh);
    if (threadIdx.x == 0) {
        const float clip_den = a.clip_mask ? fmaxf(__ldcg(a.ws.total_msum + 1), 1.f) : fmaxf(total_m, 1.f);
        const float clip_ratio = f_clip / clip_den;
        const float ent_mean = f_ent * inv_total;
        float final_loss = f_loss + f_kl * a.p.kl_loss_coef;
        if (use_ent_loss && a.ent != nullptr) final_loss -= ent_mean * a.p.entropy_loss_coef;
        a.stats[SKYB200_STAT_POLICY_LOSS] = f_loss;
        a.stats[SKYB200_STAT_CLIP_RATIO] = clip_ratio;
        a.stats[SKYB200_STAT_KL_LOSS] = f_kl;
        a.stats[SKYB200_STAT_ENTROPY] = ent_mean;
        a.stats[SKYB200_STAT_FINAL_LOSS] = final_loss;
        a.stats[SKYB200_STAT_MASK_SUM] = total_m;
        a.stats[SKYB200_STAT_RATIO_MEAN] = f_ratio * inv_total;
        a.stats[SKYB200_STAT_RATIO_MAX] = f_max;
        a.stats[SKYB200_STAT_RATIO_MIN] = f_min;
        for (int i = SKYB200_STAT_RATIO_MIN + 1; i < SKYB200_LOSS_NSTATS; ++i) a.stats[i] = 0.f;
        a.ws.counters[1] = 0u;
    }
}

// ---------------------------------------------------------------------------- chunked (flat token range) variant
// Used by the fused step: the loss of a [B, A] batch is evaluated chunk by chunk as the head produces log-probs.
// Normalisers come from a mask-only pre-pass, so a chunk never needs another chunk's log-probs
// (GSPO needs whole sequences per chunk; the caller aligns its chunks to rows for it).
constexpr int kFlatTile = kLossThreads * 4 * 2;  // 2048 tokens per block

struct FlatArgs {
    const float* lp;
    const float* olp;
    const float* adv;
    const float* mask;
    const float* base;
    const float* ent;
    const float* tis;        // detached per-token weight (off-policy correction) or nullptr
    const float* clip_mask;  // mask of the clip_ratio metric when it differs from `mask`, or nullptr
    float* d_lp;
    float* d_ent;
    LossChunkWs ws;
    const int32_t* cu;  // packed layout: row offsets [B + 1], else nullptr
    int B;
    int vec4;
    int64_t A;
    int64_t t0, t1;
    int64_t out_base;  // d_lp / d_ent are chunk-local: element i of the batch is written at [i - out_base]
    int64_t slot0;
    skyb200_loss_params p;
};

// packed layout: the row that owns flat token i = the largest r with cu[r] <= i
__device__ __forceinline__ int row_of_token(const int32_t* __restrict__ cu, int B, int64_t i) {
    int lo = 0, hi = B;  // invariant: cu[lo] <= i < cu[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(__ldg(cu + mid)) <= i)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// kType: -2 = clipped-surrogate family (the three north-star losses, no dead branches of the others), -1 = any type
template <int kType>
__global__ void __launch_bounds__(kLossThreads) loss_flat_kernel(FlatArgs a) {
    __shared__ float sh[34];
    const int64_t A = a.A;
    const float total_m = __ldcg(a.ws.total_msum);
    const float Bf = static_cast<float>(a.B);
    const float inv_total = 1.f / fmaxf(total_m, 1.f);
    const bool gspo = (a.p.loss_type == SKYB200_LOSS_GSPO);
    const bool use_ent_loss = (a.p.use_entropy_loss != 0);
    LossAcc acc;
    TokenCtx cx;
    cx.gspo = gspo;
    cx.dual = (a.p.loss_type == SKYB200_LOSS_DUAL_CLIP);
    cx.lo = 1.f - a.p.eps_clip_low;
    cx.hi = 1.f + a.p.eps_clip_high;
    cx.seq_ratio = 1.f;
    cx.seq_pass = 1.f;
    cx.clip_c = a.p.clip_ratio_c;
    cx.kl_type = a.p.kl_type;
    cx.loss_type = a.p.loss_type;
    cx.tau_pos = a.p.sapo_tau_pos;
    cx.tau_neg = a.p.sapo_tau_neg;
    cx.has_mask = (a.mask != nullptr);
    cx.has_base = (a.base != nullptr);
    cx.d_ent_val = use_ent_loss ? -a.p.entropy_loss_coef * inv_total : 0.f;

    auto set_row = [&](int64_t row) {
        const float row_len = a.cu ? static_cast<float>(a.cu[row + 1] - a.cu[row]) : static_cast<float>(A);
        const float row_m = a.mask ? __ldcg(&a.ws.row_msum[row]) : row_len;
        const float inv_row = 1.f / fmaxf(row_m, 1.f);
        if (a.p.reduction == SKYB200_REDUCE_TOKEN_MEAN)
            cx.w_scale = inv_total;
        else if (a.p.reduction == SKYB200_REDUCE_SEQUENCE_MEAN)
            cx.w_scale = inv_row / Bf;
        else if (a.p.reduction == SKYB200_REDUCE_SUM)
            cx.w_scale = 1.f;
        else
            cx.w_scale = 1.f / (a.p.max_seq_len * Bf);
        cx.kl_scale = inv_row / Bf;
        cx.kl_grad_scale = a.p.kl_has_grad ? a.p.kl_loss_coef * cx.kl_scale : 0.f;
        if (gspo) {
            const float lw = __ldcg(&a.ws.row_lw[row]);
            cx.seq_ratio = expf(fminf(lw, 10.f));
            cx.seq_pass = (lw <= 10.f) ? 1.f : 0.f;
        }
    };

    const int64_t base_t = a.t0 + static_cast<int64_t>(blockIdx.x) * kFlatTile;
    if (a.vec4) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int64_t i = base_t + (static_cast<int64_t>(it) * kLossThreads + threadIdx.x) * 4;
            if (i < a.t1) {
                set_row(i / A);  // A % 4 == 0: the four tokens share a row
                const float4 lp = ldg_stream(reinterpret_cast<const float4*>(a.lp + i));
                const float4 olp = ldg_stream(reinterpret_cast<const float4*>(a.olp + i));
                const float4 adv = ldg_stream(reinterpret_cast<const float4*>(a.adv + i));
                float4 m = make_float4(1.f, 1.f, 1.f, 1.f), bs = make_float4(0.f, 0.f, 0.f, 0.f), en = bs;
                if (a.mask) m = ldg_stream(reinterpret_cast<const float4*>(a.mask + i));
                if (a.base) bs = ldg_stream(reinterpret_cast<const float4*>(a.base + i));
                if (a.ent) en = ldg_stream(reinterpret_cast<const float4*>(a.ent + i));
                float4 ti = make_float4(1.f, 1.f, 1.f, 1.f), cm = make_float4(-1.f, -1.f, -1.f, -1.f);
                if (a.tis) ti = ldg_stream(reinterpret_cast<const float4*>(a.tis + i));
                if (a.clip_mask) cm = ldg_stream(reinterpret_cast<const float4*>(a.clip_mask + i));
                float4 g;
                g.x = loss_token<kType>(cx, acc, lp.x, olp.x, adv.x, m.x, bs.x, ti.x, cm.x);
                g.y = loss_token<kType>(cx, acc, lp.y, olp.y, adv.y, m.y, bs.y, ti.y, cm.y);
                g.z = loss_token<kType>(cx, acc, lp.z, olp.z, adv.z, m.z, bs.z, ti.z, cm.z);
                g.w = loss_token<kType>(cx, acc, lp.w, olp.w, adv.w, m.w, bs.w, ti.w, cm.w);
                if (a.d_lp) *reinterpret_cast<float4*>(a.d_lp + (i - a.out_base)) = g;
                if (a.ent) {
                    acc.s_ent = fmaf(en.x, m.x, acc.s_ent);
                    acc.s_ent = fmaf(en.y, m.y, acc.s_ent);
                    acc.s_ent = fmaf(en.z, m.z, acc.s_ent);
                    acc.s_ent = fmaf(en.w, m.w, acc.s_ent);
                    if (a.d_ent)
                        *reinterpret_cast<float4*>(a.d_ent + (i - a.out_base)) =
                            make_float4(cx.d_ent_val * m.x, cx.d_ent_val * m.y, cx.d_ent_val * m.z, cx.d_ent_val * m.w);
                }
            }
        }
    } else {
#pragma unroll 2
        for (int it = 0; it < 8; ++it) {
            const int64_t i = base_t + static_cast<int64_t>(it) * kLossThreads + threadIdx.x;
            if (i < a.t1) {
                set_row(a.cu ? row_of_token(a.cu, a.B, i) : i / A);
                const float m = a.mask ? a.mask[i] : 1.f;
                const float g = loss_token<kType>(cx, acc, a.lp[i], a.olp[i], a.adv[i], m, a.base ? a.base[i] : 0.f,
                                                  a.tis ? a.tis[i] : 1.f, a.clip_mask ? a.clip_mask[i] : -1.f);
                if (a.d_lp) a.d_lp[i - a.out_base] = g;
                if (a.ent) {
                    acc.s_ent = fmaf(a.ent[i], m, acc.s_ent);
                    if (a.d_ent) a.d_ent[i - a.out_base] = cx.d_ent_val * m;
                }
            }
        }
    }
    const float s_loss = block_sum(acc.s_loss, sh);
    const float s_clip = block_sum(acc.s_clip, sh);
    const float s_kl = block_sum(acc.s_kl, sh);
    const float s_ent = block_sum(acc.s_ent, sh);
    const float s_ratio = block_sum(acc.s_ratio, sh);
    const float r_max = block_max(acc.r_max, sh);
    const float r_min = block_min(acc.r_min, sh);
    if (threadIdx.x == 0) {
        float* p = a.ws.partials + (a.slot0 + blockIdx.x) * kNumPartials;
        p[0] = s_loss;
        p[1] = s_clip;
        p[2] = s_kl;
        p[3] = s_ent;
        p[4] = s_ratio;
        p[5] = r_max;
        p[6] = r_min;
    }
}

__global__ void __launch_bounds__(kLossThreads)
loss_finalize_kernel(LossChunkWs ws, int64_t nslots, int has_entropy, int has_clip_mask, float* __restrict__ stats,
                     skyb200_loss_params p) {
    __shared__ float sh[34];
    float f_loss = 0.f, f_clip = 0.f, f_kl = 0.f, f_ent = 0.f, f_ratio = 0.f, f_max = -INFINITY, f_min = INFINITY;
    for (int64_t b = threadIdx.x; b < nslots; b += blockDim.x) {
        const float* q = ws.partials + b * kNumPartials;
        f_loss += q[0];
        f_clip += q[1];
        f_kl += q[2];
        f_ent += q[3];
        f_ratio += q[4];
        f_max = fmaxf(f_max, q[5]);
        f_min = fminf(f_min, q[6]);
    }
    f_loss = block_sum(f_loss, sh);
    f_clip = block_sum(f_clip, sh);
    f_kl = block_sum(f_kl, sh);
    f_ent = block_sum(f_ent, sh);
    f_ratio = block_sum(f_ratio, sh);
    f_max = block_max(f_max, sh);
    f_min = block_min(f_min, sh);
    if (threadIdx.x == 0) {
        const float total_m = *ws.total_msum;
        const float inv_total = 1.f / fmaxf(total_m, 1.f);
        const float ent_mean = f_ent * inv_total;
        float final_loss = f_loss + f_kl * p.kl_loss_coef;
        if (p.use_entropy_loss && has_entropy) final_loss -= ent_mean * p.entropy_loss_coef;
        stats[SKYB200_STAT_POLICY_LOSS] = f_loss;
        // the clip_ratio metric keeps the mask it had BEFORE the off-policy correction (ppo_utils.py:570 vs :579-583)
        stats[SKYB200_STAT_CLIP_RATIO] = has_clip_mask ? f_clip / fmaxf(ws.total_msum[1], 1.f) : f_clip * inv_total;
        stats[SKYB200_STAT_KL_LOSS] = f_kl;
        stats[SKYB200_STAT_ENTROPY] = ent_mean;
        stats[SKYB200_STAT_FINAL_LOSS] = final_loss;
        stats[SKYB200_STAT_MASK_SUM] = total_m;
        stats[SKYB200_STAT_RATIO_MEAN] = f_ratio * inv_total;
        stats[SKYB200_STAT_RATIO_MAX] = f_max;
        stats[SKYB200_STAT_RATIO_MIN] = f_min;
        for (int i = SKYB200_STAT_RATIO_MIN + 1; i < SKYB200_LOSS_NSTATS; ++i) stats[i] = 0.f;
    }
}

size_t loss_chunk_ws_bytes(int64_t B, int64_t n_tokens) {
    const int64_t slots = loss_chunk_blocks(n_tokens) + 4 * 64;  // every chunk rounds its block count up
    return static_cast<size_t>(3 * ((B * 4 + 255) / 256 * 256) + 512 + (slots + B) * kNumPartials * 4 + 512);
}
int64_t loss_chunk_blocks(int64_t n_tokens) { return (n_tokens + kFlatTile - 1) / kFlatTile; }

void loss_chunk_ws_carve(void* base, int64_t B, int64_t A, LossChunkWs* ws) {
    (void)A;
    char* b = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
    const size_t rb = static_cast<size_t>((B * 4 + 255) / 256 * 256);
    ws->row_msum = reinterpret_cast<float*>(b);
    ws->row_lw = reinterpret_cast<float*>(b + rb);
    ws->row_csum = reinterpret_cast<float*>(b + 2 * rb);
    ws->total_msum = reinterpret_cast<float*>(b + 3 * rb);
    ws->counters = reinterpret_cast<unsigned int*>(b + 3 * rb + 64);
    ws->partials = reinterpret_cast<float*>(b + 3 * rb + 256);
}

int loss_prepass(const float* loss_mask, const float* clip_mask, const LossChunkWs& ws, int64_t B, int64_t A,
                 const int32_t* cu, cudaStream_t stream) {
    SKYB200_CUDA_CHECK(cudaMemsetAsync(ws.counters, 0, 16, stream));
    if (clip_mask != nullptr) {  // sum(clip_mask) -> total_msum[1] (the last block resets the counter)
        loss_rowstats_kernel<<<(unsigned)B, kLossThreads, 0, stream>>>(nullptr, nullptr, clip_mask, ws.row_csum,
                                                                       ws.row_lw, ws.total_msum + 1, ws.counters,
                                                                       (int)B, A, 0, cu);
        int rc = check_launch("loss_rowstats_kernel(clip mask)");
        if (rc) return rc;
    }
    loss_rowstats_kernel<<<(unsigned)B, kLossThreads, 0, stream>>>(nullptr, nullptr, loss_mask, ws.row_msum, ws.row_lw,
                                                                   ws.total_msum, ws.counters, (int)B, A, 0, cu);
    return check_launch("loss_rowstats_kernel(prepass)");
}

int loss_row_logratio(const float* lp, const float* olp, const float* loss_mask, const LossChunkWs& ws, int64_t row0,
                      int64_t rows, int64_t A, cudaStream_t stream) {
    loss_rowstats_kernel<<<(unsigned)rows, kLossThreads, 0, stream>>>(
        lp + row0 * A, olp + row0 * A, loss_mask ? loss_mask + row0 * A : nullptr, ws.row_msum + row0, ws.row_lw + row0,
        nullptr, ws.counters, (int)rows, A, 1);
    return check_launch("loss_rowstats_kernel(gspo)");
}

int loss_flat_range(const float* lp, const float* olp, const float* adv, const float* loss_mask, const float* base,
                    const float* ent, const float* tis, const float* clip_mask, float* d_lp, float* d_ent,
                    const LossChunkWs& ws, int64_t B, int64_t A, int64_t t0, int64_t t1, int64_t slot0,
                    const skyb200_loss_params& p, const int32_t* cu, cudaStream_t stream) {
    if (t1 <= t0) return SKYB200_OK;
    FlatArgs a;
    a.cu = cu;
    a.lp = lp;
    a.olp = olp;
    a.adv = adv;
    a.mask = loss_mask;
    a.base = base;
    a.ent = ent;
    a.tis = tis;
    a.clip_mask = clip_mask;
    a.d_lp = d_lp;
    a.d_ent = d_ent;
    a.ws = ws;
    a.B = (int)B;
    a.A = A;
    a.t0 = t0;
    a.t1 = t1;
    a.out_base = t0;  // the caller passes chunk-local d_lp / d_ent buffers
    a.slot0 = slot0;
    a.p = p;
    auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    a.vec4 = (cu == nullptr) && (A % 4 == 0) && (t0 % 4 == 0) && al16(lp) && al16(olp) && al16(adv) && al16(loss_mask) && al16(base) &&
             al16(ent) && al16(d_lp) && al16(d_ent) && al16(tis) && al16(clip_mask);
    const unsigned grid = (unsigned)loss_chunk_blocks(t1 - t0);
    if (p.loss_type <= SKYB200_LOSS_GSPO)
        loss_flat_kernel<-2><<<grid, kLossThreads, 0, stream>>>(a);
    else
        loss_flat_kernel<-1><<<grid, kLossThreads, 0, stream>>>(a);
    return check_launch("loss_flat_kernel");
}

int loss_finalize(const LossChunkWs& ws, int64_t nslots, bool has_entropy, bool has_clip_mask, float* stats,
                  const skyb200_loss_params& p, cudaStream_t stream) {
    loss_finalize_kernel<<<1, kLossThreads, 0, stream>>>(ws, nslots, has_entropy ? 1 : 0, has_clip_mask ? 1 : 0, stats,
                                                         p);
    return check_launch("loss_finalize_kernel");
}

// ---------------------------------------------------------------------------- critic (value) loss
// ppo_critic_loss (ppo_utils.py:175-193): 0.5 * mean_b masked_mean_t max((clip(v) - R)^2, (v - R)^2),
// clipfrac = masked_mean(1[surr1 > surr2]) over the whole batch.  Forward + d loss / d values in one pass.
struct CriticArgs {
    const float* values;
    const float* old_values;
    const float* returns;
    const float* mask;
    float* d_values;
    float* stats;  // [0] loss, [1] clipfrac, [2] mask_sum
    LossWs ws;
    int B;
    int64_t A;
    float value_clip;
    int use_clip;
};

__global__ void __launch_bounds__(kLossThreads) critic_main_kernel(CriticArgs a) {
    __shared__ float sh[34];
    __shared__ bool is_last;
    const int row = blockIdx.y;
    const int64_t rbase = static_cast<int64_t>(row) * a.A;
    const float total_m = __ldcg(a.ws.total_msum);
    const float row_m = a.mask ? __ldcg(&a.ws.row_msum[row]) : static_cast<float>(a.A);
    const float w_row = 0.5f / (fmaxf(row_m, 1.f) * static_cast<float>(a.B));
    float s_loss = 0.f, s_clip = 0.f;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * kLossTile;
#pragma unroll 4
    for (int it = 0; it < kLossIters * kLossVec; ++it) {
        const int64_t t = t0 + static_cast<int64_t>(it) * kLossThreads + threadIdx.x;
        if (t < a.A) {
            const int64_t i = rbase + t;
            const float v = a.values[i], R = a.returns[i];
            const float m = a.mask ? a.mask[i] : 1.f;
            const float e2 = v - R;
            float loss = e2 * e2, g = 2.f * e2;
            if (a.use_clip) {
                const float ov = a.old_values[i];
                const float dv = v - ov;
                const float vc = ov + fminf(fmaxf(dv, -a.value_clip), a.value_clip);
                const float pass = (dv >= -a.value_clip && dv <= a.value_clip) ? 1.f : 0.f;
                const float e1 = vc - R;
                const float s1 = e1 * e1, s2 = e2 * e2;
                const float g1 = 2.f * e1 * pass, g2 = 2.f * e2;
                loss = fmaxf(s1, s2);
                g = (s1 > s2) ? g1 : ((s1 < s2) ? g2 : 0.5f * (g1 + g2));  // torch.max splits ties 50/50
                s_clip = fmaf((s1 > s2) ? 1.f : 0.f, m, s_clip);
            }
            const float w = m * w_row;
            s_loss = fmaf(loss, w, s_loss);
            if (a.d_values) a.d_values[i] = g * w;
        }
    }
    const int nblk = gridDim.x * gridDim.y;
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    s_loss = block_sum(s_loss, sh);
    s_clip = block_sum(s_clip, sh);
    if (threadIdx.x == 0) {
        float* p = a.ws.partials + static_cast<int64_t>(blk) * kNumPartials;
        p[0] = s_loss;
        p[1] = s_clip;
        __threadfence();
        const unsigned int prev = atomicAdd(a.ws.counters + 1, 1u);
        is_last = (prev == static_cast<unsigned int>(nblk) - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float f_loss = 0.f, f_clip = 0.f;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        const float* p = a.ws.partials + static_cast<int64_t>(b) * kNumPartials;
        f_loss += __ldcg(p + 0);
        f_clip += __ldcg(p + 1);
    }
    f_loss = block_sum(f_loss, sh);
    f_clip = block_sum(f_clip, sh);
    if (threadIdx.x == 0) {
        const float denom = a.mask ? fmaxf(total_m, 1.f) : static_cast<float>(a.B) * static_cast<float>(a.A);
        a.stats[0] = f_loss;
        a.stats[1] = f_clip / denom;
        a.stats[2] = total_m;
        a.stats[3] = 0.f;
        a.ws.counters[1] = 0u;
    }
}

// ---------------------------------------------------------------------------- small helpers
__global__ void scale_to_sum_kernel(const float* mean, float* out, float n) { *out = *mean * n; }

__global__ void scale_inplace_kernel(float* __restrict__ x, const float* __restrict__ s, int64_t n) {
    const float k = __ldg(s);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        x[i] *= k;
}

__global__ void __launch_bounds__(256) masked_sum_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                         float* __restrict__ partials, int64_t n) {
    __shared__ float sh[34];
    float sx = 0.f, sm = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float mm = m ? m[i] : 1.f;
        sx = fmaf(x[i], mm, sx);
        sm += mm;
    }
    sx = block_sum(sx, sh);
    sm = block_sum(sm, sh);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = sx;
        partials[2 * blockIdx.x + 1] = sm;
    }
}
__global__ void __launch_bounds__(256) masked_mean_final_kernel(const float* __restrict__ partials, int nblk,
                                                                float* __restrict__ out) {
    __shared__ float sh[34];
    float sx = 0.f, sm = 0.f;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
        sx += partials[2 * i];
        sm += partials[2 * i + 1];
    }
    sx = block_sum(sx, sh);
    sm = block_sum(sm, sh);
    if (threadIdx.x == 0) *out = sx / fmaxf(sm, 1.f);
}

__global__ void approx_kl_kernel(const float* __restrict__ lp, const float* __restrict__ base,
                                 const float* __restrict__ mask, float* __restrict__ out, int64_t n, int type) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float dk;
        float k = kl_estimate(lp[i], base[i], type, &dk);
        if (mask) k *= mask[i];
        out[i] = k;
    }
}

static int grid_for(int64_t n, int threads, int per_sm) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace skyb200

using namespace skyb200;

extern "C" size_t skyb200_policy_loss_workspace(int64_t B, int64_t A) {
    return loss_ws_layout(B, A, nullptr, nullptr) + 256;
}

extern "C" int skyb200_policy_loss(const float* log_probs, const float* old_log_probs, const float* advantages,
                                   const float* loss_mask, const float* base_log_probs, const float* entropy,
                                   float* d_log_probs, float* d_entropy, float* stats, void* workspace,
                                   size_t workspace_bytes, int64_t B, int64_t A, const skyb200_loss_params* params,
                                   const float* tis, const float* clip_mask, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(log_probs && old_log_probs && advantages && stats && workspace && params, "null pointer");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < 65536, "need 0 < B < 65536 and A > 0");
    SKYB200_REQUIRE(params->loss_type >= 0 && params->loss_type <= SKYB200_LOSS_LAST, "bad loss_type");
    SKYB200_REQUIRE(params->reduction >= 0 && params->reduction <= SKYB200_REDUCE_SUM, "bad reduction");
    SKYB200_REQUIRE(params->loss_type != SKYB200_LOSS_SAPO || (params->sapo_tau_pos > 0.f && params->sapo_tau_neg > 0.f),
                    "sapo needs positive tau_pos / tau_neg");
    SKYB200_REQUIRE(params->kl_type >= 0 && params->kl_type <= SKYB200_KL_K3, "bad kl_type");
    SKYB200_REQUIRE(params->reduction != SKYB200_REDUCE_SEQ_MEAN_TOKEN_SUM_NORM || params->max_seq_len > 0.f,
                    "max_seq_len must be provided for seq_mean_token_sum_norm");
    if (workspace_bytes < skyb200_policy_loss_workspace(B, A))
        return fail(SKYB200_EWORKSPACE, "policy_loss workspace too small");
    LossWs ws;
    // 256-byte align the workspace base
    char* wbase = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    loss_ws_layout(B, A, &ws, wbase);
    // no memset node and no global counter in the pre-pass: it writes the per-row mask sums / normalisers and clears
    // the token pass's ticket; the token pass folds the global mask sum from the row sums itself
    const int need_lw = params->loss_type == SKYB200_LOSS_GSPO;
    loss_rowstats_kernel<<<(unsigned)B, kLossThreads, 0, stream>>>(log_probs, old_log_probs, loss_mask, ws.row_msum,
                                                                   ws.row_lw, nullptr, ws.counters, (int)B, A,
                                                                   need_lw, nullptr, ws.row_scale, ws.counters + 1);
    int rc = check_launch("loss_rowstats_kernel");
    if (rc) return rc;
    if (clip_mask != nullptr) {  // sum(clip_mask) -> total_msum[1]
        rc = skyb200_masked_mean(clip_mask, nullptr, ws.total_msum + 2, ws.tmp, 148 * 4 * 2 * 4, B * A, stream_);
        if (rc) return rc;
        scale_to_sum_kernel<<<1, 1, 0, stream>>>(ws.total_msum + 2, ws.total_msum + 1, static_cast<float>(B * A));
        rc = check_launch("scale_to_sum_kernel");
        if (rc) return rc;
    }
    LossArgs a;
    a.tis = tis;
    a.clip_mask = clip_mask;
    a.lp = log_probs;
    a.olp = old_log_probs;
    a.adv = advantages;
    a.mask = loss_mask;
    a.base = base_log_probs;
    a.ent = entropy;
    a.d_lp = d_log_probs;
    a.d_ent = d_entropy;
    a.stats = stats;
    a.ws = ws;
    a.B = (int)B;
    a.A = A;
    a.p = *params;
    auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    a.vec4 = (A % 4 == 0) && al16(log_probs) && al16(old_log_probs) && al16(advantages) && al16(loss_mask) &&
             al16(base_log_probs) && al16(entropy) && al16(d_log_probs) && al16(d_entropy) && al16(tis) &&
             al16(clip_mask);
    const int64_t tiles = ((A + kLossThreads * 4 - 1) / (kLossThreads * 4)) * B;
    // 128-bit path: 3 blocks per SM with two tiles (2 x 5 x 16 B per thread) in flight; scalar path: 4 blocks, one tile
    const int per_sm = a.vec4 ? 3 : 4;
    const unsigned grid = (unsigned)std::max<int64_t>(
        1, std::min<int64_t>(tiles, std::min<int64_t>((int64_t)num_sms() * per_sm, (int64_t)kMaxMainGrid)));
    // instantiated per loss type x {no base, base with the reference's value-only k3 KL term, base (general)} x
    // {128-bit path without / with the off-policy-correction arrays, scalar path}: dead branches cost throughput
    const bool k3_value_only = (params->kl_type == SKYB200_KL_K3) && (params->kl_has_grad == 0);
    const bool extra = (tis != nullptr) || (clip_mask != nullptr);
#define SKYB200_LOSS_LAUNCH(T, BASE, KL)                                                                                \
    (a.vec4 ? (extra ? launch_pdl(loss_main_kernel<2, 3, T, BASE, KL, true, true>, dim3(grid), dim3(kLossThreads), 0,   \
                                  stream, a)                                                                            \
                     : launch_pdl(loss_main_kernel<2, 3, T, BASE, KL, true, false>, dim3(grid), dim3(kLossThreads), 0,  \
                                  stream, a))                                                                           \
            : launch_pdl(loss_main_kernel<1, 4, T, BASE, KL, false, true>, dim3(grid), dim3(kLossThreads), 0, stream, a))
#define SKYB200_LOSS_CASE(T)                                                            \
    case T:                                                                             \
        if (!base_log_probs)                                                            \
            SKYB200_CUDA_CHECK(SKYB200_LOSS_LAUNCH(T, 0, -1));                          \
        else if (k3_value_only)                                                         \
            SKYB200_CUDA_CHECK(SKYB200_LOSS_LAUNCH(T, 1, SKYB200_KL_K3));               \
        else                                                                            \
            SKYB200_CUDA_CHECK(SKYB200_LOSS_LAUNCH(T, 1, -1));                          \
        break;
    switch (params->loss_type) {
        SKYB200_LOSS_CASE(SKYB200_LOSS_REGULAR)
        SKYB200_LOSS_CASE(SKYB200_LOSS_DUAL_CLIP)
        SKYB200_LOSS_CASE(SKYB200_LOSS_GSPO)
        SKYB200_LOSS_CASE(SKYB200_LOSS_SAPO)
        SKYB200_LOSS_CASE(SKYB200_LOSS_CISPO)
        SKYB200_LOSS_CASE(SKYB200_LOSS_CROSS_ENTROPY)
        SKYB200_LOSS_CASE(SKYB200_LOSS_IMPORTANCE_SAMPLING)
        SKYB200_LOSS_CASE(SKYB200_LOSS_KL_COV)
        default:
            return fail(SKYB200_EINVAL, "bad loss_type %d", params->loss_type);
    }
#undef SKYB200_LOSS_CASE
#undef SKYB200_LOSS_LAUNCH
    return check_launch("loss_main_kernel");
}

extern "C" int skyb200_scale_inplace(float* x, const float* scale_dev, int64_t n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n == 0) return SKYB200_OK;
    SKYB200_REQUIRE(x && scale_dev, "null pointer");
    scale_inplace_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(x, scale_dev, n);
    return check_launch("scale_inplace_kernel");
}

extern "C" int skyb200_masked_mean(const float* x, const float* mask, float* out, void* workspace,
                                   size_t workspace_bytes, int64_t n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(x && out && workspace, "null pointer");
    const int nblk = grid_for(n, 256, 4);
    if (workspace_bytes < (size_t)nblk * 8) return fail(SKYB200_EWORKSPACE, "masked_mean workspace too small");
    float* partials = static_cast<float*>(workspace);
    masked_sum_kernel<<<nblk, 256, 0, stream>>>(x, mask, partials, n);
    int rc = check_launch("masked_sum_kernel");
    if (rc) return rc;
    masked_mean_final_kernel<<<1, 256, 0, stream>>>(partials, nblk, out);
    return check_launch("masked_mean_final_kernel");
}

extern "C" int skyb200_approx_kl(const float* log_probs, const float* base_log_probs, const float* loss_mask,
                                 float* out, int64_t n, int32_t kl_type, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n == 0) return SKYB200_OK;
    SKYB200_REQUIRE(log_probs && base_log_probs && out, "null pointer");
    SKYB200_REQUIRE(kl_type >= 0 && kl_type <= SKYB200_KL_K3, "Invalid KL estimator type");
    approx_kl_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(log_probs, base_log_probs, loss_mask, out, n, kl_type);
    return check_launch("approx_kl_kernel");
}

extern "C" int skyb200_critic_loss(const float* values, const float* old_values, const float* returns,
                                   const float* loss_mask, float* d_values, float* stats, void* workspace,
                                   size_t workspace_bytes, int64_t B, int64_t A, float value_clip, int32_t use_clip,
                                   void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(values && returns && stats && workspace, "null pointer");
    SKYB200_REQUIRE(!use_clip || old_values != nullptr, "value clipping needs old_values");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < 65536, "need 0 < B < 65536 and A > 0");
    if (workspace_bytes < skyb200_policy_loss_workspace(B, A))
        return fail(SKYB200_EWORKSPACE, "critic_loss workspace too small");
    LossWs ws;
    char* wbase = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    loss_ws_layout(B, A, &ws, wbase);
    SKYB200_CUDA_CHECK(cudaMemsetAsync(ws.counters, 0, 16, stream));
    loss_rowstats_kernel<<<(unsigned)B, kLossThreads, 0, stream>>>(nullptr, nullptr, loss_mask, ws.row_msum, ws.row_lw,
                                                                   ws.total_msum, ws.counters, (int)B, A, 0);
    int rc = check_launch("loss_rowstats_kernel");
    if (rc) return rc;
    CriticArgs a;
    a.values = values;
    a.old_values = old_values;
    a.returns = returns;
    a.mask = loss_mask;
    a.d_values = d_values;
    a.stats = stats;
    a.ws = ws;
    a.B = (int)B;
    a.A = A;
    a.value_clip = value_clip;
    a.use_clip = use_clip;
    const unsigned chunks = (unsigned)((A + kLossTile - 1) / kLossTile);
    critic_main_kernel<<<dim3(chunks, (unsigned)B), kLossThreads, 0, stream>>>(a);
    return check_launch("critic_main_kernel");
}
