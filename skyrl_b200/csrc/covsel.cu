// Covariance-based token selection for the clip_cov / kl_cov policy losses (detached; no gradient flows through it).
//
// Reference arithmetic (skyrl-train/skyrl_train/utils/ppo_utils.py):
//   compute_policy_loss_clip_cov :765-829  cov = (adv - masked_mean(adv)) * (lp - masked_mean(lp)); candidates are valid
//                                          tokens not clipped by the PPO rule with lb < cov < ub; a RANDOM subset of
//                                          min(max(int(clip_ratio * sum(mask)), 1), #candidates) of them is dropped.
//   compute_policy_loss_kl_cov   :833-888  cov over the valid tokens (plain means); the max(1, int(n_valid * frac))
//                                          tokens with the LARGEST cov get an extra  coef * |lp - olp|  term.
//
// Both are "select exactly k tokens by key": key = cov (kl_cov) or a counter-based random number (clip_cov).  Keys
// are order-preserving uint32 (0 = not a candidate).  An 8-bit-per-pass radix select finds the k-th largest key T;
// tokens with key > T are taken, ties on T are taken in index order until k is reached - deterministic for a seed.
// Nothing syncs with the host: k itself is computed on the device from the mask statistics.
#include <algorithm>

#include "common.cuh"

namespace skyb200 {

namespace {

constexpr int kCovThreads = 256;
constexpr int kCovTile = 4096;  // tokens per block in the index-ordered tie pass
constexpr int kMomentFields = 6;

struct CovState {
    double sum_m, sum_am, sum_lm, sum_a_valid, sum_l_valid;
    unsigned long long n_valid;
    float mean_a, mean_l;
    unsigned int prefix;  // radix-select prefix; after 4 passes the threshold key T
    unsigned int done;    // nothing to select
    unsigned long long k_rem;
    unsigned long long n_keys;
};

struct CovWs {
    CovState* st;
    double* partials;     // [blocks, 6]
    unsigned int* hist;   // [4, 256]
    unsigned int* tiecnt; // [tiles]
    unsigned int* keys;   // [N]
    size_t total;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int moment_blocks(int64_t N) { return (int)std::max<int64_t>(1, std::min<int64_t>((N + 2047) / 2048, 1184)); }

CovWs carve(void* base, int64_t N) {
    CovWs w;
    size_t off = 0;
    char* b = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* p = b ? b + off : nullptr;
        off += align256(bytes);
        return p;
    };
    const int64_t tiles = (N + kCovTile - 1) / kCovTile;
    w.st = reinterpret_cast<CovState*>(take(sizeof(CovState)));
    w.hist = reinterpret_cast<unsigned int*>(take(4 * 256 * sizeof(unsigned int)));
    w.partials = reinterpret_cast<double*>(take((size_t)moment_blocks(N) * kMomentFields * sizeof(double)));
    w.tiecnt = reinterpret_cast<unsigned int*>(take((size_t)std::max<int64_t>(tiles, 1) * sizeof(unsigned int)));
    w.keys = reinterpret_cast<unsigned int*>(take((size_t)std::max<int64_t>(N, 1) * sizeof(unsigned int)));
    w.total = off + 256;
    return w;
}

__device__ __forceinline__ unsigned int ordered_key(float x) {
    if (x == 0.f) x = 0.f;  // -0 and +0 compare equal in the reference's topk: give them one key
    const unsigned int u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone in x; >= 1 for every non-NaN float
}

__device__ __forceinline__ unsigned int random_key(unsigned long long seed, unsigned long long i) {
    unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;  // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (unsigned int)(z >> 32) | 1u;
}

__global__ void __launch_bounds__(kCovThreads)
cov_moments_kernel(const float* __restrict__ lp, const float* __restrict__ adv, const float* __restrict__ mask,
                   double* __restrict__ partials, int64_t N) {
    __shared__ double sh[34];
    double s_m = 0, s_am = 0, s_lm = 0, s_a = 0, s_l = 0, n = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f, a = adv[i], l = lp[i];
        s_m += m;
        s_am += (double)(a * m);  // the reference multiplies in fp32, then sums
        s_lm += (double)(l * m);
        if (m > 0.f) {
            s_a += a;
            s_l += l;
            n += 1.0;
        }
    }
    double v[kMomentFields] = {s_m, s_am, s_lm, s_a, s_l, n};
#pragma unroll
    for (int f = 0; f < kMomentFields; ++f) {
        const double t = block_sum(v[f], sh);
        if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * kMomentFields + f] = t;
    }
}

// one thread folds the block partials in a fixed order (bit-reproducible means) and resets the select state
__global__ void cov_fold_kernel(const double* __restrict__ partials, int nblk, CovState* st, unsigned int* hist,
                                int mode) {
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) hist[i] = 0u;
    if (threadIdx.x != 0) return;
    double v[kMomentFields] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nblk; ++b)
        for (int f = 0; f < kMomentFields; ++f) v[f] += partials[(int64_t)b * kMomentFields + f];
    st->sum_m = v[0];
    st->sum_am = v[1];
    st->sum_lm = v[2];
    st->sum_a_valid = v[3];
    st->sum_l_valid = v[4];
    st->n_valid = (unsigned long long)v[5];
    if (mode == SKYB200_COV_CLIP) {  // masked_mean: sum(x * m) / clamp(sum(m), min=1)
        const double d = fmax(v[0], 1.0);
        st->mean_a = (float)(v[1] / d);
        st->mean_l = (float)(v[2] / d);
    } else {  // plain mean over the valid tokens
        const double d = fmax(v[5], 1.0);
        st->mean_a = (float)(v[3] / d);
        st->mean_l = (float)(v[4] / d);
    }
    st->prefix = 0u;
    st->done = 0u;
    st->k_rem = 0ull;
    st->n_keys = 0ull;
}

__global__ void __launch_bounds__(kCovThreads)
cov_keys_kernel(const float* __restrict__ lp, const float* __restrict__ olp, const float* __restrict__ adv,
                const float* __restrict__ mask, const CovState* __restrict__ st, unsigned int* __restrict__ keys,
                unsigned int* __restrict__ hist0, int64_t N, skyb200_cov_params p) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const float mean_a = st->mean_a, mean_l = st->mean_l;
    const float lo = 1.f - p.eps_clip_low, hi = 1.f + p.eps_clip_high;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.f, a = adv[i], l = lp[i];
        unsigned int key = 0u;
        if (m > 0.f) {
            const float cov = (a - mean_a) * (l - mean_l);
            if (p.mode == SKYB200_COV_CLIP) {
                const float ratio = expf(l - olp[i]);
                const float l1 = -a * ratio, l2 = -a * fminf(fmaxf(ratio, lo), hi);
                if (!(l2 > l1) && cov < p.cov_ub && cov > p.cov_lb) key = random_key(p.seed, (unsigned long long)i);
            } else {
                key = ordered_key(cov);
            }
        }
        keys[i] = key;
        if (key != 0u) atomicAdd(&sh[key >> 24], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist0[threadIdx.x], sh[threadIdx.x]);
}

// histogram of digit `pass` (8 bits, from the top) over the keys that match the prefix found so far
__global__ void __launch_bounds__(kCovThreads)
cov_hist_kernel(const unsigned int* __restrict__ keys, const CovState* __restrict__ st, unsigned int* __restrict__ hist,
                int64_t N, int pass) {
    if (st->done) return;
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const int hi_shift = 32 - 8 * pass, shift = 24 - 8 * pass;
    const unsigned int want = st->prefix >> hi_shift;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned int key = keys[i];
        if (key != 0u && (key >> hi_shift) == want) atomicAdd(&sh[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// single thread: choose the digit that holds the k-th largest key; pass 0 also fixes k
__global__ void cov_pick_kernel(CovState* st, const unsigned int* __restrict__ hist, int pass, skyb200_cov_params p,
                                float* __restrict__ counts) {
    if (threadIdx.x != 0 || st->done) return;
    if (pass == 0) {
        unsigned long long n_keys = 0;
        for (int d = 0; d < 256; ++d) n_keys += hist[d];
        st->n_keys = n_keys;
        long long k;
        if (p.mode == SKYB200_COV_CLIP) {  // max(int(clip_ratio * mask.sum().item()), 1), then min with #candidates
            k = (long long)(p.frac * (double)(float)st->sum_m);
            if (k < 1) k = 1;
        } else {  // max(1, int(n_valid * frac))
            k = (long long)((double)st->n_valid * p.frac);
            if (k < 1) k = 1;
        }
        if ((unsigned long long)k > n_keys) k = (long long)n_keys;
        st->k_rem = (unsigned long long)k;
        if (counts) {
            counts[0] = (float)k;
            counts[1] = (float)n_keys;
        }
        if (k == 0) {
            st->done = 1u;
            st->prefix = 0xFFFFFFFFu;
            return;
        }
    }
    unsigned long long cum = 0;
    const unsigned long long k = st->k_rem;
    for (int d = 255; d >= 0; --d) {
        const unsigned long long h = hist[d];
        if (cum + h >= k) {
            st->prefix |= (unsigned int)d << (24 - 8 * pass);
            st->k_rem = k - cum;  // how many are still needed from inside this digit
            return;
        }
        cum += h;
    }
}

__global__ void __launch_bounds__(kCovThreads)
cov_tie_count_kernel(const unsigned int* __restrict__ keys, const CovState* __restrict__ st,
                     unsigned int* __restrict__ tiecnt, int64_t N) {
    __shared__ unsigned int total;
    if (threadIdx.x == 0) total = 0u;
    __syncthreads();
    const unsigned int T = st->prefix;
    const int64_t t0 = (int64_t)blockIdx.x * kCovTile;
    unsigned int c = 0;
    for (int j = threadIdx.x; j < kCovTile; j += kCovThreads) {
        const int64_t i = t0 + j;
        if (i < N && keys[i] == T) ++c;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) tiecnt[blockIdx.x] = total;
}

// exclusive scan of the per-tile tie counts (one block; tiles are few: N / 4096)
__global__ void cov_tie_scan_kernel(unsigned int* __restrict__ tiecnt, int64_t tiles) {
    if (threadIdx.x != 0) return;
    unsigned int run = 0;
    for (int64_t t = 0; t < tiles; ++t) {
        const unsigned int c = tiecnt[t];
        tiecnt[t] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(kCovThreads)
cov_mark_kernel(const unsigned int* __restrict__ keys, const CovState* __restrict__ st,
                const unsigned int* __restrict__ tiebase, float* __restrict__ out, int64_t N, float v_sel,
                float v_other) {
    __shared__ unsigned int warp_cnt[kCovThreads / 32];
    __shared__ unsigned int run_sh;
    const unsigned int T = st->prefix;
    const unsigned long long need = st->done ? 0ull : st->k_rem;  // ties on T still to take, in index order
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) run_sh = tiebase[blockIdx.x];
    __syncthreads();
    const int64_t t0 = (int64_t)blockIdx.x * kCovTile;
    for (int j = 0; j < kCovTile; j += kCovThreads) {
        const int64_t i = t0 + j + threadIdx.x;
        const unsigned int key = (i < N) ? keys[i] : 0u;
        const bool tie = (i < N) && !st->done && key == T;
        const unsigned int bal = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        unsigned int before = run_sh;
        for (int x = 0; x < w; ++x) before += warp_cnt[x];
        const unsigned int rank = before + __popc(bal & ((1u << lane) - 1u));
        const bool sel = (i < N) && !st->done && (key > T || (tie && (unsigned long long)rank < need));
        if (i < N) out[i] = sel ? v_sel : v_other;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tot = 0;
            for (int x = 0; x < kCovThreads / 32; ++x) tot += warp_cnt[x];
            run_sh += tot;
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace skyb200

using namespace skyb200;

extern "C" size_t skyb200_cov_select_workspace(int64_t B, int64_t A) { return carve(nullptr, B * A).total; }

extern "C" int skyb200_cov_select(const float* log_probs, const float* old_log_probs, const float* advantages,
                                  const float* loss_mask, float* out, float* counts, void* workspace,
                                  size_t workspace_bytes, int64_t B, int64_t A, const skyb200_cov_params* params,
                                  void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(params != nullptr, "null params");
    SKYB200_REQUIRE(params->mode == SKYB200_COV_KL || params->mode == SKYB200_COV_CLIP, "bad mode");
    const int64_t N = B * A;
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(B > 0 && A > 0 && N < (1ll << 40), "bad shape");
    SKYB200_REQUIRE(log_probs && advantages && out && workspace, "null pointer");
    SKYB200_REQUIRE(params->mode != SKYB200_COV_CLIP || old_log_probs != nullptr, "clip_cov needs old_log_probs");
    SKYB200_REQUIRE(params->frac >= 0.0, "fraction must be >= 0");
    if (workspace_bytes < carve(nullptr, N).total) return fail(SKYB200_EWORKSPACE, "cov_select workspace too small");
    int rc = skyb200_check_device();
    if (rc) return rc;
    char* wb = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const CovWs w = carve(wb, N);
    const int mb = moment_blocks(N);
    const int64_t tiles = (N + kCovTile - 1) / kCovTile;
    const int sweep = (int)std::max<int64_t>(1, std::min<int64_t>((N + 4 * kCovThreads - 1) / (4 * kCovThreads),
                                                                  (int64_t)num_sms() * 8));
    cov_moments_kernel<<<mb, kCovThreads, 0, stream>>>(log_probs, advantages, loss_mask, w.partials, N);
    if ((rc = check_launch("cov_moments_kernel"))) return rc;
    cov_fold_kernel<<<1, 256, 0, stream>>>(w.partials, mb, w.st, w.hist, params->mode);
    if ((rc = check_launch("cov_fold_kernel"))) return rc;
    cov_keys_kernel<<<sweep, kCovThreads, 0, stream>>>(log_probs, old_log_probs, advantages, loss_mask, w.st, w.keys,
                                                       w.hist, N, *params);
    if ((rc = check_launch("cov_keys_kernel"))) return rc;
    for (int pass = 0; pass < 4; ++pass) {
        if (pass > 0) {
            cov_hist_kernel<<<sweep, kCovThreads, 0, stream>>>(w.keys, w.st, w.hist + 256 * pass, N, pass);
            if ((rc = check_launch("cov_hist_kernel"))) return rc;
        }
        cov_pick_kernel<<<1, 32, 0, stream>>>(w.st, w.hist + 256 * pass, pass, *params, counts);
        if ((rc = check_launch("cov_pick_kernel"))) return rc;
    }
    cov_tie_count_kernel<<<(unsigned)tiles, kCovThreads, 0, stream>>>(w.keys, w.st, w.tiecnt, N);
    if ((rc = check_launch("cov_tie_count_kernel"))) return rc;
    cov_tie_scan_kernel<<<1, 32, 0, stream>>>(w.tiecnt, tiles);
    if ((rc = check_launch("cov_tie_scan_kernel"))) return rc;
    cov_mark_kernel<<<(unsigned)tiles, kCovThreads, 0, stream>>>(w.keys, w.st, w.tiecnt, out, N,
                                                                 params->value_selected, params->value_other);
    return check_launch("cov_mark_kernel");
}
