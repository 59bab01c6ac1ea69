// Advantage estimation for the policy update: GRPO / RLOO group statistics, the GAE reverse scan with
// global whitening, and the KL-in-reward penalty.  All HBM-bound; rows are [B, A] fp32, row-major.
//
// Reference arithmetic (skyrl-train/skyrl_train/):
//   compute_grpo_outcome_advantage   utils/ppo_utils.py:1132-1182
//   compute_rloo_outcome_advantage   utils/ppo_utils.py:1048-1098
//   compute_gae_advantage_return     utils/ppo_utils.py:1101-1129  (no reset on mask!)
//   masked_whiten / masked_var       utils/ppo_utils.py:148-172
//   apply_reward_kl_penalty          trainer.py:981-1035
#include "common.cuh"

namespace skyb200 {

__device__ __forceinline__ uint32_t ptx_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}

template <int kMaskType>
__device__ __forceinline__ float mask_at(const void* mask, int64_t i) {
    if constexpr (kMaskType == SKYB200_MASK_F32) return static_cast<const float*>(mask)[i];
    if constexpr (kMaskType == SKYB200_MASK_I64) return static_cast<float>(static_cast<const long long*>(mask)[i]);
    return 1.f;
}

// ------------------------------------------------------------------------------------------ GRPO / RLOO

// scores[row] = sum_t rewards[row, t]   (unmasked, as the reference)
__global__ void __launch_bounds__(256) row_sum_kernel(const float* __restrict__ x, float* __restrict__ scores,
                                                      int64_t A) {
    __shared__ float sh[34];
    pdl_launch_dependents();  // the broadcast pass may start prefetching its mask tiles while this one still runs
    const int64_t base = static_cast<int64_t>(blockIdx.x) * A;
    float s = 0.f;
    const bool vec = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    if (vec) {
        // eight independent 128-bit loads in flight per thread (a 32 KB row is one round): the pass is latency-bound
        // otherwise - one block per row leaves only ~3.5 blocks per SM for a 512-row batch
        const float4* x4 = reinterpret_cast<const float4*>(x + base);
        const int64_t n4 = A / 4;
        for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 8 * blockDim.x) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t i = i0 + static_cast<int64_t>(u) * blockDim.x;
                v[u] = (i < n4) ? ldg_stream(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
    } else {
        for (int64_t i = threadIdx.x; i < A; i += blockDim.x) s += x[base + i];
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) scores[blockIdx.x] = s;
}

constexpr int kBcastTile = 4096;  // elements per block of the broadcast pass

// Per row: group mean / unbiased std from the (tiny) group, then out[row, t] = normalised_score * mask[row, t].
template <int kMaskType>
__global__ void __launch_bounds__(256)
group_broadcast_kernel(const float* __restrict__ scores, const void* __restrict__ mask,
                       const int32_t* __restrict__ group_of_row, const int32_t* __restrict__ group_start,
                       const int32_t* __restrict__ group_rows, float* __restrict__ out, int64_t A, int mode,
                       float eps) {
    __shared__ float sh[34];
    const int row = blockIdx.y;
    const int64_t base = static_cast<int64_t>(row) * A;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * kBcastTile;
    const int64_t t1 = min(A, t0 + kBcastTile);
    const bool vec = (A % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(mask) & 15u) == 0) && kMaskType != SKYB200_MASK_NONE;
    // the mask tile is requested first: its DRAM latency overlaps the chain of dependent loads of the group statistics
    constexpr int kPer = kBcastTile / (256 * 4);
    float4 m[kPer];
    if (vec) {
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int64_t t = t0 + (static_cast<int64_t>(u) * 256 + threadIdx.x) * 4;
            m[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < t1) {
                if constexpr (kMaskType == SKYB200_MASK_F32) {
                    m[u] = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(mask) + base + t));
                } else if constexpr (kMaskType == SKYB200_MASK_I64) {
                    const longlong2* mp =
                        reinterpret_cast<const longlong2*>(static_cast<const long long*>(mask) + base + t);
                    const longlong2 a = __ldcs(mp), b = __ldcs(mp + 1);
                    m[u] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
                }
            }
        }
    }
    const int g = group_of_row[row];
    const int beg = group_start[g], end = group_start[g + 1];
    const int n = end - beg;
    pdl_wait();  // the row sums of the preceding kernel are complete and visible from here on
    float value;
    {
        float s = 0.f;
        for (int i = beg + threadIdx.x; i < end; i += blockDim.x) s += scores[group_rows[i]];
        s = block_sum(s, sh);
        const float mean = s / static_cast<float>(n);
        const float sc = scores[row];
        if (mode == 2) {  // RLOO: (score - mean) * n / (n - 1); singleton -> 0
            value = (n > 1) ? (sc - mean) * (static_cast<float>(n) / static_cast<float>(n - 1)) : 0.f;
        } else if (n == 1) {  // GRPO singleton: mean 0, std 1
            value = (mode == 0) ? sc / (1.f + eps) : sc;
        } else if (mode == 1) {
            value = sc - mean;
        } else {
            float q = 0.f;
            for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
                const float d = scores[group_rows[i]] - mean;
                q = fmaf(d, d, q);
            }
            q = block_sum(q, sh);
            const float sd = sqrtf(q / static_cast<float>(n - 1));
            value = (sc - mean) / (sd + eps);
        }
    }
    if (vec) {
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int64_t t = t0 + (static_cast<int64_t>(u) * 256 + threadIdx.x) * 4;
            if (t < t1)
                __stcs(reinterpret_cast<float4*>(out + base + t),
                       make_float4(value * m[u].x, value * m[u].y, value * m[u].z, value * m[u].w));
        }
    } else {
        for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x)
            out[base + t] = value * mask_at<kMaskType>(mask, base + t);
    }
}

// ---- GRPO / RLOO in ONE launch when every prompt group has the same size (the trainer's layout: n samples per
// prompt): a thread-block CLUSTER per group.  CTA `rank` of the cluster owns kRows rows of its group; every load of
// the pass (the rewards AND the mask of those rows) is in flight at once, the mask stays in registers, the row sums
// are exchanged through distributed shared memory (one cluster barrier), and each CTA scales its own mask rows.
// 12 B per position cross HBM exactly once - no score round trip through global memory, no second launch.
template <int kMaskType, int kPer>
__device__ __forceinline__ void load_mask_row(const void* __restrict__ mask, int64_t base, int64_t n4, int T,
                                              float4 (&m)[kPer]) {
    if constexpr (kMaskType == SKYB200_MASK_F32) {
        const float4* m4 = reinterpret_cast<const float4*>(static_cast<const float*>(mask) + base);
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int64_t i = static_cast<int64_t>(u) * T + threadIdx.x;
            m[u] = (i < n4) ? __ldcs(m4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if constexpr (kMaskType == SKYB200_MASK_I64) {
        const longlong2* m2 = reinterpret_cast<const longlong2*>(static_cast<const long long*>(mask) + base);
        longlong2 a[kPer], b[kPer];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int64_t i = static_cast<int64_t>(u) * T + threadIdx.x;
            a[u] = b[u] = make_longlong2(0, 0);
            if (i < n4) {
                a[u] = __ldcs(m2 + 2 * i);
                b[u] = __ldcs(m2 + 2 * i + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kPer; ++u) m[u] = make_float4((float)a[u].x, (float)a[u].y, (float)b[u].x, (float)b[u].y);
    } else {
#pragma unroll
        for (int u = 0; u < kPer; ++u) m[u] = make_float4(1.f, 1.f, 1.f, 1.f);
    }
}

template <int kMaskType, int kRows, int kPer>
__global__ void __launch_bounds__(512)
group_cluster_kernel(const float* __restrict__ rewards, const void* __restrict__ mask,
                     const int32_t* __restrict__ group_rows, float* __restrict__ out, float* __restrict__ scores,
                     int64_t A, int n, int mode, float eps) {
    __shared__ float sh[34];
    __shared__ float my_scores[kRows];  // read by the other CTAs of the cluster
    __shared__ float my_value[kRows];
    const int T = blockDim.x;
    uint32_t csize, rank;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int g = blockIdx.x / csize;
    const int64_t n4 = A / 4;
    int rows[kRows];
    float4 m[kRows][kPer];
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
        rows[j] = group_rows[static_cast<int64_t>(g) * n + rank * kRows + j];
        const float4* x4 = reinterpret_cast<const float4*>(rewards + static_cast<int64_t>(rows[j]) * A);
        float4 r[kPer];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int64_t i = static_cast<int64_t>(u) * T + threadIdx.x;
            r[u] = (i < n4) ? ldg_stream(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        load_mask_row<kMaskType, kPer>(mask, static_cast<int64_t>(rows[j]) * A, n4, T, m[j]);
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < kPer; ++u) s += (r[u].x + r[u].y) + (r[u].z + r[u].w);
        s = block_sum(s, sh);
        if (threadIdx.x == 0) {
            my_scores[j] = s;
            scores[rows[j]] = s;
        }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float sc = 0.f;
        if (lane < n) {
            const uint32_t addr = ptx_mapa(static_cast<uint32_t>(__cvta_generic_to_shared(&my_scores[lane % kRows])),
                                           static_cast<uint32_t>(lane / kRows));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(sc) : "r"(addr) : "memory");
        }
        const float nf = static_cast<float>(n);
        const float mean = warp_sum(sc) / nf;
        const float d = (lane < n) ? sc - mean : 0.f;
        const float q = warp_sum(d * d);
        if (lane < kRows) {
            const float mine = my_scores[lane];
            float value;
            if (mode == 2) {
                value = (n > 1) ? (mine - mean) * (nf / static_cast<float>(n - 1)) : 0.f;
            } else if (n == 1) {
                value = (mode == 0) ? mine / (1.f + eps) : mine;
            } else if (mode == 1) {
                value = mine - mean;
            } else {
                value = (mine - mean) / (sqrtf(q / static_cast<float>(n - 1)) + eps);
            }
            my_value[lane] = value;
        }
    }
    // every remote read of this CTA's scores precedes the peer's arrive; the matching wait is the last instruction
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
        const float value = my_value[j];
        float4* o4 = reinterpret_cast<float4*>(out + static_cast<int64_t>(rows[j]) * A);
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int64_t i = static_cast<int64_t>(u) * T + threadIdx.x;
            if (i < n4)
                __stcs(o4 + i, make_float4(value * m[j][u].x, value * m[j][u].y, value * m[j][u].z, value * m[j][u].w));
        }
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ GAE

constexpr int kGaeThreads = 256;
constexpr int kGaePer = 16;                       // contiguous positions per thread
constexpr int kGaeTile = kGaeThreads * kGaePer;   // 4096 positions per tile
constexpr int kGaePitch = kGaePer + 1;            // +1 float of padding per thread segment: no bank conflicts

// The recurrence  adv_t = delta_t + c * adv_{t+1}  (c = gamma*lambda) is an affine map per position; a thread's
// 16-position segment composes to  carry_out = mul * carry_in + add.  Segments combine right-to-left.
struct Affine {
    double mul, add;
};
__device__ __forceinline__ Affine compose_left_of(const Affine& left, const Affine& right) {
    // apply `right` first (it is further right in the row), then `left`
    Affine r;
    r.mul = left.mul * right.mul;
    r.add = left.mul * right.add + left.add;
    return r;
}

// Tile pipeline per row (right to left, carry between tiles):
//   load   coalesced 128-bit loads of rewards / values / mask (12 requests in flight per thread) -> shared memory
//   pass 1 each thread scans its 16 consecutive positions with zero carry-in -> one affine map per thread
//   carry  suffix-composition of the 256 maps (warp shuffles + 8 warp totals) -> carry entering every thread
//   pass 2 the same 16 positions again with the true carry-in (exactly the sequential recurrence, in fp64)
//   store  coalesced 128-bit stores of advantages and returns (values come from shared memory), row moments
// The last block to finish folds the per-row moments in a fixed order (no separate fold launch).
//
// kSplit: one block per (row, 4096-position tile) instead of one per row - a batch of few long rows (8 trajectories of
// 32768 positions) otherwise occupies 8 of 148 SMs.  Every block loads its tile and runs pass 1 independently; only
// the scalar carry travels along the row: the block of tile k waits for the carry tile k + 1 published (a flag in
// global memory), applies its own 256 composed maps to it and publishes the result BEFORE its pass 2.  Work items are
// handed out by an atomic ticket in the order "rows ascending, tiles right to left", so the block a waiter depends on
// always holds an earlier ticket and is running or done: no deadlock whatever the hardware's block order.  The carry
// arithmetic is the same sequence of fp64 operations as in the row-per-block form: results are bit-identical.
struct GaeChain {
    double* carry;        // [B * T] carry leaving tile (row, tile) to its left neighbour
    unsigned int* flag;   // [B * T] 1 = carry published
    unsigned int* ticket; // [1]
    int tiles_per_row;
};

template <int kMaskType, bool kSplit>
__global__ void __launch_bounds__(kGaeThreads, 4)
gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const void* __restrict__ mask,
                float* __restrict__ adv_raw, float* __restrict__ returns, double* __restrict__ row_stats, int64_t A,
                float gamma, float lambd, double* __restrict__ moments, unsigned int* __restrict__ counter, int B,
                GaeChain chain) {
    __shared__ float s_r[kGaeThreads * kGaePitch + 1];  // rewards -> deltas -> advantages
    __shared__ float s_v[kGaeThreads * kGaePitch + 1];  // values, plus V of the position right of the tile
    __shared__ double s_wmul[kGaeThreads / 32], s_wadd[kGaeThreads / 32];
    __shared__ double s_carry;
    __shared__ double sh_d[34];
    __shared__ bool is_last;
    pdl_launch_dependents();  // lets the whitening pass get scheduled behind the last blocks of the scan
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int row = blockIdx.x;
    int my_tile = 0;
    if constexpr (kSplit) {
        __shared__ unsigned int s_ticket;
        if (tid == 0) s_ticket = atomicAdd(chain.ticket, 1u);
        __syncthreads();
        row = static_cast<int>(s_ticket / static_cast<unsigned int>(chain.tiles_per_row));
        my_tile = chain.tiles_per_row - 1 - static_cast<int>(s_ticket % static_cast<unsigned int>(chain.tiles_per_row));
    }
    const int64_t base = static_cast<int64_t>(row) * A;
    const double c = static_cast<double>(gamma) * static_cast<double>(lambd);
    double c16 = 1.0;
#pragma unroll
    for (int j = 0; j < kGaePer; ++j) c16 *= c;
    auto pad = [](int p) { return p + p / kGaePer; };  // 17-float pitch per 16 positions: conflict-free
    const bool vec = (A % 4 == 0) &&
                     (((reinterpret_cast<uintptr_t>(rewards) | reinterpret_cast<uintptr_t>(values) |
                        reinterpret_cast<uintptr_t>(adv_raw) | reinterpret_cast<uintptr_t>(returns) |
                        reinterpret_cast<uintptr_t>(mask)) & 15u) == 0);

    double sum_m = 0.0, sum_mx = 0.0, sum_mxx = 0.0;
    if (tid == 0) s_carry = 0.0;
    const int64_t num_tiles = (A + kGaeTile - 1) / kGaeTile;
    const int64_t tile_hi = kSplit ? my_tile : num_tiles - 1, tile_lo = kSplit ? my_tile : 0;
    for (int64_t tile = tile_hi; tile >= tile_lo; --tile) {
        const int64_t t0 = tile * kGaeTile;
        __syncthreads();
        float4 mk[4];
        if (vec) {
            float4 r4[4], v4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t t = t0 + (k * kGaeThreads + tid) * 4;
                r4[k] = v4[k] = mk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < A) {
                    r4[k] = __ldcs(reinterpret_cast<const float4*>(rewards + base + t));
                    v4[k] = __ldcs(reinterpret_cast<const float4*>(values + base + t));
                    if constexpr (kMaskType == SKYB200_MASK_F32) {
                        mk[k] = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(mask) + base + t));
                    } else if constexpr (kMaskType == SKYB200_MASK_I64) {
                        const longlong2* mp =
                            reinterpret_cast<const longlong2*>(static_cast<const long long*>(mask) + base + t);
                        const longlong2 ma = __ldcs(mp), mb = __ldcs(mp + 1);
                        mk[k] = make_float4((float)ma.x, (float)ma.y, (float)mb.x, (float)mb.y);
                    } else {
                        mk[k] = make_float4(1.f, 1.f, 1.f, 1.f);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o = pad((k * kGaeThreads + tid) * 4);
                s_r[o] = r4[k].x, s_r[o + 1] = r4[k].y, s_r[o + 2] = r4[k].z, s_r[o + 3] = r4[k].w;
                s_v[o] = v4[k].x, s_v[o + 1] = v4[k].y, s_v[o + 2] = v4[k].z, s_v[o + 3] = v4[k].w;
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < kGaePer; ++k) {
                const int p = k * kGaeThreads + tid;
                const int64_t t = t0 + p;
                s_r[pad(p)] = (t < A) ? rewards[base + t] : 0.f;
                s_v[pad(p)] = (t < A) ? values[base + t] : 0.f;
            }
        }
        if (tid == 0) s_v[pad(kGaeTile)] = (t0 + kGaeTile < A) ? values[base + t0 + kGaeTile] : 0.f;
        __syncthreads();
        // pass 1: delta_t = r_t + gamma * V_{t+1} - V_t (V_A = 0; positions >= A hold zeros), zero carry-in
        const int o0 = tid * kGaePitch;
        double run = 0.0;
#pragma unroll
        for (int j = kGaePer - 1; j >= 0; --j) {
            const float vn = s_v[(j == kGaePer - 1) ? o0 + kGaePitch : o0 + j + 1];  // next thread's first / slot 4096
            const float d = s_r[o0 + j] + gamma * vn - s_v[o0 + j];
            s_r[o0 + j] = d;
            run = static_cast<double>(d) + c * run;
        }
        // suffix-compose the per-thread affine maps: inclusive suffix scan inside the warp, then across the 8 warps
        Affine incl;
        incl.mul = c16;
        incl.add = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double om = __shfl_down_sync(0xffffffffu, incl.mul, o);
            const double oa = __shfl_down_sync(0xffffffffu, incl.add, o);
            if (lane + o < 32) {
                Affine right;
                right.mul = om;
                right.add = oa;
                incl = compose_left_of(incl, right);
            }
        }
        if (lane == 0) {
            s_wmul[warp] = incl.mul;
            s_wadd[warp] = incl.add;
        }
        if constexpr (kSplit) {
            // the carry entering this tile comes from the block of the tile to the right (the rightmost tile starts at 0)
            if (tid == 0 && tile + 1 < num_tiles) {
                const int64_t nb = static_cast<int64_t>(row) * chain.tiles_per_row + tile + 1;
                unsigned int ready = 0u;
                while (true) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(ready) : "l"(chain.flag + nb));
                    if (ready != 0u) break;
                    __nanosleep(64);
                }
                s_carry = __ldcg(chain.carry + nb);
            }
        }
        __syncthreads();
        double cin = s_carry;  // carry entering this warp = warps to the right applied to the tile carry
        for (int w = kGaeThreads / 32 - 1; w > warp; --w) cin = s_wmul[w] * cin + s_wadd[w];
        const double nm = __shfl_down_sync(0xffffffffu, incl.mul, 1);
        const double na = __shfl_down_sync(0xffffffffu, incl.add, 1);
        const double my_cin = (lane == 31) ? cin : (nm * cin + na);
        __syncthreads();
        if (tid == 0) {
            s_carry = s_wmul[0] * cin + s_wadd[0];  // carry into the next (left) tile
            if constexpr (kSplit) {
                if (tile > 0) {  // publish it before pass 2: the neighbour only needs the scalar
                    const int64_t me = static_cast<int64_t>(row) * chain.tiles_per_row + tile;
                    chain.carry[me] = s_carry;
                    __threadfence();
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(chain.flag + me), "r"(1u) : "memory");
                }
            }
        }
        // pass 2: the recurrence with the true carry-in
        run = my_cin;
#pragma unroll
        for (int j = kGaePer - 1; j >= 0; --j) {
            run = static_cast<double>(s_r[o0 + j]) + c * run;
            s_r[o0 + j] = static_cast<float>(run);
        }
        __syncthreads();
        auto moments_of = [&](float a, float m) {
            const double md = static_cast<double>(m), ad = static_cast<double>(a);
            sum_m += md;
            sum_mx += md * ad;
            sum_mxx += md * ad * ad;
        };
        if (vec) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = (k * kGaeThreads + tid) * 4;
                const int64_t t = t0 + p;
                if (t < A) {
                    const int o = pad(p);
                    const float4 a4 = make_float4(s_r[o], s_r[o + 1], s_r[o + 2], s_r[o + 3]);
                    // the raw advantages are re-read by the whitening pass right behind this kernel: a plain store keeps
                    // them in L2 (the returns are final: streaming store) - 34.8 -> 32.8 us for GAE + whitening on
                    // 512 x 8192, 92.1 -> 86.0 us at 4x the batch (profiles/r02z_hbm_probe.log)
                    *reinterpret_cast<float4*>(adv_raw + base + t) = a4;
                    __stcs(reinterpret_cast<float4*>(returns + base + t),
                           make_float4(a4.x + s_v[o], a4.y + s_v[o + 1], a4.z + s_v[o + 2], a4.w + s_v[o + 3]));
                    moments_of(a4.x, mk[k].x);
                    moments_of(a4.y, mk[k].y);
                    moments_of(a4.z, mk[k].z);
                    moments_of(a4.w, mk[k].w);
                }
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < kGaePer; ++k) {
                const int p = k * kGaeThreads + tid;
                const int64_t t = t0 + p;
                if (t < A) {
                    const float a = s_r[pad(p)];
                    adv_raw[base + t] = a;
                    returns[base + t] = a + s_v[pad(p)];
                    moments_of(a, mask_at<kMaskType>(mask, base + t));
                }
            }
        }
    }
    sum_m = block_sum(sum_m, sh_d);
    sum_mx = block_sum(sum_mx, sh_d);
    sum_mxx = block_sum(sum_mxx, sh_d);
    // one partial per block: per row, or per (row, tile) in row-major order when the rows are split
    const int n_parts = kSplit ? B * chain.tiles_per_row : B;
    const int part = kSplit ? row * chain.tiles_per_row + my_tile : row;
    if (tid == 0) {
        row_stats[3 * part + 0] = sum_m;
        row_stats[3 * part + 1] = sum_mx;
        row_stats[3 * part + 2] = sum_mxx;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == static_cast<unsigned int>(n_parts) - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double n = 0.0, sx = 0.0, sxx = 0.0;  // same fixed order as fold_moments_kernel
    for (int i = tid; i < n_parts; i += kGaeThreads) {
        n += __ldcg(&row_stats[3 * i + 0]);
        sx += __ldcg(&row_stats[3 * i + 1]);
        sxx += __ldcg(&row_stats[3 * i + 2]);
    }
    n = block_sum(n, sh_d);
    sx = block_sum(sx, sh_d);
    sxx = block_sum(sxx, sh_d);
    if (tid == 0) {
        moments[0] = n;
        moments[1] = sx;
        moments[2] = sxx;
        *counter = 0u;
    }
}

// One block: fold the per-row moments in a fixed order -> moments[3] = {sum m, sum m*x, sum m*x*x} (double)
__global__ void __launch_bounds__(256) fold_moments_kernel(const double* __restrict__ row_stats, int B,
                                                           double* __restrict__ moments) {
    __shared__ double sh_d[34];
    double n = 0.0, sx = 0.0, sxx = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        n += row_stats[3 * i + 0];
        sx += row_stats[3 * i + 1];
        sxx += row_stats[3 * i + 2];
    }
    n = block_sum(n, sh_d);
    sx = block_sum(sx, sh_d);
    sxx = block_sum(sxx, sh_d);
    if (threadIdx.x == 0) {
        moments[0] = n;
        moments[1] = sx;
        moments[2] = sxx;
    }
}

// moments -> stats {n, mean, var (Bessel), rstd}; masked_mean clamps its denominator to >= 1
__global__ void whiten_stats_kernel(const double* __restrict__ moments, float* __restrict__ stats_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = moments[0], sx = moments[1], sxx = moments[2];
    const double den = fmax(n, 1.0);
    const double mean = sx / den;
    double var = (sxx - 2.0 * mean * sx + mean * mean * n) / den;  // masked_mean((x - mean)^2, mask)
    if (n > 1.0) var *= n / (n - 1.0);                             // Bessel (the host raises for n in {0, 1})
    stats_out[0] = static_cast<float>(n);
    stats_out[1] = static_cast<float>(mean);
    stats_out[2] = static_cast<float>(var);
    stats_out[3] = static_cast<float>(1.0 / sqrt(static_cast<double>(static_cast<float>(var)) + 1e-8));
}

// moments -> {n, mean, var, rstd} exactly as whiten_stats_kernel computes them
__device__ __forceinline__ void whiten_stats_of(const double* __restrict__ moments, float (&st)[4]) {
    const double n = moments[0], sx = moments[1], sxx = moments[2];
    const double den = fmax(n, 1.0);
    const double mean = sx / den;
    double var = (sxx - 2.0 * mean * sx + mean * mean * n) / den;
    if (n > 1.0) var *= n / (n - 1.0);
    st[0] = static_cast<float>(n);
    st[1] = static_cast<float>(mean);
    st[2] = static_cast<float>(var);
    st[3] = static_cast<float>(1.0 / sqrt(static_cast<double>(static_cast<float>(var)) + 1e-8));
}

// x = (x - mean) * rstd in place; every block derives the statistics from the 3 moments itself (no separate
// statistics launch), block 0 publishes them.  Four 128-bit loads in flight per thread.
__global__ void __launch_bounds__(256) whiten_apply_kernel(float* __restrict__ x, const double* __restrict__ moments,
                                                           float* __restrict__ stats_out, int64_t n) {
    pdl_wait();  // moments (and x itself) come from the preceding scan
    float st[4];
    whiten_stats_of(moments, st);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats_out[0] = st[0];
        stats_out[1] = st[1];
        stats_out[2] = st[2];
        stats_out[3] = st[3];
    }
    const float mean = st[1], rstd = st[3];
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
        float4* x4 = reinterpret_cast<float4*>(x);
        const int64_t n4 = n / 4;
        for (int64_t i0 = tid; i0 < n4; i0 += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < n4) v[u] = __ldcs(x4 + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < n4) {
                    v[u].x = (v[u].x - mean) * rstd;
                    v[u].y = (v[u].y - mean) * rstd;
                    v[u].z = (v[u].z - mean) * rstd;
                    v[u].w = (v[u].w - mean) * rstd;
                    __stcs(x4 + i, v[u]);
                }
            }
        }
    } else {
        for (int64_t i = tid; i < n; i += stride) x[i] = (x[i] - mean) * rstd;
    }
}

// ------------------------------------------------------------------------------------------ REINFORCE++
// returns_t = r_t + gamma * (returns_{t+1} * mask_{t+1})   (reset after EOS; ppo_utils.py:1036-1040)
// Same affine suffix scan as GAE, with a position-dependent multiplier c_t = gamma * mask_{t+1}.
template <int kMaskType>
__global__ void __launch_bounds__(kGaeThreads)
rpp_scan_kernel(const float* __restrict__ rewards, const void* __restrict__ mask, float* __restrict__ returns,
                double* __restrict__ row_stats, int64_t A, float gamma) {
    __shared__ float s_val[kGaeThreads * kGaePitch];
    __shared__ float s_mul[kGaeThreads * kGaePitch];
    __shared__ double s_wmul[kGaeThreads / 32], s_wadd[kGaeThreads / 32];
    __shared__ double s_carry;
    __shared__ double sh_d[34];
    const int row = blockIdx.x;
    const int64_t base = static_cast<int64_t>(row) * A;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double sum_m = 0.0, sum_mx = 0.0, sum_mxx = 0.0;
    if (tid == 0) s_carry = 0.0;
    const int64_t num_tiles = (A + kGaeTile - 1) / kGaeTile;
    for (int64_t tile = num_tiles - 1; tile >= 0; --tile) {
        const int64_t t0 = tile * kGaeTile;
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kGaePer; ++k) {
            const int p = k * kGaeThreads + tid;
            const int64_t t = t0 + p;
            float d = 0.f, c = 0.f;
            if (t < A) {
                d = rewards[base + t];
                c = (t + 1 < A) ? gamma * mask_at<kMaskType>(mask, base + t + 1) : 0.f;
            }
            const int o = (p / kGaePer) * kGaePitch + (p % kGaePer);
            s_val[o] = d;
            s_mul[o] = c;
        }
        __syncthreads();
        // pass 1: this thread's 16 positions with zero carry-in -> its affine map (product of multipliers, local sum)
        double run = 0.0, prod = 1.0;
#pragma unroll
        for (int j = kGaePer - 1; j >= 0; --j) {
            const double c = static_cast<double>(s_mul[tid * kGaePitch + j]);
            run = static_cast<double>(s_val[tid * kGaePitch + j]) + c * run;
            prod *= c;
        }
        Affine incl;
        incl.mul = prod;
        incl.add = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double om = __shfl_down_sync(0xffffffffu, incl.mul, o);
            const double oa = __shfl_down_sync(0xffffffffu, incl.add, o);
            if (lane + o < 32) {
                Affine right;
                right.mul = om;
                right.add = oa;
                incl = compose_left_of(incl, right);
            }
        }
        if (lane == 0) {
            s_wmul[warp] = incl.mul;
            s_wadd[warp] = incl.add;
        }
        __syncthreads();
        double cin = s_carry;
        for (int w = kGaeThreads / 32 - 1; w > warp; --w) cin = s_wmul[w] * cin + s_wadd[w];
        const double nm = __shfl_down_sync(0xffffffffu, incl.mul, 1);
        const double na = __shfl_down_sync(0xffffffffu, incl.add, 1);
        const double my_cin = (lane == 31) ? cin : (nm * cin + na);
        __syncthreads();
        if (tid == 0) s_carry = s_wmul[0] * cin + s_wadd[0];
        // pass 2: the recurrence again with the true carry-in (no 2 x 16 fp64 register arrays)
        run = my_cin;
#pragma unroll
        for (int j = kGaePer - 1; j >= 0; --j) {
            run = static_cast<double>(s_val[tid * kGaePitch + j]) + static_cast<double>(s_mul[tid * kGaePitch + j]) * run;
            s_val[tid * kGaePitch + j] = static_cast<float>(run);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kGaePer; ++k) {
            const int p = k * kGaeThreads + tid;
            const int64_t t = t0 + p;
            if (t < A) {
                const float a = s_val[(p / kGaePer) * kGaePitch + (p % kGaePer)];
                returns[base + t] = a;
                const double m = static_cast<double>(mask_at<kMaskType>(mask, base + t));
                const double ad = static_cast<double>(a);
                sum_m += m;
                sum_mx += m * ad;
                sum_mxx += m * ad * ad;
            }
        }
    }
    sum_m = block_sum(sum_m, sh_d);
    sum_mx = block_sum(sum_mx, sh_d);
    sum_mxx = block_sum(sum_mxx, sh_d);
    if (tid == 0) {
        row_stats[3 * row + 0] = sum_m;
        row_stats[3 * row + 1] = sum_mx;
        row_stats[3 * row + 2] = sum_mxx;
    }
}

// out = (x - mean) * rstd * mask      (REINFORCE++: whitened returns re-masked, ppo_utils.py:1042-1043)
template <int kMaskType>
__global__ void __launch_bounds__(256)
whiten_masked_kernel(const float* __restrict__ x, const void* __restrict__ mask, const float* __restrict__ stats,
                     float* __restrict__ out, int64_t n) {
    const float mean = stats[1], rstd = stats[3];
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = (x[i] - mean) * rstd * mask_at<kMaskType>(mask, i);
}

// ------------------------------------------------------------------------------------------ batch normalisation
// normalize_advantages_dict (ppo_utils.py:127-145): mean over ALL positions (unmasked), variance over the mask.
// moments: {sum x, sum m, sum m*x, sum m*x*x}
template <int kMaskType>
__global__ void __launch_bounds__(256)
adv_moments_kernel(const float* __restrict__ x, const void* __restrict__ mask, double* __restrict__ partials,
                   int64_t n) {
    __shared__ double sh_d[34];
    double sx = 0.0, sm = 0.0, smx = 0.0, smxx = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double v = static_cast<double>(x[i]);
        const double m = static_cast<double>(mask_at<kMaskType>(mask, i));
        sx += v;
        sm += m;
        smx += m * v;
        smxx += m * v * v;
    }
    sx = block_sum(sx, sh_d);
    sm = block_sum(sm, sh_d);
    smx = block_sum(smx, sh_d);
    smxx = block_sum(smxx, sh_d);
    if (threadIdx.x == 0) {
        partials[4 * blockIdx.x + 0] = sx;
        partials[4 * blockIdx.x + 1] = sm;
        partials[4 * blockIdx.x + 2] = smx;
        partials[4 * blockIdx.x + 3] = smxx;
    }
}
__global__ void __launch_bounds__(256) adv_norm_stats_kernel(const double* __restrict__ partials, int nblk, int64_t n,
                                                             float* __restrict__ stats_out) {
    __shared__ double sh_d[34];
    double sx = 0.0, sm = 0.0, smx = 0.0, smxx = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
        sx += partials[4 * i + 0];
        sm += partials[4 * i + 1];
        smx += partials[4 * i + 2];
        smxx += partials[4 * i + 3];
    }
    sx = block_sum(sx, sh_d);
    sm = block_sum(sm, sh_d);
    smx = block_sum(smx, sh_d);
    smxx = block_sum(smxx, sh_d);
    if (threadIdx.x == 0) {
        const double mean = sx / static_cast<double>(n);
        const double ss = smxx - 2.0 * mean * smx + mean * mean * sm;  // sum m*(x - mean)^2
        const float var = static_cast<float>(ss) / static_cast<float>(sm);  // no clamp on the count (reference)
        const float rstd = rsqrtf(fmaxf(var, 1e-8f));
        stats_out[0] = static_cast<float>(sm);
        stats_out[1] = static_cast<float>(mean);
        stats_out[2] = var;
        stats_out[3] = rstd;
    }
}
__global__ void __launch_bounds__(256) affine_out_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                         float* __restrict__ out, int64_t n) {
    const float mean = stats[1], rstd = stats[3];
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = (x[i] - mean) * rstd;
}

// ------------------------------------------------------------------------------------------ KL in reward

__device__ __forceinline__ float kl_value(float lp, float base, int type) {
    const float d = lp - base;
    switch (type) {
        case SKYB200_KL_K1:
            return d;
        case SKYB200_KL_ABS:
            return fabsf(d);
        case SKYB200_KL_K2:
            return 0.5f * d * d;
        default: {
            const float x = fminf(fmaxf(base - lp, -20.f), 20.f);
            const float k = expf(x) - x - 1.f;
            return fminf(fmaxf(k, -10.f), 10.f);
        }
    }
}

__global__ void __launch_bounds__(256)
reward_kl_kernel(const float* __restrict__ lp, const float* __restrict__ base_lp, const float* __restrict__ mask,
                 const float* __restrict__ rewards, float* __restrict__ rewards_out, float* __restrict__ row_mean,
                 float* __restrict__ row_max, int64_t A, int kl_type, float coef) {
    __shared__ float sh[34];
    pdl_launch_dependents();
    const int row = blockIdx.x;
    const int64_t base = static_cast<int64_t>(row) * A;
    float s_km = 0.f, s_m = 0.f, mx = 0.f;
    auto one = [&](float l, float b, float m, float r) {
        float kl = kl_value(l, b, kl_type);
        if (mask) kl *= m;
        mx = fmaxf(mx, fabsf(kl));
        s_km = fmaf(kl, m, s_km);
        s_m += m;
        return r - kl * coef;
    };
    const bool vec = (A % 4 == 0) &&
                     (((reinterpret_cast<uintptr_t>(lp) | reinterpret_cast<uintptr_t>(base_lp) |
                        reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(rewards) |
                        reinterpret_cast<uintptr_t>(rewards_out)) & 15u) == 0);
    if (vec) {
        // two rounds of 128-bit loads (up to 8 requests) in flight per thread
        const int64_t n4 = A / 4;
        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 2 * blockDim.x) {
            float4 l[2], b[2], m[2], r[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t i = i0 + u * blockDim.x;
                if (i < n4) {
                    const int64_t e = base + 4 * i;
                    l[u] = __ldcs(reinterpret_cast<const float4*>(lp + e));
                    b[u] = __ldcs(reinterpret_cast<const float4*>(base_lp + e));
                    r[u] = __ldcs(reinterpret_cast<const float4*>(rewards + e));
                    m[u] = mask ? __ldcs(reinterpret_cast<const float4*>(mask + e)) : one4;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t i = i0 + u * blockDim.x;
                if (i < n4) {
                    float4 o;
                    o.x = one(l[u].x, b[u].x, m[u].x, r[u].x);
                    o.y = one(l[u].y, b[u].y, m[u].y, r[u].y);
                    o.z = one(l[u].z, b[u].z, m[u].z, r[u].z);
                    o.w = one(l[u].w, b[u].w, m[u].w, r[u].w);
                    __stcs(reinterpret_cast<float4*>(rewards_out + base + 4 * i), o);
                }
            }
        }
    } else {
        for (int64_t t = threadIdx.x; t < A; t += blockDim.x) {
            const int64_t i = base + t;
            rewards_out[i] = one(lp[i], base_lp[i], mask ? mask[i] : 1.f, rewards[i]);
        }
    }
    s_km = block_sum(s_km, sh);
    s_m = block_sum(s_m, sh);
    mx = block_max(mx, sh);
    if (threadIdx.x == 0) {
        row_mean[row] = mask ? s_km / fmaxf(s_m, 1.f) : s_km / static_cast<float>(A);
        row_max[row] = mx;
    }
}

__global__ void __launch_bounds__(256) row_means_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        int B, float* __restrict__ out) {
    __shared__ float sh[34];
    pdl_wait();  // per-row results of the preceding kernel
    float sa = 0.f, sb = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        sa += a[i];
        sb += b[i];
    }
    sa = block_sum(sa, sh);
    sb = block_sum(sb, sh);
    if (threadIdx.x == 0) {
        out[0] = sa / static_cast<float>(B);
        out[1] = sb / static_cast<float>(B);
    }
}

}  // namespace skyb200

using namespace skyb200;

extern "C" int skyb200_group_advantage(const float* rewards, const void* response_mask, int32_t mask_type,
                                       const int32_t* group_of_row, const int32_t* group_start,
                                       const int32_t* group_rows, float* out, float* scores_ws, int64_t B, int64_t A,
                                       int32_t num_groups, int32_t mode, float epsilon, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B == 0 || A == 0) return SKYB200_OK;
    SKYB200_REQUIRE(rewards && group_of_row && group_start && group_rows && out && scores_ws, "null pointer");
    SKYB200_REQUIRE(mask_type == SKYB200_MASK_NONE || response_mask != nullptr, "mask pointer missing");
    SKYB200_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (grpo), 1 (grpo no std) or 2 (rloo)");
    SKYB200_REQUIRE(B < 65536 && num_groups > 0, "need B < 65536 and num_groups > 0");
    row_sum_kernel<<<(unsigned)B, 256, 0, stream>>>(rewards, scores_ws, A);
    int rc = check_launch("row_sum_kernel");
    if (rc) return rc;
    const dim3 grid((unsigned)((A + kBcastTile - 1) / kBcastTile), (unsigned)B);
    switch (mask_type) {
        case SKYB200_MASK_F32:
            SKYB200_CUDA_CHECK(launch_pdl(group_broadcast_kernel<SKYB200_MASK_F32>, grid, dim3(256), 0, stream,
                                          scores_ws, response_mask, group_of_row, group_start, group_rows, out, A,
                                          mode, epsilon));
            break;
        case SKYB200_MASK_I64:
            SKYB200_CUDA_CHECK(launch_pdl(group_broadcast_kernel<SKYB200_MASK_I64>, grid, dim3(256), 0, stream,
                                          scores_ws, response_mask, group_of_row, group_start, group_rows, out, A,
                                          mode, epsilon));
            break;
        case SKYB200_MASK_NONE:
            SKYB200_CUDA_CHECK(launch_pdl(group_broadcast_kernel<SKYB200_MASK_NONE>, grid, dim3(256), 0, stream,
                                          scores_ws, response_mask, group_of_row, group_start, group_rows, out, A,
                                          mode, epsilon));
            break;
        default:
            return fail(SKYB200_EINVAL, "bad mask_type %d", mask_type);
    }
    return check_launch("group_broadcast_kernel");
}

// Shape gate of the one-launch cluster form (also exported: the host mirror asks before it picks the entry point).
struct GroupClusterPlan {
    int cluster, rows_per_cta, per, threads;
};
static bool group_cluster_plan(int64_t A, int64_t n, GroupClusterPlan* p) {
    if (A <= 0 || A % 4 != 0 || A > 32768 || n < 1) return false;
    int cluster, rows;
    if (n <= 8) {
        cluster = (int)n, rows = 1;
    } else if (n <= 16 && n % 2 == 0) {
        cluster = (int)(n / 2), rows = 2;
    } else {
        return false;
    }
    int per = 8, threads = 256;
    if (A > 8192) threads = 512;
    if (A > 16384) {
        if (rows != 1) return false;
        per = 16;
    }
    if (p) *p = {cluster, rows, per, threads};
    return true;
}
extern "C" int skyb200_group_advantage_uniform_supported(int64_t A, int64_t group_size) {
    return group_cluster_plan(A, group_size, nullptr) ? 1 : 0;
}

template <int kMaskType, int kRows, int kPer>
static int launch_group_cluster(const GroupClusterPlan& pl, const float* rewards, const void* mask,
                                const int32_t* group_rows, float* out, float* scores, int64_t B, int64_t A, int n,
                                int mode, float eps, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B / n * pl.cluster));
    cfg.blockDim = dim3((unsigned)pl.threads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)pl.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SKYB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, group_cluster_kernel<kMaskType, kRows, kPer>, rewards, mask, group_rows,
                                          out, scores, A, n, mode, eps));
    return SKYB200_OK;
}
template <int kMaskType>
static int dispatch_group_cluster(const GroupClusterPlan& pl, const float* rewards, const void* mask,
                                  const int32_t* group_rows, float* out, float* scores, int64_t B, int64_t A, int n,
                                  int mode, float eps, cudaStream_t stream) {
    if (pl.rows_per_cta == 2)
        return launch_group_cluster<kMaskType, 2, 8>(pl, rewards, mask, group_rows, out, scores, B, A, n, mode, eps, stream);
    if (pl.per == 16)
        return launch_group_cluster<kMaskType, 1, 16>(pl, rewards, mask, group_rows, out, scores, B, A, n, mode, eps, stream);
    return launch_group_cluster<kMaskType, 1, 8>(pl, rewards, mask, group_rows, out, scores, B, A, n, mode, eps, stream);
}

extern "C" int skyb200_group_advantage_uniform(const float* rewards, const void* response_mask, int32_t mask_type,
                                               const int32_t* group_rows, float* out, float* scores_ws, int64_t B,
                                               int64_t A, int32_t group_size, int32_t mode, float epsilon,
                                               void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B == 0 || A == 0) return SKYB200_OK;
    SKYB200_REQUIRE(rewards && group_rows && out && scores_ws, "null pointer");
    SKYB200_REQUIRE(mask_type == SKYB200_MASK_NONE || response_mask != nullptr, "mask pointer missing");
    SKYB200_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (grpo), 1 (grpo no std) or 2 (rloo)");
    SKYB200_REQUIRE(group_size >= 1 && B % group_size == 0, "B must be a whole number of groups of group_size rows");
    GroupClusterPlan pl;
    SKYB200_REQUIRE(group_cluster_plan(A, group_size, &pl),
                    "shape not supported by the cluster form (ask skyb200_group_advantage_uniform_supported)");
    SKYB200_REQUIRE(((reinterpret_cast<uintptr_t>(rewards) | reinterpret_cast<uintptr_t>(out) |
                      reinterpret_cast<uintptr_t>(response_mask)) & 15u) == 0, "pointers must be 16-byte aligned");
    switch (mask_type) {
        case SKYB200_MASK_F32:
            return dispatch_group_cluster<SKYB200_MASK_F32>(pl, rewards, response_mask, group_rows, out, scores_ws, B, A,
                                                            group_size, mode, epsilon, stream);
        case SKYB200_MASK_I64:
            return dispatch_group_cluster<SKYB200_MASK_I64>(pl, rewards, response_mask, group_rows, out, scores_ws, B, A,
                                                            group_size, mode, epsilon, stream);
        case SKYB200_MASK_NONE:
            return dispatch_group_cluster<SKYB200_MASK_NONE>(pl, rewards, response_mask, group_rows, out, scores_ws, B,
                                                             A, group_size, mode, epsilon, stream);
        default:
            return fail(SKYB200_EINVAL, "bad mask_type %d", mask_type);
    }
}

// Rows are split into one block per 4096-position tile when that is what fills the chip: fewer rows than two blocks
// per SM, more than one tile, at most kGaeMaxSplit tiles per row.  (With more rows the row-per-block form is faster:
// 512 x 8192 measured 34.8 us whole-row vs 43.0 us split - a second wave of blocks and the carry hand-off buy nothing
// once every SM already holds several rows; 8 x 32768 and 64 x 8192 are the shapes that gain.)
constexpr int kGaeMaxSplit = 32;
constexpr int64_t kGaeSplitRows = 296;  // 148 SMs x 2
static int64_t gae_split_tiles(int64_t B, int64_t A) {
    const int64_t T = (A + kGaeTile - 1) / kGaeTile;
    return (T > 1 && T <= kGaeMaxSplit && B < kGaeSplitRows) ? T : 1;
}
// partial moments (3 doubles per block) + 3 moments + counters, then the carry chain of the split form
static size_t gae_parts_cap(int64_t B) { return (size_t)std::max<int64_t>(B, std::min<int64_t>(B, kGaeSplitRows) * kGaeMaxSplit); }
extern "C" size_t skyb200_gae_workspace(int64_t B) {
    const size_t parts = gae_parts_cap(B);
    return (parts * 3 + 8) * sizeof(double) + parts * (sizeof(double) + sizeof(unsigned int)) + 1024;
}

extern "C" int skyb200_gae_scan(const float* rewards, const float* values, const void* response_mask,
                                int32_t mask_type, float* adv_raw, float* returns, double* moments, void* workspace,
                                size_t workspace_bytes, int64_t B, int64_t A, float gamma, float lambd,
                                void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(rewards && values && adv_raw && returns && moments && workspace, "null pointer");
    SKYB200_REQUIRE(mask_type == SKYB200_MASK_NONE || response_mask != nullptr, "mask pointer missing");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < (1ll << 31), "bad shape");
    SKYB200_REQUIRE((reinterpret_cast<uintptr_t>(moments) & 7u) == 0, "moments must be 8-byte aligned");
    if (workspace_bytes < skyb200_gae_workspace(B)) return fail(SKYB200_EWORKSPACE, "gae workspace too small");
    // layout: [3 moments of skyb200_gae_advantage at row_stats + 3 * B][partial moments: 3 * parts doubles, starting
    // behind them][counter, ticket][carry: parts doubles][flag: parts u32]
    double* ws0 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const size_t parts_cap = gae_parts_cap(B);
    double* row_stats = ws0 + 3 * B + 4;  // (ws0 .. ws0 + 3B + 3 stays free for the callers' moments slot)
    unsigned int* counter = reinterpret_cast<unsigned int*>(row_stats + 3 * parts_cap);
    GaeChain chain;
    chain.ticket = counter + 1;
    chain.carry = reinterpret_cast<double*>(counter + 4);
    chain.flag = reinterpret_cast<unsigned int*>(chain.carry + parts_cap);
    const int64_t T = gae_split_tiles(B, A);
    chain.tiles_per_row = (int)T;
    const bool split = T > 1;
    if (split) {  // counter + ticket + the flags of this call's B * T tiles in one memset (the carries need none)
        SKYB200_CUDA_CHECK(cudaMemsetAsync(counter, 0, 16, stream));
        SKYB200_CUDA_CHECK(cudaMemsetAsync(chain.flag, 0, (size_t)(B * T) * sizeof(unsigned int), stream));
    } else {
        SKYB200_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    }
    const unsigned grid = (unsigned)(B * T);
#define SKYB200_GAE_LAUNCH(MT)                                                                                        \
    if (split)                                                                                                        \
        gae_scan_kernel<MT, true><<<grid, kGaeThreads, 0, stream>>>(rewards, values, response_mask, adv_raw, returns, \
                                                                    row_stats, A, gamma, lambd, moments, counter,    \
                                                                    (int)B, chain);                                   \
    else                                                                                                              \
        gae_scan_kernel<MT, false><<<grid, kGaeThreads, 0, stream>>>(rewards, values, response_mask, adv_raw,        \
                                                                     returns, row_stats, A, gamma, lambd, moments,   \
                                                                     counter, (int)B, chain);
    switch (mask_type) {
        case SKYB200_MASK_F32:
            SKYB200_GAE_LAUNCH(SKYB200_MASK_F32)
            break;
        case SKYB200_MASK_I64:
            SKYB200_GAE_LAUNCH(SKYB200_MASK_I64)
            break;
        case SKYB200_MASK_NONE:
            SKYB200_GAE_LAUNCH(SKYB200_MASK_NONE)
            break;
        default:
            return fail(SKYB200_EINVAL, "bad mask_type %d", mask_type);
    }
#undef SKYB200_GAE_LAUNCH
    return check_launch("gae_scan_kernel");
}

extern "C" int skyb200_whiten_from_moments(float* x, const double* moments, float* stats_out, int64_t n,
                                           void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(x && moments && stats_out, "null pointer");
    if (n == 0) {
        whiten_stats_kernel<<<1, 32, 0, stream>>>(moments, stats_out);
        return check_launch("whiten_stats_kernel");
    }
    int64_t blocks = (n / 16 + 255) / 256;  // 4 x 128-bit accesses per thread
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SKYB200_CUDA_CHECK(launch_pdl(whiten_apply_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, moments,
                                  stats_out, n));
    return check_launch("whiten_apply_kernel");
}

extern "C" int skyb200_gae_advantage(const float* rewards, const float* values, const void* response_mask,
                                     int32_t mask_type, float* advantages, float* returns, float* stats_out,
                                     void* workspace, size_t workspace_bytes, int64_t B, int64_t A, float gamma,
                                     float lambd, void* stream_) {
    SKYB200_REQUIRE(stats_out && workspace, "null pointer");
    if (workspace_bytes < skyb200_gae_workspace(B)) return fail(SKYB200_EWORKSPACE, "gae workspace too small");
    // moments live at the (8-byte aligned) tail of the workspace
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    double* moments = reinterpret_cast<double*>(base) + 3 * B;
    int rc = skyb200_gae_scan(rewards, values, response_mask, mask_type, advantages, returns, moments, workspace,
                              workspace_bytes, B, A, gamma, lambd, stream_);
    if (rc) return rc;
    return skyb200_whiten_from_moments(advantages, moments, stats_out, B * A, stream_);
}

extern "C" int skyb200_reward_kl_penalty(const float* log_probs, const float* base_log_probs, const float* loss_mask,
                                         const float* rewards, float* rewards_out, float* metrics_out, float* row_ws,
                                         int64_t B, int64_t A, int32_t kl_type, float kl_coef, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(log_probs && base_log_probs && rewards && rewards_out && metrics_out && row_ws, "null pointer");
    SKYB200_REQUIRE(kl_type >= 0 && kl_type <= SKYB200_KL_K3, "Invalid KL estimator type");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < (1ll << 31), "bad shape");
    const float coef = kl_coef > 0.f ? kl_coef : 0.f;  // max(0, kl_loss_coef)
    reward_kl_kernel<<<(unsigned)B, 256, 0, stream>>>(log_probs, base_log_probs, loss_mask, rewards, rewards_out,
                                                      row_ws, row_ws + B, A, kl_type, coef);
    int rc = check_launch("reward_kl_kernel");
    if (rc) return rc;
    SKYB200_CUDA_CHECK(launch_pdl(row_means_kernel, dim3(1), dim3(256), 0, stream, (const float*)row_ws,
                                  (const float*)(row_ws + B), (int)B, metrics_out));
    return check_launch("row_means_kernel");
}

// out[b, :] = src[idx[b], :]   (step-wise trajectories: broadcast the last step's advantages, trainer.py:803-804)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          float* __restrict__ out, int64_t A) {
    const int64_t r = idx[blockIdx.y];
    const float* s = src + r * A;
    float* o = out + static_cast<int64_t>(blockIdx.y) * A;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * 4096;
    const int64_t t1 = min(A, t0 + 4096);
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) o[t] = s[t];
}

extern "C" int skyb200_gather_rows(const float* src, const int32_t* row_index, float* out, int64_t num_src_rows,
                                   int64_t B, int64_t A, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B == 0 || A == 0) return SKYB200_OK;
    SKYB200_REQUIRE(src && row_index && out, "null pointer");
    SKYB200_REQUIRE(num_src_rows > 0 && B < 65536, "need source rows and B < 65536");
    gather_rows_kernel<<<dim3((unsigned)((A + 4095) / 4096), (unsigned)B), 256, 0, stream>>>(src, row_index, out, A);
    return check_launch("gather_rows_kernel");
}

static int grid_cap(int64_t n, int per_sm) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int skyb200_reinforce_pp_advantage(const float* rewards, const void* response_mask, int32_t mask_type,
                                              float* advantages, float* returns, float* stats_out, void* workspace,
                                              size_t workspace_bytes, int64_t B, int64_t A, float gamma,
                                              void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(rewards && response_mask && advantages && returns && stats_out && workspace, "null pointer");
    SKYB200_REQUIRE(mask_type == SKYB200_MASK_F32 || mask_type == SKYB200_MASK_I64, "bad mask_type");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < (1ll << 31), "bad shape");
    if (workspace_bytes < skyb200_gae_workspace(B)) return fail(SKYB200_EWORKSPACE, "workspace too small");
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    double* row_stats = reinterpret_cast<double*>(base);
    double* moments = row_stats + 3 * B;
    const int64_t n = B * A;
    if (mask_type == SKYB200_MASK_F32)
        rpp_scan_kernel<SKYB200_MASK_F32><<<(unsigned)B, kGaeThreads, 0, stream>>>(rewards, response_mask, returns,
                                                                                  row_stats, A, gamma);
    else
        rpp_scan_kernel<SKYB200_MASK_I64><<<(unsigned)B, kGaeThreads, 0, stream>>>(rewards, response_mask, returns,
                                                                                  row_stats, A, gamma);
    int rc = check_launch("rpp_scan_kernel");
    if (rc) return rc;
    fold_moments_kernel<<<1, 256, 0, stream>>>(row_stats, (int)B, moments);
    rc = check_launch("fold_moments_kernel");
    if (rc) return rc;
    whiten_stats_kernel<<<1, 32, 0, stream>>>(moments, stats_out);
    rc = check_launch("whiten_stats_kernel");
    if (rc) return rc;
    if (mask_type == SKYB200_MASK_F32)
        whiten_masked_kernel<SKYB200_MASK_F32><<<grid_cap(n, 8), 256, 0, stream>>>(returns, response_mask, stats_out,
                                                                                  advantages, n);
    else
        whiten_masked_kernel<SKYB200_MASK_I64><<<grid_cap(n, 8), 256, 0, stream>>>(returns, response_mask, stats_out,
                                                                                  advantages, n);
    return check_launch("whiten_masked_kernel");
}

extern "C" size_t skyb200_normalize_advantages_workspace(void) { return (size_t)(148 * 8 * 4 * sizeof(double) * 2 + 512); }

extern "C" int skyb200_normalize_advantages(const float* advantages, const void* response_mask, int32_t mask_type,
                                            float* out, float* stats_out, void* workspace, size_t workspace_bytes,
                                            int64_t n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(advantages && response_mask && out && stats_out && workspace, "null pointer");
    SKYB200_REQUIRE(mask_type == SKYB200_MASK_F32 || mask_type == SKYB200_MASK_I64, "bad mask_type");
    SKYB200_REQUIRE(n > 0, "empty input");
    if (workspace_bytes < skyb200_normalize_advantages_workspace())
        return fail(SKYB200_EWORKSPACE, "workspace too small");
    double* partials = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    int nblk = grid_cap(n, 8);
    if (nblk > 148 * 8) nblk = 148 * 8;
    if (mask_type == SKYB200_MASK_F32)
        adv_moments_kernel<SKYB200_MASK_F32><<<nblk, 256, 0, stream>>>(advantages, response_mask, partials, n);
    else
        adv_moments_kernel<SKYB200_MASK_I64><<<nblk, 256, 0, stream>>>(advantages, response_mask, partials, n);
    int rc = check_launch("adv_moments_kernel");
    if (rc) return rc;
    adv_norm_stats_kernel<<<1, 256, 0, stream>>>(partials, nblk, n, stats_out);
    rc = check_launch("adv_norm_stats_kernel");
    if (rc) return rc;
    affine_out_kernel<<<grid_cap(n, 8), 256, 0, stream>>>(advantages, stats_out, out, n);
    return check_launch("affine_out_kernel");
}
