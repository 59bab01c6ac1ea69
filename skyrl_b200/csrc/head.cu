// Fused LM head for the policy update: hidden[N,H] x weight[V,H]^T -> per-token log-prob of the label and
// entropy, forward and backward, without ever holding [N, V] logits.
//
// Forward  (per token chunk):  tcgen05 GEMM whose epilogue reduces each 128x256 logits tile (still in TMEM)
//                              to online-softmax partials, then a merge kernel over the vocab slabs.
// Backward (per token chunk):  (1) the same GEMM again, epilogue = d(logits) in bf16 into a bounded scratch
//                              [chunk, V]; (2) d_hidden = dS * W (B operand MN-major straight from the
//                              nn.Linear weight); (3) d_weight += dS^T * hidden (both operands MN-major).
//
// Replaces model_wrapper.py:309-363 + torch_utils.py:59-177 of the reference (lm_head GEMM, logits.div_(T),
// flash-attn cross-entropy, chunked entropy) — see include/skyb200.h.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm_sm100.cuh"
#include "internal.h"

namespace skyb200 {

// ------------------------------------------------------------------------------------------ host utils

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// 2D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch in elements.
static int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                          uint32_t box_inner, uint32_t box_outer) {
    auto enc = get_encode_fn();
    if (enc == nullptr) return fail(SKYB200_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return fail(SKYB200_EINVAL, "TMA base not 16B aligned");
    if ((pitch_elems * 2) % 16 != 0) return fail(SKYB200_EINVAL, "row pitch must be a multiple of 8 bf16");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SKYB200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SKYB200_OK;
}

// 2D fp32 tensor map for TMA-store / TMA-reduce epilogues (32 x 32 boxes, 128-byte swizzle)
static int make_tmap_f32_out(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems) {
    auto enc = get_encode_fn();
    if (enc == nullptr) return fail(SKYB200_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return fail(SKYB200_EINVAL, "TMA base not 16B aligned");
    if ((pitch_elems * 4) % 16 != 0) return fail(SKYB200_EINVAL, "row pitch must be a multiple of 4 fp32");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_elems * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SKYB200_ECUDA, "cuTensorMapEncodeTiled (fp32 out) failed with CUresult %d", (int)r);
    return SKYB200_OK;
}

static GemmDebug* g_dbg_host = nullptr;  // mapped pinned memory
static GemmDebug* g_dbg_dev = nullptr;

static GemmDebug* debug_record_dev() {
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        if (cudaHostAlloc(&h, sizeof(GemmDebug), cudaHostAllocMapped) == cudaSuccess) {
            memset(h, 0, sizeof(GemmDebug));
            void* d = nullptr;
            if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
                g_dbg_host = static_cast<GemmDebug*>(h);
                g_dbg_dev = static_cast<GemmDebug*>(d);
            }
        }
    });
    return g_dbg_dev;
}

// 8192 epoch counters per device for the best-effort lockstep of the deep-K GEMMs (see gemm_sm100.cuh)
static unsigned int* lockstep_counters() {
    static unsigned int* per_dev[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (per_dev[dev] == nullptr) {
        void* p = nullptr;
        if (cudaMalloc(&p, 8192 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
        per_dev[dev] = static_cast<unsigned int*>(p);
    }
    return per_dev[dev];
}

static int env_int0(const char* name, int dflt);
// split-phase lockstep: K blocks between a CTA's arrival at the epoch counter and its wait for the others
static int lock_slack() { static int v = env_int0("SKYB200_LOCK_SLACK", 0); return v < 0 ? 0 : v; }

// A: K-major -> [M, K] with pitch lda; MN-major -> [K, M] with pitch lda.  Same for B with N.
template <int kCtas, int BN, int kStages, bool kAMn, bool kBMn, class Epi>
static int launch_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int group_m,
                       const typename Epi::Params& ep, cudaStream_t stream, int tile_order = 0,
                       uint64_t hint_a = kEvictNormal, uint64_t hint_b = kEvictNormal, int sync_every = 0,
                       const GemmMaps* maps_in = nullptr) {
    using S = GemmSmem<kCtas, BN, kStages, Epi::kSmemBytes>;
    if (M <= 0 || N <= 0 || K <= 0) return SKYB200_OK;
    CUtensorMap ta, tb;
    int rc;
    if (kAMn)
        rc = make_tmap_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64);
    else
        rc = make_tmap_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, kBM);
    if (rc) return rc;
    if (kBMn)
        rc = make_tmap_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64);
    else
        rc = make_tmap_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, GemmTile<kCtas, BN>::kSubN);
    if (rc) return rc;

    auto kern = gemm_sm100_kernel<kCtas, BN, kStages, kAMn, kBMn, Epi>;
    static bool attr_set[64] = {false};  // per template instantiation and per device (the attribute is per context)
    int dev = 0;
    SKYB200_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        SKYB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kTotal));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int tile_m = kBM * kCtas;
    const int64_t tiles = (int64_t)((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN);
    const int clusters = (int)std::min<int64_t>(tiles, num_sms() / kCtas);

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCtas);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GemmDebug* dbg = debug_record_dev();
    if (group_m < 1) group_m = 1;
    unsigned int* sync_ctr = nullptr;
    if (sync_every > 0) {
        sync_ctr = lockstep_counters();
        if (sync_ctr != nullptr) SKYB200_CUDA_CHECK(cudaMemsetAsync(sync_ctr, 0, 8192 * sizeof(unsigned int), stream));
    }
    GemmMaps maps = {nullptr, nullptr, nullptr, nullptr};
    if (maps_in != nullptr) maps = *maps_in;
    SKYB200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, group_m, tile_order, hint_a, hint_b, ep, dbg,
                                          sync_ctr, sync_every > 0 ? sync_every : 1,
                                          std::min(lock_slack(), std::max(sync_every, 1) - 1), maps));
    return SKYB200_OK;
}

// ------------------------------------------------------------------------------------------ configuration

// The production configuration: CTA pairs (cta_group::2) on 256x256 tiles, 6-stage TMA ring.
constexpr int kHeadCtas = 2;
constexpr int kHeadBN = 256;
constexpr int kHeadStages = 6;
constexpr int kTileM = kBM * kHeadCtas;

// Token chunk = 37 pair-tiles: 37 * n_tiles is a multiple of the 74 resident CTA pairs, so every GEMM of the
// chunk runs in whole waves on the 148 SMs.
static int env_int(const char* name, int dflt);
// 37 pair-tiles = 9472 tokens; SKYB200_CHUNK_TILES overrides (multiples of 37 keep whole waves).  With the wide
// (256 x 512) backward tiles the d_hidden GEMM has half as many tiles per token row, so whole waves need 74 pair-tiles
// = 18944 tokens (74 * ceil(H / 512) tiles); the deeper d_weight reduction (296 K blocks per tile) also halves the
// share of its now un-overlapped epilogue and the fp32 read-modify-write traffic per token.
static int wide_mask();
static int64_t default_chunk() {
    static int64_t v = (int64_t)env_int("SKYB200_CHUNK_TILES", wide_mask() ? 74 : 37) * kTileM;
    return v;
}
#define kDefaultChunk default_chunk()

static int64_t pick_chunk(int64_t N, int64_t chunk) {
    if (chunk <= 0) chunk = kDefaultChunk;
    chunk = std::min<int64_t>(chunk, ((N + kTileM - 1) / kTileM) * kTileM);
    chunk = ((chunk + kTileM - 1) / kTileM) * kTileM;
    return std::max<int64_t>(chunk, kTileM);
}
static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (v == nullptr || *v == 0) return dflt;
    const int x = atoi(v);
    return x > 0 ? x : dflt;
}
// tile-order group sizes (tunable for experiments; defaults measured on the 7B head)
static int fwd_group_m() { static int v = env_int("SKYB200_FWD_GROUP_M", 19); return v; }
static int dh_group_m() { static int v = env_int("SKYB200_DH_GROUP_M", 8); return v; }
static uint64_t hint_of(const char* name, int dflt) {
    const int v = env_int(name, dflt);  // 1 = normal, 2 = evict_first, 3 = evict_last
    return v == 2 ? kEvictFirst : (v == 3 ? kEvictLast : kEvictNormal);
}
static uint64_t hint_fwd_a() { static uint64_t v = hint_of("SKYB200_HINT_FWD_A", 3); return v; }
static uint64_t hint_dw_b() { static uint64_t v = hint_of("SKYB200_HINT_DW_B", 3); return v; }
// d_weight GEMM: the d(logits) panels (A) and the fp32 output traffic (C) are each touched once per launch
static uint64_t hint_dw_a() { static uint64_t v = hint_of("SKYB200_HINT_DW_A", 2); return v; }
static uint64_t hint_dw_c() { static uint64_t v = hint_of("SKYB200_HINT_DW_C", 2); return v; }
static uint64_t hint_dh_a() { static uint64_t v = hint_of("SKYB200_HINT_DH_A", 1); return v; }
static uint64_t hint_dh_b() { static uint64_t v = hint_of("SKYB200_HINT_DH_B", 1); return v; }
static int env_int0(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v == nullptr || *v == 0) ? dflt : atoi(v);
}
static int dh_lockstep() { static int v = env_int0("SKYB200_DH_LOCKSTEP", 128); return v; }
static int fwd_lockstep() { static int v = env_int0("SKYB200_FWD_LOCKSTEP", 0); return v; }
static int dw_lockstep() { static int v = env_int0("SKYB200_DW_LOCKSTEP", 128); return v; }
// d_weight tile order (N-fastest inside groups of `group_n` N tiles): the B panels of a group - group_n x [rows, 256]
// bf16 of the hidden chunk - must STAY in L2 while the d(logits) panels stream past them once per group.  The whole
// 7B-head chunk (14 panels, 68 MB) does not: ncu (profiles/r02h) shows it re-read from DRAM once per round of M tiles
// (13.0 GB read per launch, tensor pipe 80 % active).  Groups are sized to <= 36 MB of B panels and balanced: two
// groups of 7 for the 7B head (A is then read twice, 5.8 GB, but the 7.9 GB of hidden re-reads are gone); together
// with evict-first on the streamed A panels and the fp32 output traffic: 8.37 -> 8.08 ms per launch sustained
// (profiles/r02i_bwd_ab.jsonl).  SKYB200_DW_GROUP_N overrides.
static int dw_group_n(int64_t rows, int64_t H, int64_t V, int bn = kHeadBN) {
    static int forced = env_int0("SKYB200_DW_GROUP_N", 0);
    if (forced > 0) return forced;
    // DRAM-traffic model per launch, minimised over the group size g (n_groups = ceil(n_tiles / g)):
    //   A (dS, rows x V bf16) is streamed once per group;
    //   B: a group whose panels fit (<= 36 MB) is read once; one that does not is re-read once per round of
    //      74 / g concurrently resident M tiles.
    // 7B head, 9472-row chunk: g = 14 -> 10.5 GB (measured 10.9 + the reduce), g = 7 -> 5.9 GB.  A GSPO chunk of one
    // 32768-token sequence on the 32B-shaped head (B = 335 MB, no group fits): g = 10 -> 47 GB vs 64 GB for g = 20.
    const int64_t n_tiles = (H + bn - 1) / bn, m_tiles = (V + kTileM - 1) / kTileM;
    const double panel = (double)std::max<int64_t>(rows, 1) * bn * 2;
    const double a_bytes = (double)rows * (double)V * 2.0;
    double best = -1.0;
    int best_g = (int)n_tiles;
    for (int64_t g = 1; g <= n_tiles; ++g) {
        const int64_t groups = (n_tiles + g - 1) / g;
        if ((n_tiles + groups - 1) / groups != g) continue;  // only balanced group sizes
        const bool fits = g * panel <= (double)(36ll << 20);
        const double b_bytes = fits ? (double)n_tiles * panel
                                    : (double)groups * (double)m_tiles * (double)(g * g) * panel / 74.0;
        const double t = a_bytes * (double)groups + b_bytes;
        if (best < 0.0 || t < best) {
            best = t;
            best_g = (int)g;
        }
    }
    return best_g;
}
// skip all-zero d(logits) rows (masked tokens) in the backward GEMMs of the fused step (0 disables)
static int skip_zero_rows() { static int v = env_int0("SKYB200_SKIP_ZERO_ROWS", 1); return v; }
// TIMING EXPERIMENTS ONLY (results are wrong with any bit set): bound what a change could gain before building it.
//   1 = skip the d(logits) transform pass, 2 = d_weight epilogue stores instead of accumulating (no fp32 re-read),
//   4 = forward without the bf16 tile store, 8 = skip the d_hidden GEMM, 16 = skip the d_weight GEMM
static int whatif() { static int v = env_int0("SKYB200_WHATIF", 0); return v; }
// cap on the resident blocks of the d(logits) transform pass (0 = one block per item)
static int transform_blocks() { static int v = env_int0("SKYB200_TRANSFORM_BLOCKS", 0); return v; }
static int transform_evict() { static int v = env_int0("SKYB200_TRANSFORM_EVICT", 0); return v; }


// ------------------------------------------------------------------------------------------ backward GEMM dispatch

// Wide tiles (gemm_sm100.cuh, BN == 512) for the two deep-K backward GEMMs: 25 % fewer operand bytes per flop.  Measured
// (profiles/r02q_*, one B200, alternating processes): d_hidden 1228 -> 1406 TFLOP/s sustained, d_weight (accumulating)
// 1199 -> 1361, the fused step of config 2 2880 -> 2680 ms with every GEMM then at the power-capped tensor peak.
// SKYB200_WIDE: bit 1 = d_hidden, bit 2 = d_weight (default both), bit 4 = even where tile quantisation says otherwise;
// 0 = the 256 x 256 tiles everywhere.
constexpr int kWideBN = 512;
constexpr int kWideStages = 4;
static int wide_mask() { static int v = env_int0("SKYB200_WIDE", 3); return v; }
static int dh_group_m_wide() { static int v = env_int("SKYB200_DH_GROUP_M_WIDE", 7); return v; }

// fraction of the machine a [m_tiles x ceil(n / bn)] tile grid keeps busy: whole waves of 74 CTA pairs x useful columns
static double tile_grid_efficiency(int64_t m_tiles, int64_t n, int bn) {
    const int64_t n_tiles = (n + bn - 1) / bn, tiles = m_tiles * n_tiles, pairs = std::max(1, num_sms() / kHeadCtas);
    if (tiles <= 0) return 0.0;
    const int64_t waves = (tiles + pairs - 1) / pairs;
    return ((double)tiles / (double)(waves * pairs)) * ((double)n / (double)(n_tiles * bn));
}
// The wide form is worth +13..15 % at equal tile-grid efficiency and still +8 % at 0.875 of it (a 37-tile chunk of the 7B
// head: 259 tiles = 3.5 waves): take it unless quantisation gives back more than that.
static bool use_wide(int bit, int64_t m_tiles, int64_t n) {
    if (!(wide_mask() & bit)) return false;
    if (wide_mask() & 4) return true;
    return tile_grid_efficiency(m_tiles, n, kWideBN) >= 0.85 * tile_grid_efficiency(m_tiles, n, kHeadBN);
}

// d_hidden[rows, H] (bf16) = S[rows, V] * W[V, H]      (A K-major, B MN-major)
static int launch_d_hidden(const void* S, int64_t ldS, const void* weight, int rows, int64_t H, int64_t V,
                           const EpiStoreBF16::Params& eh, cudaStream_t stream, const GemmMaps* maps) {
    if (use_wide(1, (rows + kTileM - 1) / kTileM, H))
        return launch_gemm<kHeadCtas, kWideBN, kWideStages, false, true, EpiStoreBF16>(
            S, ldS, weight, H, rows, (int)H, (int)V, dh_group_m_wide(), eh, stream, 0, hint_dh_a(), hint_dh_b(),
            dh_lockstep(), maps);
    return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, true, EpiStoreBF16>(
        S, ldS, weight, H, rows, (int)H, (int)V, dh_group_m(), eh, stream, 0, hint_dh_a(), hint_dh_b(), dh_lockstep(),
        maps);
}

// d_weight[V, H] (fp32) (+)= S^T[V, rows] * hidden[rows, H]   (both MN-major); Epi = EpiTmaF32 or EpiF32Scatter
template <class Epi>
static int launch_d_weight(const void* S, int64_t ldS, const void* hidden, int rows, int64_t H, int64_t V,
                           const typename Epi::Params& ew, cudaStream_t stream, const GemmMaps* maps) {
    if (use_wide(2, (V + kTileM - 1) / kTileM, H))
        return launch_gemm<kHeadCtas, kWideBN, kWideStages, true, true, Epi>(
            S, ldS, hidden, H, (int)V, (int)H, rows, dw_group_n(rows, H, V, kWideBN), ew, stream, 1, hint_dw_a(),
            hint_dw_b(), dw_lockstep(), maps);
    return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, true, true, Epi>(
        S, ldS, hidden, H, (int)V, (int)H, rows, dw_group_n(rows, H, V), ew, stream, 1, hint_dw_a(), hint_dw_b(),
        dw_lockstep(), maps);
}


static int64_t num_slabs(int64_t V) { return (V + kHeadBN - 1) / kHeadBN; }

// ------------------------------------------------------------------------------------------ merge kernel

// Combine the per-slab online-softmax partials of one token chunk.
//   lse2 = M + log2(S);  logprob = acc_label/T - ln2*lse2;  entropy = ln2*lse2 - (T_sum/S)/T
// A block owns 32 token rows; its 8 warps split the vocabulary slabs (warp w takes slabs w, w + 8, ...), every load
// is a coalesced 128-byte row segment, each thread keeps a running (max, sum, weighted sum) that is rescaled when
// the max moves, and warp 0 folds the 8 partial results in a fixed order.  ceil(rows / 32) blocks: ~300 for a
// 9472-token chunk instead of the 37 a row-per-thread layout would give.
constexpr int kMergeRows = 32;
__device__ __forceinline__ void online_merge(float& m, float& s, float& t, float pm, float ps, float pt) {
    const float nm = fmaxf(m, pm);
    const float a = exp2f(m - nm), b = exp2f(pm - nm);  // m = -inf on the first step: a = 0
    s = fmaf(s, a, ps * b);
    t = fmaf(t, a, pt * b);
    m = nm;
}
// returns true in the threads (warp 0, row < rows) that hold the merged (M, S, T) of row `r`
__device__ __forceinline__ bool merge_slab_partials(const float* __restrict__ part_m, const float* __restrict__ part_s,
                                                    const float* __restrict__ part_t, int64_t ld, int slabs, int rows,
                                                    int& r, float& M, float& S, float& T) {
    __shared__ float sm[8][kMergeRows], ss[8][kMergeRows], st[8][kMergeRows];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    r = blockIdx.x * kMergeRows + lane;
    float m = -INFINITY, sv = 0.f, tv = 0.f;
    if (r < rows) {
#pragma unroll 4
        for (int sl = w; sl < slabs; sl += 8) {
            const int64_t o = (int64_t)sl * ld + r;
            online_merge(m, sv, tv, __ldcs(part_m + o), __ldcs(part_s + o), __ldcs(part_t + o));
        }
    }
    sm[w][lane] = m;
    ss[w][lane] = sv;
    st[w][lane] = tv;
    __syncthreads();
    if (w != 0 || r >= rows) return false;
    M = sm[0][lane], S = ss[0][lane], T = st[0][lane];
#pragma unroll
    for (int x = 1; x < 8; ++x)
        if (x < slabs) online_merge(M, S, T, sm[x][lane], ss[x][lane], st[x][lane]);
    return true;
}

__global__ void __launch_bounds__(256) head_merge_kernel(const float* __restrict__ part_m,
                                                         const float* __restrict__ part_s,
                                                         const float* __restrict__ part_t, int64_t ld, int slabs,
                                                         const float* __restrict__ label_acc,
                                                         const float* __restrict__ entropy_mask, float inv_T,
                                                         float* __restrict__ logprobs, float* __restrict__ entropy,
                                                         float* __restrict__ entropy_raw,
                                                         float* __restrict__ lse2_out, int rows) {
    int r;
    float M, S, T;
    if (!merge_slab_partials(part_m, part_s, part_t, ld, slabs, rows, r, M, S, T)) return;
    const float lse2 = M + log2f(S);
    const float lse = lse2 * 0.6931471805599453f;
    lse2_out[r] = lse2;
    logprobs[r] = label_acc[r] * inv_T - lse;
    if (entropy != nullptr || entropy_raw != nullptr) {
        float h = lse - (T / S) * inv_T;
        if (entropy_raw != nullptr) entropy_raw[r] = h;
        if (entropy_mask != nullptr) h *= entropy_mask[r];
        if (entropy != nullptr) entropy[r] = h;
    }
}

// Vocab-parallel variant: fold the slab partials of ONE vocab shard into per-token (max, sum, weighted sum,
// label accumulator); the cross-rank combine happens on these 4 numbers per token.
__global__ void __launch_bounds__(256) head_merge_stats_kernel(const float* __restrict__ part_m,
                                                               const float* __restrict__ part_s,
                                                               const float* __restrict__ part_t, int64_t ld, int slabs,
                                                               const float* __restrict__ label_acc,
                                                               float* __restrict__ stats, int rows) {
    int r;
    float M, S, T;
    if (!merge_slab_partials(part_m, part_s, part_t, ld, slabs, rows, r, M, S, T)) return;
    reinterpret_cast<float4*>(stats)[r] = make_float4(M, S, T, label_acc[r]);
}

// Cross-shard combine of the vocab-parallel head: stats [tp, N, 4] (all-gathered) -> log-prob, entropy, lse2.
__global__ void __launch_bounds__(256) head_combine_stats_kernel(const float4* __restrict__ stats, int tp, int64_t N,
                                                                 float inv_T, float* __restrict__ logprobs,
                                                                 float* __restrict__ entropy,
                                                                 float* __restrict__ lse2_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float M = -INFINITY;
    for (int s = 0; s < tp; ++s) M = fmaxf(M, stats[(int64_t)s * N + r].x);
    float S = 0.f, Tw = 0.f, Z = 0.f;
    for (int s = 0; s < tp; ++s) {
        const float4 q = stats[(int64_t)s * N + r];
        const float w = exp2f(q.x - M);
        S = fmaf(q.y, w, S);
        Tw = fmaf(q.z, w, Tw);
        Z += q.w;  // only the shard that owns the label contributes
    }
    const float lse2 = M + log2f(S);
    const float lse = lse2 * 0.6931471805599453f;
    logprobs[r] = Z * inv_T - lse;
    entropy[r] = lse - (Tw / S) * inv_T;
    lse2_out[r] = lse2;
}

// In-place turn of the bf16 tile store [rows, ld] of the fused step's forward (EpiLseStoreRel: slab-relative
// log2-logits u) into d(logits):
//   l = u + (part_m[slab][row] - lse2[row])          log2 of the probability
//   p = 2^l ; d = inv_T * ( g_lp*1[j == label] - p*(g_lp + g_ent*(l*ln2 + H)) )
// One block row per token; each thread owns 16 consecutive vocab entries (one 32-byte access each way), which lie in
// one 256-wide slab.
__global__ void __launch_bounds__(256)
dlogits_transform_kernel(__nv_bfloat16* __restrict__ S, int64_t ld, int V, const float* __restrict__ part_m,
                         int64_t ld_part, const float* __restrict__ lse2, const float* __restrict__ g_lp,
                         const float* __restrict__ g_ent, const float* __restrict__ g_ent_mask,
                         const float* __restrict__ entropy_raw, const int64_t* __restrict__ labels, float inv_T,
                         int rows, int store_evict_first) {
    // items = (token row, 4096-wide column block), row-major; a block walks them with a grid stride so the launch can
    // be throttled to a few resident blocks per SM (it runs under the GEMMs of the previous chunk)
    const int cbs = (V + 4095) / 4096;
    const int64_t total = (int64_t)rows * cbs;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int row = (int)(item / cbs);
    const int col0 = ((int)(item % cbs) * 256 + threadIdx.x) * 16;
    if (col0 >= V) continue;
    const float glp = g_lp[row];
    float ge = 0.f, H = 0.f;
    if (g_ent != nullptr) {
        ge = g_ent[row];
        if (g_ent_mask != nullptr) ge *= g_ent_mask[row];
        H = entropy_raw[row];
    }
    const int lab = static_cast<int>(labels[row]);
    const float a0 = (glp + ge * H) * (-inv_T);
    const float b0 = ge * 0.6931471805599453f * (-inv_T);
    const float lab_add = glp * inv_T;
    const bool all_zero = (glp == 0.f && ge == 0.f);
    __nv_bfloat16* ptr = S + static_cast<int64_t>(row) * ld + col0;
    // reference point of this thread's slab against the row's log-sum-exp
    const float ref = all_zero ? 0.f : __ldg(part_m + (int64_t)(col0 / kHeadBN) * ld_part + row) - lse2[row];
    if (col0 + 16 <= V && (ld & 15) == 0) {
        uint32_t r[8];
        if (all_zero) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = 0u;
        } else {
            ld_global_256_stream(ptr, r);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const __nv_bfloat162 z2 = *reinterpret_cast<const __nv_bfloat162*>(&r[k]);
                float2 z = __bfloat1622float2(z2);
                const float u0 = z.x + ref, u1 = z.y + ref;
                float d0 = ex2_approx(u0) * fmaf(u0, b0, a0);
                float d1 = ex2_approx(u1) * fmaf(u1, b0, a0);
                if (col0 + 2 * k == lab) d0 += lab_add;
                if (col0 + 2 * k + 1 == lab) d1 += lab_add;
                r[k] = pack_bf16x2(d0, d1);
            }
        }
        // re-read by the two GEMMs next - but 2.9 GB per chunk never survive in L2 until then: evict-first keeps the
        // pass from displacing the operand panels of the GEMMs it runs under (SKYB200_TRANSFORM_EVICT, A/B in r02m)
        if (store_evict_first)
            st_global_256_stream(ptr, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
        else
            st_global_256(ptr, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
    } else {
        for (int k = 0; k < 16 && col0 + k < V; ++k) {
            float d = 0.f;
            if (!all_zero) {
                const float u = __bfloat162float(ptr[k]) + ref;
                d = ex2_approx(u) * fmaf(u, b0, a0);
                if (col0 + k == lab) d += lab_add;
            }
            ptr[k] = __float2bfloat16_rn(d);
        }
    }
    }
}

// ------------------------------------------------------------------------------------------ logits path

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
    return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}

// One block per row of materialised logits: online (max, sum exp, sum exp*z) + label gather.
template <typename T>
__global__ void __launch_bounds__(512) logits_row_kernel(const T* __restrict__ logits,
                                                         const int64_t* __restrict__ labels,
                                                         const float* __restrict__ entropy_mask,
                                                         float* __restrict__ logprobs, float* __restrict__ entropy,
                                                         float* __restrict__ lse_out, int64_t V) {
    __shared__ float sh[34];
    const int64_t row = blockIdx.x;
    const T* z = logits + row * V;
    float m = -INFINITY, s = 0.f, t = 0.f;
    for (int64_t j = threadIdx.x; j < V; j += blockDim.x) {
        const float x = to_f32<T>(z[j]);
        if (x > m) {
            const float c = expf(m - x);
            s *= c;
            t *= c;
            m = x;
        }
        const float e = expf(x - m);
        s += e;
        t = fmaf(e, x, t);
    }
    const float M = block_max(m, sh);
    const float w = (m == -INFINITY) ? 0.f : expf(m - M);
    const float S = block_sum(s * w, sh);
    const float Tt = block_sum(t * w, sh);
    if (threadIdx.x == 0) {
        const float lse = M + logf(S);
        if (lse_out != nullptr) lse_out[row] = lse;
        if (logprobs != nullptr) logprobs[row] = to_f32<T>(z[labels[row]]) - lse;
        if (entropy != nullptr) {
            float h = lse - Tt / S;
            if (entropy_mask != nullptr) h *= entropy_mask[row];
            entropy[row] = h;
        }
    }
}

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
    return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

// One block per row: d_j = g_lp*(onehot - p_j) - g_ent*p_j*(log p_j + H); may run in place over the logits.
template <typename T>
__global__ void __launch_bounds__(512) logits_backward_kernel(const T* __restrict__ logits,
                                                              const int64_t* __restrict__ labels,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ entropy,
                                                              const float* __restrict__ g_lp,
                                                              const float* __restrict__ g_ent, T* d_logits,
                                                              int64_t V) {
    const int64_t row = blockIdx.x;
    const T* z = logits + row * V;
    T* d = d_logits + row * V;
    const float l = lse[row];
    const float glp = g_lp ? g_lp[row] : 0.f;
    const float ge = g_ent ? g_ent[row] : 0.f;
    const float H = g_ent ? entropy[row] : 0.f;
    const int64_t lab = labels ? labels[row] : -1;
    for (int64_t j = threadIdx.x; j < V; j += blockDim.x) {
        const float lpj = to_f32<T>(z[j]) - l;
        const float p = expf(lpj);
        float g = -p * (glp + ge * (lpj + H));
        if (j == lab) g += glp;
        d[j] = from_f32<T>(g);
    }
}

// ---- zero-row skipping of the backward GEMMs (fused step and head_backward)
// A token whose d(log-prob) and d(entropy) are both zero (masked position) has an all-zero d(logits) row: its
// 256-row tile contributes nothing to d_hidden and its 64-row K block nothing to d_weight.  live_rows_kernel turns
// the per-token gradients of a chunk into compacted lists of live tiles / K blocks (layout below); the GEMMs walk the
// lists (GemmMaps), dead d_hidden tiles are zero-filled.  Results are bit-identical to computing the zeros.
//   maps[0] = #live M tiles, maps[1] = #live K blocks, maps[2 .. 2+mt) = live flag per M tile,
//   maps[2+mt .. 2+2mt) = m_map, maps[2+2mt .. 2+2mt+kt) = k_map          (mt = chunk/256, kt = chunk/64)
static size_t live_maps_ints(int64_t chunk) { return (size_t)(2 + 2 * (chunk / kTileM) + chunk / kBK + 8); }

__global__ void __launch_bounds__(256) live_rows_kernel(const float* __restrict__ d_lp, const float* __restrict__ d_ent,
                                                        const float* __restrict__ ent_mask, int rows, int* maps, int mt,
                                                        int kt) {
    extern __shared__ unsigned char live_k[];  // [kt]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int kb = warp; kb < kt; kb += 8) {
        bool live = false;
        for (int j = lane; j < kBK; j += 32) {
            const int r = kb * kBK + j;
            if (r < rows) {
                float ge = 0.f;
                if (d_ent != nullptr) ge = d_ent[r] * (ent_mask ? ent_mask[r] : 1.f);
                live |= !(d_lp[r] == 0.f && ge == 0.f);
            }
        }
        live = __any_sync(0xffffffffu, live);
        if (lane == 0) live_k[kb] = live ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int* m_live = maps + 2;
        int* m_map = maps + 2 + mt;
        int* k_map = maps + 2 + 2 * mt;
        int nk = 0, nm = 0;
        constexpr int per = kTileM / kBK;
        for (int m = 0; m < mt; ++m) {
            bool any = false;
            for (int j = 0; j < per; ++j) {
                const int kb = m * per + j;
                if (kb < kt && live_k[kb]) {
                    k_map[nk++] = kb;
                    any = true;
                }
            }
            m_live[m] = any ? 1 : 0;
            if (any) m_map[nm++] = m;
        }
        maps[0] = nm;
        maps[1] = nk;
    }
}

// zero the d_hidden rows of the tiles the compacted GEMM does not visit
__global__ void __launch_bounds__(256) zero_dead_tiles_kernel(__nv_bfloat16* __restrict__ C, int64_t ldc, int H,
                                                              int rows, const int* __restrict__ m_live) {
    const int m = blockIdx.x;
    if (m_live[m]) return;
    const int r1 = min(rows, (m + 1) * kTileM);
    const int vec = H / 8;  // H % 8 == 0 (checked by the caller)
    for (int r = m * kTileM + blockIdx.y; r < r1; r += gridDim.y) {
        uint4* dst = reinterpret_cast<uint4*>(C + (int64_t)r * ldc);
        for (int j = threadIdx.x; j < vec; j += blockDim.x) dst[j] = make_uint4(0u, 0u, 0u, 0u);
    }
}


}  // namespace skyb200

using namespace skyb200;

// ================================================================================================ C ABI

extern "C" size_t skyb200_head_forward_workspace(int64_t N, int64_t H, int64_t V, int64_t chunk) {
    (void)H;
    const int64_t c = pick_chunk(N, chunk);
    // 3 partial arrays [slabs][c] + label_acc [c]
    return (size_t)((3 * num_slabs(V) + 1) * c * sizeof(float) + 256);
}

extern "C" size_t skyb200_head_backward_workspace(int64_t N, int64_t H, int64_t V, int64_t chunk) {
    (void)H;
    const int64_t c = pick_chunk(N, chunk);
    // dS scratch (bf16) + the live tile / K-block lists of the zero-row skipping
    return (size_t)round_up(c * round_up(V, 64) * 2, 256) + live_maps_ints(c) * sizeof(int) + 1024;
}

extern "C" int skyb200_head_forward(const void* hidden, const void* weight, const int64_t* labels,
                                    const float* entropy_mask, float* logprobs, float* entropy, float* lse2,
                                    void* workspace, size_t workspace_bytes, int64_t N, int64_t H, int64_t V,
                                    float temperature, int64_t chunk, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(hidden && weight && labels && logprobs && lse2 && workspace, "null pointer");
    SKYB200_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    SKYB200_REQUIRE(temperature > 0.f, "temperature must be > 0");
    SKYB200_REQUIRE(N < (1ll << 31) && V < (1ll << 31) && H < (1ll << 31), "dimension too large");
    if (workspace_bytes < skyb200_head_forward_workspace(N, H, V, chunk))
        return fail(SKYB200_EWORKSPACE, "head_forward workspace too small");
    int rc = skyb200_check_device();
    if (rc) return rc;

    const int64_t c = pick_chunk(N, chunk);
    const int slabs = (int)num_slabs(V);
    float* part_m = static_cast<float*>(workspace);
    float* part_s = part_m + (int64_t)slabs * c;
    float* part_t = part_s + (int64_t)slabs * c;
    float* label_acc = part_t + (int64_t)slabs * c;
    const float inv_T = 1.0f / temperature;
    const __nv_bfloat16* hid = static_cast<const __nv_bfloat16*>(hidden);

    for (int64_t r0 = 0; r0 < N; r0 += c) {
        const int rows = (int)std::min<int64_t>(c, N - r0);
        EpiLsePartials::Params ep;
        ep.scale_log2 = 1.4426950408889634f * inv_T;
        ep.part_m = part_m;
        ep.part_s = part_s;
        ep.part_t = part_t;
        ep.ld = c;
        ep.labels = labels + r0;
        ep.label_acc = label_acc;
        ep.logits = nullptr;
        ep.ld_logits = 0;
        // a label outside [0, V) never matches a slab: give it a defined (zero) logit
        SKYB200_CUDA_CHECK(cudaMemsetAsync(label_acc, 0, (size_t)rows * sizeof(float), stream));
        rc = launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiLsePartials>(
            hid + r0 * H, H, weight, H, rows, (int)V, (int)H, fwd_group_m(), ep, stream);
        if (rc) return rc;
        head_merge_kernel<<<(rows + kMergeRows - 1) / kMergeRows, 256, 0, stream>>>(
            part_m, part_s, part_t, c, slabs, label_acc, entropy_mask ? entropy_mask + r0 : nullptr, inv_T,
            logprobs + r0, entropy ? entropy + r0 : nullptr, nullptr, lse2 + r0, rows);
        rc = check_launch("head_merge_kernel");
        if (rc) return rc;
    }
    return SKYB200_OK;
}

extern "C" int skyb200_head_forward_stats(const void* hidden, const void* weight, const int64_t* labels,
                                          float* stats, void* workspace, size_t workspace_bytes, int64_t N,
                                          int64_t H, int64_t V, float temperature, int64_t chunk, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(hidden && weight && labels && stats && workspace, "null pointer");
    SKYB200_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    SKYB200_REQUIRE(temperature > 0.f, "temperature must be > 0");
    SKYB200_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 15u) == 0, "stats must be 16-byte aligned");
    SKYB200_REQUIRE(N < (1ll << 31) && V < (1ll << 31) && H < (1ll << 31), "dimension too large");
    if (workspace_bytes < skyb200_head_forward_workspace(N, H, V, chunk))
        return fail(SKYB200_EWORKSPACE, "head_forward workspace too small");
    int rc = skyb200_check_device();
    if (rc) return rc;
    const int64_t c = pick_chunk(N, chunk);
    const int slabs = (int)num_slabs(V);
    float* part_m = static_cast<float*>(workspace);
    float* part_s = part_m + (int64_t)slabs * c;
    float* part_t = part_s + (int64_t)slabs * c;
    float* label_acc = part_t + (int64_t)slabs * c;
    const __nv_bfloat16* hid = static_cast<const __nv_bfloat16*>(hidden);
    for (int64_t r0 = 0; r0 < N; r0 += c) {
        const int rows = (int)std::min<int64_t>(c, N - r0);
        EpiLsePartials::Params ep;
        ep.scale_log2 = 1.4426950408889634f / temperature;
        ep.part_m = part_m;
        ep.part_s = part_s;
        ep.part_t = part_t;
        ep.ld = c;
        ep.labels = labels + r0;
        ep.label_acc = label_acc;
        ep.logits = nullptr;
        ep.ld_logits = 0;
        SKYB200_CUDA_CHECK(cudaMemsetAsync(label_acc, 0, (size_t)rows * sizeof(float), stream));
        rc = launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiLsePartials>(
            hid + r0 * H, H, weight, H, rows, (int)V, (int)H, fwd_group_m(), ep, stream);
        if (rc) return rc;
        head_merge_stats_kernel<<<(rows + kMergeRows - 1) / kMergeRows, 256, 0, stream>>>(part_m, part_s, part_t, c, slabs, label_acc,
                                                                        stats + 4 * r0, rows);
        rc = check_launch("head_merge_stats_kernel");
        if (rc) return rc;
    }
    return SKYB200_OK;
}

extern "C" int skyb200_head_combine_stats(const float* stats_all, int32_t tp, int64_t N, float temperature,
                                          float* logprobs, float* entropy, float* lse2, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(stats_all && logprobs && entropy && lse2, "null pointer");
    SKYB200_REQUIRE(tp >= 1 && temperature > 0.f, "need tp >= 1 and temperature > 0");
    SKYB200_REQUIRE((reinterpret_cast<uintptr_t>(stats_all) & 15u) == 0, "stats must be 16-byte aligned");
    int rc = skyb200_check_device();
    if (rc) return rc;
    head_combine_stats_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(stats_all), tp, N, 1.0f / temperature, logprobs, entropy, lse2);
    return check_launch("head_combine_stats_kernel");
}

extern "C" int skyb200_head_backward(const void* hidden, const void* weight, const int64_t* labels,
                                     const float* lse2, const float* entropy, const float* g_logprobs,
                                     const float* g_entropy, void* d_hidden, float* d_weight, int32_t accumulate,
                                     void* workspace, size_t workspace_bytes, int64_t N, int64_t H, int64_t V,
                                     float temperature, int64_t chunk, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(hidden && weight && labels && lse2 && g_logprobs && workspace, "null pointer");
    SKYB200_REQUIRE(d_hidden || d_weight, "nothing to compute");
    SKYB200_REQUIRE(g_entropy == nullptr || entropy != nullptr, "g_entropy needs the forward entropy");
    SKYB200_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    SKYB200_REQUIRE(temperature > 0.f, "temperature must be > 0");
    SKYB200_REQUIRE(N < (1ll << 31) && V < (1ll << 31) && H < (1ll << 31), "dimension too large");
    if (workspace_bytes < skyb200_head_backward_workspace(N, H, V, chunk))
        return fail(SKYB200_EWORKSPACE, "head_backward workspace too small");
    int rc = skyb200_check_device();
    if (rc) return rc;
    if (N == 0) {
        if (d_weight && !accumulate) SKYB200_CUDA_CHECK(cudaMemsetAsync(d_weight, 0, (size_t)V * H * 4, stream));
        return SKYB200_OK;
    }

    const int64_t c = pick_chunk(N, chunk);
    const int64_t ldS = round_up(V, 64);
    __nv_bfloat16* dS = static_cast<__nv_bfloat16*>(workspace);
    int* maps = reinterpret_cast<int*>(static_cast<char*>(workspace) + round_up(c * ldS * 2, 256));
    const int mt = (int)(c / kTileM), kt = (int)(c / kBK);
    const GemmMaps m_maps = {maps + 2 + mt, maps, nullptr, nullptr};
    const GemmMaps k_maps = {nullptr, nullptr, maps + 2 + 2 * mt, maps + 1};
    const float inv_T = 1.0f / temperature;
    const __nv_bfloat16* hid = static_cast<const __nv_bfloat16*>(hidden);
    __nv_bfloat16* dhid = static_cast<__nv_bfloat16*>(d_hidden);

    for (int64_t r0 = 0; r0 < N; r0 += c) {
        const int rows = (int)std::min<int64_t>(c, N - r0);
        // rows with zero upstream gradients (masked tokens) have all-zero d(logits): skip their tiles / K blocks in
        // all three GEMMs.  Not on a non-accumulating first chunk, whose d_weight GEMM must see every K block.
        const bool acc_w = (accumulate != 0 || r0 > 0);
        const bool skip = skip_zero_rows() != 0 && (d_weight == nullptr || acc_w);
        if (skip) {
            live_rows_kernel<<<1, 256, (size_t)kt, stream>>>(g_logprobs + r0, g_entropy ? g_entropy + r0 : nullptr,
                                                             nullptr, rows, maps, mt, kt);
            if ((rc = check_launch("live_rows_kernel"))) return rc;
        }
        // (1) recompute logits tile by tile -> d(logits) bf16
        EpiDLogits::Params ed;
        ed.scale_log2 = 1.4426950408889634f * inv_T;
        ed.inv_temperature = inv_T;
        ed.lse2 = lse2 + r0;
        ed.g_lp = g_logprobs + r0;
        ed.g_ent = g_entropy ? g_entropy + r0 : nullptr;
        ed.entropy = entropy ? entropy + r0 : nullptr;
        ed.labels = labels + r0;
        ed.dS = dS;
        ed.ld = ldS;
        rc = launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiDLogits>(
            hid + r0 * H, H, weight, H, rows, (int)V, (int)H, fwd_group_m(), ed, stream, 0, kEvictNormal, kEvictNormal, 0,
            skip ? &m_maps : nullptr);
        if (rc) return rc;
        // (2) d_hidden[rows, H] = dS[rows, V] * W[V, H]      (A K-major, B MN-major)
        if (dhid != nullptr) {
            EpiStoreBF16::Params eh;
            eh.C = dhid + r0 * H;
            eh.ldc = H;
            if (skip) {
                const dim3 zgrid((unsigned)((rows + kTileM - 1) / kTileM), 32);
                zero_dead_tiles_kernel<<<zgrid, 256, 0, stream>>>(eh.C, H, (int)H, rows, maps + 2);
                if ((rc = check_launch("zero_dead_tiles_kernel"))) return rc;
            }
            rc = launch_d_hidden(dS, ldS, weight, rows, H, V, eh, stream, skip ? &m_maps : nullptr);
            if (rc) return rc;
        }
        // (3) d_weight[V, H] (+)= dS^T[V, rows] * hidden[rows, H]   (both MN-major)
        if (d_weight != nullptr) {
            EpiTmaF32::Params ew;
            if ((rc = make_tmap_f32_out(&ew.tmap_out, d_weight, (uint64_t)H, (uint64_t)V, (uint64_t)H))) return rc;
            ew.accumulate = acc_w ? 1 : 0;
            ew.hint = hint_dw_c();
            rc = launch_d_weight<EpiTmaF32>(dS, ldS, hid + r0 * H, rows, H, V, ew, stream, skip ? &k_maps : nullptr);
            if (rc) return rc;
        }
    }
    return SKYB200_OK;
}

// ------------------------------------------------------------------------------------------ fused step
namespace {

struct StepPlan {
    int64_t chunk;   // token CAPACITY of the per-chunk buffers (multiple of 256)
    int64_t stride;  // tokens a chunk actually covers: == chunk, except GSPO (whole sequences: k * A <= chunk)
    int64_t ldS;     // pitch of the bf16 logits / d_logits store
    int slabs;
    // everything below exists twice (parity of the chunk index): the small kernels of chunk c+1 run on the
    // auxiliary stream while the backward GEMMs of chunk c run on the caller's stream
    size_t off_part[2], off_label[2], off_lse2[2], off_entraw[2], off_dlp[2], off_dent[2], off_S[2], off_maps[2];
    size_t off_loss, total;
};

// Chunks are flat token ranges; GSPO needs whole sequences per chunk (its ratio is a per-sequence mean).
StepPlan plan_step(int64_t B, int64_t A, int64_t N, int64_t H, int64_t V, int loss_type, bool need_grad) {
    StepPlan p;
    int64_t chunk = kDefaultChunk;
    int64_t stride = 0;
    if (loss_type == SKYB200_LOSS_GSPO) {
        // whole sequences per chunk: k * A tokens in buffers of round_up(k * A, 256) rows.  The padded response
        // length A is whatever the longest response of the batch was, so it need not be a multiple of the tile
        // height: the chunk START is then not tile-aligned (every consumer takes a plain row offset) and the last
        // tile of a chunk is partial.  k (up to ~1.5 default chunks) is chosen so that the d_hidden GEMM of a chunk
        // fills whole waves of the resident CTA pairs as well as possible.
        int64_t best_k = 1;
        double best_eff = -1.0;
        for (int64_t k = 1; k * A <= std::max<int64_t>(A, kDefaultChunk + kDefaultChunk / 2); ++k) {
            const int64_t mt = (k * A + kTileM - 1) / kTileM;
            const bool wide = use_wide(1, mt, H);  // the form launch_d_hidden will pick for this chunk
            const double eff = tile_grid_efficiency(mt, H, wide ? kWideBN : kHeadBN) * (double)(k * A) /
                               (double)(mt * kTileM) * (wide ? 1.12 : 1.0);
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                best_k = k;
            }
        }
        stride = std::min<int64_t>(best_k * A, N);
        chunk = round_up(stride, kTileM);
    }
    chunk = std::min<int64_t>(chunk, round_up(N, kTileM));
    if (stride == 0) stride = chunk;
    p.chunk = chunk;
    p.stride = stride;
    p.ldS = round_up(V, 64);
    p.slabs = (int)num_slabs(V);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    for (int i = 0; i < 2; ++i) {
        p.off_part[i] = take((size_t)3 * p.slabs * chunk * 4);
        p.off_label[i] = take((size_t)chunk * 4);
        p.off_lse2[i] = take((size_t)chunk * 4);
        p.off_entraw[i] = take((size_t)chunk * 4);
        p.off_dlp[i] = take((size_t)chunk * 4);
        p.off_dent[i] = take((size_t)chunk * 4);
        p.off_S[i] = take(need_grad ? (size_t)chunk * p.ldS * 2 : 0);
        p.off_maps[i] = take(live_maps_ints(chunk) * sizeof(int));
    }
    // every chunk rounds its loss-partial slots up: 4 spare slots (128 bytes) per chunk
    p.off_loss = take(loss_chunk_ws_bytes(B, N) + (size_t)(N / stride + 2) * 8 * 4 * 4);
    p.total = off + 512;
    return p;
}

// One auxiliary stream + a few events per device, created on first use.
struct StepStreams {
    cudaStream_t aux = nullptr;
    cudaStream_t copy = nullptr;  // host -> device staging of hidden chunks (policy_step_from_host)
    cudaEvent_t start = nullptr, fwd[2] = {nullptr, nullptr}, mid[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> copied;  // one per chunk, grown on demand
    // opts->timing: (begin, end) event pairs around every GEMM launch on the caller's stream, accumulated over steps
    // until skyb200_policy_step_timing() harvests them; tkind[i] = which GEMM pair i brackets
    std::vector<cudaEvent_t> tev;
    std::vector<int> tkind;
    size_t tused = 0;
    bool ok = false;
    cudaEvent_t timing_event() {  // next unused timing event (nullptr on failure)
        if (tused == tev.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            tev.push_back(e);
        }
        return tev[tused++];
    }
    bool grow_copied(size_t n) {
        while (copied.size() < n) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
            copied.push_back(e);
        }
        return true;
    }
};
StepStreams* step_streams() {
    static StepStreams per_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    StepStreams& s = per_dev[dev];
    if (!s.ok) {
        if (cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEvent_t* evs[5] = {&s.start, &s.fwd[0], &s.fwd[1], &s.mid[0], &s.mid[1]};
        for (auto* e : evs)
            if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.ok = true;
    }
    return &s;
}

}  // namespace

extern "C" size_t skyb200_policy_step_workspace(int64_t B, int64_t A, int64_t n_tokens, int64_t H, int64_t V,
                                                int32_t loss_type, int32_t need_grad) {
    return plan_step(B, A, n_tokens > 0 ? n_tokens : B * A, H, V, loss_type, need_grad != 0).total;
}

extern "C" int skyb200_policy_step_plan(int64_t B, int64_t A, int64_t n_tokens, int64_t H, int64_t V,
                                        int32_t loss_type, int32_t need_grad, int64_t* plan_out) {
    SKYB200_REQUIRE(plan_out != nullptr, "null pointer");
    SKYB200_REQUIRE(B > 0 && A > 0 && H > 0 && V > 0, "bad shape");
    const int64_t N = n_tokens > 0 ? n_tokens : B * A;
    const StepPlan p = plan_step(B, A, N, H, V, loss_type, need_grad != 0);
    plan_out[0] = p.chunk;
    plan_out[1] = p.stride;
    plan_out[2] = (N + p.stride - 1) / p.stride;
    return SKYB200_OK;
}

// hidden_host != nullptr: `hidden` is an uninitialised device staging buffer [N, H] that this call fills chunk by
// chunk from (pinned) host memory on a copy stream, ahead of the GEMMs that read it.
static int policy_step_impl(const void* hidden_host, const void* hidden, const void* weight, const int64_t* labels,
                            const float* old_log_probs, const float* advantages, const float* loss_mask,
                            const float* base_log_probs, const float* entropy_mask, float* log_probs, float* entropy,
                            void* d_hidden, float* d_weight, int32_t accumulate, float* stats, void* workspace,
                            size_t workspace_bytes, int64_t B, int64_t A, int64_t H, int64_t V,
                            const int32_t* cu_seqlens, int64_t n_tokens, float temperature,
                            const skyb200_loss_params* params, const skyb200_step_opts* opts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const float* tis = opts ? opts->tis : nullptr;
    const float* clip_mask = opts ? opts->clip_mask : nullptr;
    const bool timing = opts != nullptr && opts->timing != 0;
    SKYB200_REQUIRE(hidden && weight && labels && old_log_probs && advantages && log_probs && entropy && stats &&
                        workspace && params,
                    "null pointer");
    SKYB200_REQUIRE(cu_seqlens == nullptr || params->loss_type != SKYB200_LOSS_GSPO,
                    "gspo needs whole sequences per chunk: use the padded [B, A] layout");
    SKYB200_REQUIRE(cu_seqlens == nullptr || n_tokens > 0, "packed layout needs n_tokens");
    SKYB200_REQUIRE(B > 0 && A > 0 && B < 65536, "need 0 < B < 65536 and A > 0");
    SKYB200_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    SKYB200_REQUIRE(temperature > 0.f, "temperature must be > 0");
    SKYB200_REQUIRE(params->loss_type >= 0 && params->loss_type <= SKYB200_LOSS_IMPORTANCE_SAMPLING,
                    "bad loss_type (the covariance losses select over the whole batch: use skyb200_policy_loss)");
    SKYB200_REQUIRE(params->reduction >= 0 && params->reduction <= SKYB200_REDUCE_SUM, "bad reduction");
    SKYB200_REQUIRE(params->kl_type >= 0 && params->kl_type <= SKYB200_KL_K3, "bad kl_type");
    SKYB200_REQUIRE(params->reduction != SKYB200_REDUCE_SEQ_MEAN_TOKEN_SUM_NORM || params->max_seq_len > 0.f,
                    "max_seq_len must be provided for seq_mean_token_sum_norm");
    const int64_t N = cu_seqlens ? n_tokens : B * A;
    SKYB200_REQUIRE(N < (1ll << 31) && V < (1ll << 31) && H < (1ll << 31), "dimension too large");
    const skyb200_dw_scatter* scatter = opts ? opts->dw_scatter : nullptr;
    if (scatter != nullptr) {
        SKYB200_REQUIRE(scatter->world >= 2 && scatter->world <= EpiF32Scatter::kMaxWorld && scatter->rank >= 0 &&
                            scatter->rank < scatter->world,
                        "dw_scatter: world must be 2..8 and 0 <= rank < world");
        SKYB200_REQUIRE(V % (32 * scatter->world) == 0, "dw_scatter: V must be a multiple of 32 * world");
        for (int r = 0; r < scatter->world; ++r)
            SKYB200_REQUIRE(scatter->shard[r] != nullptr && (reinterpret_cast<uintptr_t>(scatter->shard[r]) & 15u) == 0,
                            "dw_scatter: every shard pointer must be set and 16-byte aligned");
    }
    const bool need_grad = (d_hidden != nullptr) || (d_weight != nullptr) || (scatter != nullptr);
    const StepPlan plan = plan_step(B, A, N, H, V, params->loss_type, need_grad);
    if (workspace_bytes < plan.total) return fail(SKYB200_EWORKSPACE, "policy_step workspace too small");
    int rc = skyb200_check_device();
    if (rc) return rc;
    StepStreams* ss = step_streams();
    if (ss == nullptr) return fail(SKYB200_ECUDA, "could not create the auxiliary stream / events");
    // SKYB200_AUX_SERIAL=1 (experiment): the small kernels run on the caller's stream instead of overlapping the GEMMs
    static const int aux_serial = env_int0("SKYB200_AUX_SERIAL", 0);
    cudaStream_t aux = aux_serial ? stream : ss->aux;

    char* wb = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const int64_t c = plan.chunk;        // buffer capacity (rows) of a chunk
    const int64_t stride = plan.stride;  // tokens per chunk (== c unless GSPO keeps whole sequences together)
    LossChunkWs lws;
    loss_chunk_ws_carve(wb + plan.off_loss, B, A, &lws);
    const float inv_T = 1.0f / temperature;
    const float scale_log2 = 1.4426950408889634f * inv_T;
    const __nv_bfloat16* hid = static_cast<const __nv_bfloat16*>(hidden);
    __nv_bfloat16* dhid = static_cast<__nv_bfloat16*>(d_hidden);
    const bool use_ent_grad = need_grad && params->use_entropy_loss != 0;
    const int64_t n_chunks = (N + stride - 1) / stride;

    auto fptr = [&](size_t off) { return reinterpret_cast<float*>(wb + off); };
    // opts->timing: bracket a GEMM launch on the caller's stream with two timing events (kind 0 = forward,
    // 1 = d_hidden, 2 = d_weight); the pair sees the kernel alone because every stream wait precedes `begin`
    auto t_begin = [&]() -> int {
        if (!timing) return SKYB200_OK;
        cudaEvent_t e = ss->timing_event();
        if (e == nullptr) return fail(SKYB200_ECUDA, "could not create a timing event");
        SKYB200_CUDA_CHECK(cudaEventRecord(e, stream));
        return SKYB200_OK;
    };
    auto t_end = [&](int kind) -> int {
        if (!timing) return SKYB200_OK;
        cudaEvent_t e = ss->timing_event();
        if (e == nullptr) return fail(SKYB200_ECUDA, "could not create a timing event");
        SKYB200_CUDA_CHECK(cudaEventRecord(e, stream));
        ss->tkind.push_back(kind);
        return SKYB200_OK;
    };

    // ---- stage A (caller's stream): logits tiles -> online-softmax partials (+ bf16 copy of the tile)
    auto issue_fwd = [&](int64_t ci) -> int {
        const int p = (int)(ci & 1);
        const int64_t t0 = ci * stride;
        const int rows = (int)std::min<int64_t>(stride, N - t0);
        float* part_m = fptr(plan.off_part[p]);
        LseStoreParams ep;  // the plain epilogue takes its base-class slice
        ep.scale_log2 = scale_log2;
        ep.part_m = part_m;
        ep.part_s = part_m + (int64_t)plan.slabs * c;
        ep.part_t = part_m + 2 * (int64_t)plan.slabs * c;
        ep.ld = c;
        ep.labels = labels + t0;
        ep.label_acc = fptr(plan.off_label[p]);
        const bool store = need_grad && !(whatif() & 4);
        ep.logits = store ? reinterpret_cast<__nv_bfloat16*>(wb + plan.off_S[p]) : nullptr;
        ep.ld_logits = plan.ldS;
        // the parity buffers were last read by the auxiliary-stream kernels of chunk ci - 2
        if (ci >= 2) SKYB200_CUDA_CHECK(cudaStreamWaitEvent(stream, ss->mid[p], 0));
        if (hidden_host != nullptr) SKYB200_CUDA_CHECK(cudaStreamWaitEvent(stream, ss->copied[ci], 0));
        SKYB200_CUDA_CHECK(cudaMemsetAsync(ep.label_acc, 0, (size_t)rows * sizeof(float), stream));
        int r;
        // TMA store of the tile: [rows, V] bf16 with pitch ldS, 64-column x 32-row boxes (one per epilogue warp)
        if (store && (r = make_tmap_bf16(&ep.tmap_out, ep.logits, (uint64_t)V, (uint64_t)rows, (uint64_t)plan.ldS, 64, 32)))
            return r;
        if ((r = t_begin())) return r;
        if (store)  // with the backward to follow: slab-relative bf16 tile store for the two backward GEMMs
            r = launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiLseStoreRel>(
                hid + t0 * H, H, weight, H, rows, (int)V, (int)H, fwd_group_m(), ep, stream, 0, hint_fwd_a(),
                kEvictNormal, fwd_lockstep());
        else
            r = launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiLsePartials>(
                hid + t0 * H, H, weight, H, rows, (int)V, (int)H, fwd_group_m(),
                static_cast<const LsePartialsParams&>(ep), stream, 0, hint_fwd_a(), kEvictNormal, fwd_lockstep());
        if (r) return r;
        if ((r = t_end(0))) return r;
        SKYB200_CUDA_CHECK(cudaEventRecord(ss->fwd[p], stream));
        return SKYB200_OK;
    };
    // ---- stage B (auxiliary stream, overlaps the GEMMs): merge -> loss of the token range -> d(logits) in place
    auto issue_mid = [&](int64_t ci, int64_t slot) -> int {
        const int p = (int)(ci & 1);
        const int64_t t0 = ci * stride;
        const int rows = (int)std::min<int64_t>(stride, N - t0);
        float* part_m = fptr(plan.off_part[p]);
        float* lse2 = fptr(plan.off_lse2[p]);
        float* ent_raw = fptr(plan.off_entraw[p]);
        float* d_lp = fptr(plan.off_dlp[p]);
        float* d_ent = fptr(plan.off_dent[p]);
        SKYB200_CUDA_CHECK(cudaStreamWaitEvent(aux, ss->fwd[p], 0));
        head_merge_kernel<<<(rows + kMergeRows - 1) / kMergeRows, 256, 0, aux>>>(
            part_m, part_m + (int64_t)plan.slabs * c, part_m + 2 * (int64_t)plan.slabs * c, c, plan.slabs,
            fptr(plan.off_label[p]), entropy_mask ? entropy_mask + t0 : nullptr, inv_T, log_probs + t0, entropy + t0,
            ent_raw, lse2, rows);
        int r = check_launch("head_merge_kernel");
        if (r) return r;
        if (params->loss_type == SKYB200_LOSS_GSPO) {
            r = loss_row_logratio(log_probs, old_log_probs, loss_mask, lws, t0 / A, rows / A, A, aux);
            if (r) return r;
        }
        r = loss_flat_range(log_probs, old_log_probs, advantages, loss_mask, base_log_probs, entropy, tis, clip_mask,
                            need_grad ? d_lp : nullptr, use_ent_grad ? d_ent : nullptr, lws, B, A, t0,
                            t0 + rows, slot, *params, cu_seqlens, aux);
        if (r) return r;
        if (need_grad) {
            const int64_t items = (int64_t)rows * ((V + 4095) / 4096);
            const int64_t cap = transform_blocks() > 0 ? transform_blocks() : items;
            if (!(whatif() & 1)) {
                dlogits_transform_kernel<<<(unsigned)std::min<int64_t>(items, cap), 256, 0, aux>>>(
                    reinterpret_cast<__nv_bfloat16*>(wb + plan.off_S[p]), plan.ldS, (int)V, part_m, c, lse2, d_lp,
                    use_ent_grad ? d_ent : nullptr, (use_ent_grad && entropy_mask) ? entropy_mask + t0 : nullptr,
                    ent_raw, labels + t0, inv_T, rows, transform_evict());
                r = check_launch("dlogits_transform_kernel");
                if (r) return r;
            }
            if (skip_zero_rows()) {
                const int mt = (int)(c / kTileM), kt = (int)(c / kBK);
                live_rows_kernel<<<1, 256, (size_t)kt, aux>>>(
                    d_lp, use_ent_grad ? d_ent : nullptr, (use_ent_grad && entropy_mask) ? entropy_mask + t0 : nullptr,
                    rows, reinterpret_cast<int*>(wb + plan.off_maps[p]), mt, kt);
                r = check_launch("live_rows_kernel");
                if (r) return r;
            }
        }
        SKYB200_CUDA_CHECK(cudaEventRecord(ss->mid[p], aux));
        return SKYB200_OK;
    };
    // ---- stage C (caller's stream): d_hidden = dS * W ; d_weight (+)= dS^T * hidden
    auto issue_bwd = [&](int64_t ci) -> int {
        const int p = (int)(ci & 1);
        const int64_t t0 = ci * stride;
        const int rows = (int)std::min<int64_t>(stride, N - t0);
        __nv_bfloat16* S = reinterpret_cast<__nv_bfloat16*>(wb + plan.off_S[p]);
        SKYB200_CUDA_CHECK(cudaStreamWaitEvent(stream, ss->mid[p], 0));
        const int mt = (int)(c / kTileM);
        int* maps = reinterpret_cast<int*>(wb + plan.off_maps[p]);
        const bool skip = skip_zero_rows() != 0;
        const GemmMaps m_maps = {maps + 2 + mt, maps, nullptr, nullptr};
        const GemmMaps k_maps = {nullptr, nullptr, maps + 2 + 2 * mt, maps + 1};
        if (dhid != nullptr && !(whatif() & 8)) {
            EpiStoreBF16::Params eh;
            eh.C = dhid + t0 * H;
            eh.ldc = H;
            if (skip) {
                const dim3 zgrid((unsigned)((rows + kTileM - 1) / kTileM), 32);
                zero_dead_tiles_kernel<<<zgrid, 256, 0, stream>>>(eh.C, H, (int)H, rows, maps + 2);
                int rz = check_launch("zero_dead_tiles_kernel");
                if (rz) return rz;
            }
            int r = t_begin();
            if (r) return r;
            r = launch_d_hidden(S, plan.ldS, weight, rows, H, V, eh, stream, skip ? &m_maps : nullptr);
            if (r) return r;
            if ((r = t_end(1))) return r;
        }
        if (scatter != nullptr && !(whatif() & 16)) {
            // d_weight GEMM + reduce-scatter: every piece goes straight to its owner's shard (see EpiF32Scatter)
            EpiF32Scatter::Params es;
            const int64_t rpr = V / scatter->world;
            int r = 0;
            for (int q = 0; q < EpiF32Scatter::kMaxWorld; ++q) {
                es.shard[q] = q < scatter->world ? scatter->shard[q] : nullptr;
                // (slots beyond world are never indexed; give them a valid descriptor anyway)
                float* base = es.shard[q] ? es.shard[q] : scatter->shard[scatter->rank];
                if ((r = make_tmap_f32_out(&es.tmap[q], base, (uint64_t)H, (uint64_t)rpr, (uint64_t)H))) return r;
            }
            es.remote_red = scatter->remote_red;
            es.ld = H;
            es.rows_per_rank = (int)rpr;
            es.rank = scatter->rank;
            es.world = scatter->world;
            es.scale = scatter->scale;
            es.hint = hint_dw_c();
            if ((r = t_begin())) return r;
            r = launch_d_weight<EpiF32Scatter>(S, plan.ldS, hid + t0 * H, rows, H, V, es, stream, skip ? &k_maps : nullptr);
            if (r) return r;
            if ((r = t_end(2))) return r;
        } else if (d_weight != nullptr && !(whatif() & 16)) {
            EpiTmaF32::Params ew;
            int r = make_tmap_f32_out(&ew.tmap_out, d_weight, (uint64_t)H, (uint64_t)V, (uint64_t)H);
            if (r) return r;
            ew.accumulate = (accumulate != 0 || t0 > 0) ? 1 : 0;
            if (whatif() & 2) ew.accumulate = 0;
            ew.hint = hint_dw_c();
            if ((r = t_begin())) return r;
            r = launch_d_weight<EpiTmaF32>(S, plan.ldS, hid + t0 * H, rows, H, V, ew, stream,
                                           (skip && ew.accumulate) ? &k_maps : nullptr);
            if (r) return r;
            if ((r = t_end(2))) return r;
        }
        return SKYB200_OK;
    };

    // the auxiliary stream starts after everything already queued on the caller's stream (inputs are ready)
    rc = loss_prepass(loss_mask, clip_mask, lws, B, A, cu_seqlens, stream);
    if (rc) return rc;
    SKYB200_CUDA_CHECK(cudaEventRecord(ss->start, stream));
    SKYB200_CUDA_CHECK(cudaStreamWaitEvent(aux, ss->start, 0));
    if (hidden_host != nullptr) {
        // all chunk copies are queued now: PCIe runs ahead of the tensor cores, only chunk 0's copy is exposed
        if (!ss->grow_copied((size_t)n_chunks)) return fail(SKYB200_ECUDA, "could not create copy events");
        SKYB200_CUDA_CHECK(cudaStreamWaitEvent(ss->copy, ss->start, 0));
        const char* src = static_cast<const char*>(hidden_host);
        char* dst = reinterpret_cast<char*>(const_cast<void*>(hidden));
        for (int64_t ci = 0; ci < n_chunks; ++ci) {
            const int64_t t0 = ci * stride;
            const size_t off = (size_t)t0 * (size_t)H * sizeof(__nv_bfloat16);
            const size_t bytes = (size_t)std::min<int64_t>(stride, N - t0) * (size_t)H * sizeof(__nv_bfloat16);
            SKYB200_CUDA_CHECK(cudaMemcpyAsync(dst + off, src + off, bytes, cudaMemcpyHostToDevice, ss->copy));
            SKYB200_CUDA_CHECK(cudaEventRecord(ss->copied[ci], ss->copy));
        }
    }

    std::vector<int64_t> slot_of(n_chunks);
    {
        int64_t slot = 0;
        for (int64_t ci = 0; ci < n_chunks; ++ci) {
            slot_of[ci] = slot;
            slot += loss_chunk_blocks(std::min<int64_t>(stride, N - ci * stride));
        }
    }
    rc = issue_fwd(0);
    if (rc) return rc;
    rc = issue_mid(0, slot_of[0]);
    if (rc) return rc;
    for (int64_t ci = 0; ci < n_chunks; ++ci) {
        if (ci + 1 < n_chunks) {
            rc = issue_fwd(ci + 1);
            if (rc) return rc;
            rc = issue_mid(ci + 1, slot_of[ci + 1]);
            if (rc) return rc;
        }
        if (need_grad) {
            rc = issue_bwd(ci);
            if (rc) return rc;
        }
    }
    // join: the caller's stream continues only after the last auxiliary-stream work
    SKYB200_CUDA_CHECK(cudaStreamWaitEvent(stream, ss->mid[(n_chunks - 1) & 1], 0));
    if (n_chunks > 1) SKYB200_CUDA_CHECK(cudaStreamWaitEvent(stream, ss->mid[(n_chunks - 2) & 1], 0));
    const int64_t total_slots =
        slot_of[n_chunks - 1] + loss_chunk_blocks(std::min<int64_t>(stride, N - (n_chunks - 1) * stride));
    return loss_finalize(lws, total_slots, true, clip_mask != nullptr, stats, *params, stream);
}

extern "C" int skyb200_policy_step(const void* hidden, const void* weight, const int64_t* labels,
                                   const float* old_log_probs, const float* advantages, const float* loss_mask,
                                   const float* base_log_probs, const float* entropy_mask, float* log_probs,
                                   float* entropy, void* d_hidden, float* d_weight, int32_t accumulate, float* stats,
                                   void* workspace, size_t workspace_bytes, int64_t B, int64_t A, int64_t H, int64_t V,
                                   const int32_t* cu_seqlens, int64_t n_tokens, float temperature,
                                   const skyb200_loss_params* params, const skyb200_step_opts* opts, void* stream) {
    return policy_step_impl(nullptr, hidden, weight, labels, old_log_probs, advantages, loss_mask, base_log_probs,
                            entropy_mask, log_probs, entropy, d_hidden, d_weight, accumulate, stats, workspace,
                            workspace_bytes, B, A, H, V, cu_seqlens, n_tokens, temperature, params, opts, stream);
}

extern "C" int skyb200_policy_step_from_host(const void* hidden_host, void* hidden_staging, const void* weight,
                                             const int64_t* labels, const float* old_log_probs,
                                             const float* advantages, const float* loss_mask,
                                             const float* base_log_probs, const float* entropy_mask, float* log_probs,
                                             float* entropy, void* d_hidden, float* d_weight, int32_t accumulate,
                                             float* stats, void* workspace, size_t workspace_bytes, int64_t B,
                                             int64_t A, int64_t H, int64_t V, const int32_t* cu_seqlens,
                                             int64_t n_tokens, float temperature, const skyb200_loss_params* params,
                                             const skyb200_step_opts* opts, void* stream) {
    SKYB200_REQUIRE(hidden_host && hidden_staging, "null pointer");
    return policy_step_impl(hidden_host, hidden_staging, weight, labels, old_log_probs, advantages, loss_mask,
                            base_log_probs, entropy_mask, log_probs, entropy, d_hidden, d_weight, accumulate, stats,
                            workspace, workspace_bytes, B, A, H, V, cu_seqlens, n_tokens, temperature, params, opts,
                            stream);
}

extern "C" int skyb200_policy_step_timing(double* out) {
    SKYB200_REQUIRE(out != nullptr, "null pointer");
    StepStreams* ss = step_streams();
    if (ss == nullptr) return fail(SKYB200_ECUDA, "could not create the auxiliary stream / events");
    for (int i = 0; i < 6; ++i) out[i] = 0.0;
    const size_t pairs = ss->tkind.size();
    if (pairs > 0) SKYB200_CUDA_CHECK(cudaEventSynchronize(ss->tev[2 * pairs - 1]));
    for (size_t i = 0; i < pairs; ++i) {
        float ms = 0.f;
        SKYB200_CUDA_CHECK(cudaEventElapsedTime(&ms, ss->tev[2 * i], ss->tev[2 * i + 1]));
        const int k = ss->tkind[i];
        out[k] += ms;
        out[3 + k] += 1.0;
    }
    ss->tkind.clear();
    ss->tused = 0;
    return SKYB200_OK;
}

extern "C" int skyb200_logits_logprob_entropy(const void* logits, int32_t logits_dtype, const int64_t* labels,
                                              const float* entropy_mask, float* logprobs, float* entropy,
                                              float* lse, int64_t N, int64_t V, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(logits != nullptr, "null logits");
    SKYB200_REQUIRE(logprobs == nullptr || labels != nullptr, "labels required for logprobs");
    SKYB200_REQUIRE(V > 0 && N < (1ll << 31), "bad shape");
    if (logits_dtype == 0)
        logits_row_kernel<float><<<(unsigned)N, 512, 0, stream>>>(static_cast<const float*>(logits), labels,
                                                                  entropy_mask, logprobs, entropy, lse, V);
    else if (logits_dtype == 1)
        logits_row_kernel<__nv_bfloat16><<<(unsigned)N, 512, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(logits), labels, entropy_mask, logprobs, entropy, lse, V);
    else
        return fail(SKYB200_EINVAL, "logits_dtype must be 0 (fp32) or 1 (bf16)");
    return check_launch("logits_row_kernel");
}

extern "C" int skyb200_logits_backward(const void* logits, int32_t logits_dtype, const int64_t* labels,
                                       const float* lse, const float* entropy, const float* g_lp, const float* g_ent,
                                       void* d_logits, int64_t N, int64_t V, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (N == 0) return SKYB200_OK;
    SKYB200_REQUIRE(logits && lse && d_logits, "null pointer");
    SKYB200_REQUIRE(g_ent == nullptr || entropy != nullptr, "g_ent needs the forward entropy");
    SKYB200_REQUIRE(g_lp == nullptr || labels != nullptr, "g_lp needs labels");
    SKYB200_REQUIRE(V > 0 && N < (1ll << 31), "bad shape");
    if (logits_dtype == 0)
        logits_backward_kernel<float><<<(unsigned)N, 512, 0, stream>>>(
            static_cast<const float*>(logits), labels, lse, entropy, g_lp, g_ent, static_cast<float*>(d_logits), V);
    else if (logits_dtype == 1)
        logits_backward_kernel<__nv_bfloat16><<<(unsigned)N, 512, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(logits), labels, lse, entropy, g_lp, g_ent,
            static_cast<__nv_bfloat16*>(d_logits), V);
    else
        return fail(SKYB200_EINVAL, "logits_dtype must be 0 (fp32) or 1 (bf16)");
    return check_launch("logits_backward_kernel");
}

template <int kCtas, int BN, int kStages>
static int selftest_dispatch(const void* A, const void* B, float* C, int M, int N, int K, int a_mn, int b_mn,
                             int accumulate, cudaStream_t stream) {
    EpiStoreF32::Params ep;
    ep.C = C;
    ep.ldc = N;
    ep.accumulate = accumulate;
    const int64_t lda = a_mn ? M : K, ldb = b_mn ? N : K;
    if (!a_mn && !b_mn)
        return launch_gemm<kCtas, BN, kStages, false, false, EpiStoreF32>(A, lda, B, ldb, M, N, K, 4, ep, stream);
    if (!a_mn && b_mn)
        return launch_gemm<kCtas, BN, kStages, false, true, EpiStoreF32>(A, lda, B, ldb, M, N, K, 4, ep, stream);
    if (a_mn && !b_mn)
        return launch_gemm<kCtas, BN, kStages, true, false, EpiStoreF32>(A, lda, B, ldb, M, N, K, 4, ep, stream);
    return launch_gemm<kCtas, BN, kStages, true, true, EpiStoreF32>(A, lda, B, ldb, M, N, K, 4, ep, stream);
}

extern "C" int skyb200_gemm_selftest(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K,
                                     int32_t a_mn, int32_t b_mn, int32_t ctas, int32_t bn, int32_t accumulate,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(A && B && C, "null pointer");
    int rc = skyb200_check_device();
    if (rc) return rc;
    if (ctas == 1 && bn == 256)
        return selftest_dispatch<1, 256, 4>(A, B, C, (int)M, (int)N, (int)K, a_mn, b_mn, accumulate, stream);
    if (ctas == 1 && bn == 128)
        return selftest_dispatch<1, 128, 6>(A, B, C, (int)M, (int)N, (int)K, a_mn, b_mn, accumulate, stream);
    if (ctas == 2 && bn == 256)
        return selftest_dispatch<2, 256, 6>(A, B, C, (int)M, (int)N, (int)K, a_mn, b_mn, accumulate, stream);
    if (ctas == 2 && bn == 512)
        return selftest_dispatch<2, 512, 4>(A, B, C, (int)M, (int)N, (int)K, a_mn, b_mn, accumulate, stream);
    if (ctas == 2 && bn == 128)
        return selftest_dispatch<2, 128, 6>(A, B, C, (int)M, (int)N, (int)K, a_mn, b_mn, accumulate, stream);
    return fail(SKYB200_EINVAL, "unsupported selftest configuration ctas=%d bn=%d", ctas, bn);
}

// Tuning hook: the production GEMM configuration (cta pairs, 256x256, 6 stages) with a chosen epilogue and
// tile order, K-major operands.  epi: 0 = fp32 store, 1 = bf16 store, 2 = discard, 3 = LSE partials (head forward).
extern "C" int skyb200_gemm_bench(const void* A, const void* B, void* C, void* scratch, int64_t M, int64_t N,
                                  int64_t K, int32_t epi, int32_t group, int32_t order, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(A && B, "null pointer");
    int rc = skyb200_check_device();
    if (rc) return rc;
    if (epi == 0) {
        EpiStoreF32::Params ep{static_cast<float*>(C), N, 0};
        return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiStoreF32>(A, K, B, K, (int)M, (int)N,
                                                                                      (int)K, group, ep, stream, order);
    }
    if (epi == 1) {
        EpiStoreBF16::Params ep{static_cast<__nv_bfloat16*>(C), N};
        return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiStoreBF16>(
            A, K, B, K, (int)M, (int)N, (int)K, group, ep, stream, order);
    }
    if (epi == 2) {
        EpiNull::Params ep{static_cast<float*>(C)};
        return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiNull>(A, K, B, K, (int)M, (int)N, (int)K,
                                                                                  group, ep, stream, order);
    }
    if (epi == 3) {
        SKYB200_REQUIRE(scratch != nullptr, "epi 3 needs scratch: (3*ceil(N/256)+1)*M floats + M int64 labels");
        const int slabs = (int)num_slabs(N);
        float* part = static_cast<float*>(scratch);
        LseStoreParams ep;
        ep.scale_log2 = 1.4426950408889634f;
        ep.part_m = part;
        ep.part_s = part + (int64_t)slabs * M;
        ep.part_t = part + 2 * (int64_t)slabs * M;
        ep.ld = M;
        ep.label_acc = part + 3 * (int64_t)slabs * M;
        ep.labels = reinterpret_cast<const int64_t*>(ep.label_acc + M);
        ep.logits = static_cast<__nv_bfloat16*>(C);  // NULL: plain forward; else the fused step's forward + tile store
        ep.ld_logits = round_up(N, 64);
        if (C != nullptr) {
            if ((rc = make_tmap_bf16(&ep.tmap_out, C, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld_logits, 64, 32))) return rc;
            return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiLseStoreRel>(
                A, K, B, K, (int)M, (int)N, (int)K, group, ep, stream, order);
        }
        return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiLsePartials>(
            A, K, B, K, (int)M, (int)N, (int)K, group, static_cast<const LsePartialsParams&>(ep), stream, order);
    }
    if (epi == 4) {
        EpiSkip::Params ep{static_cast<float*>(C)};
        return launch_gemm<kHeadCtas, kHeadBN, kHeadStages, false, false, EpiSkip>(A, K, B, K, (int)M, (int)N, (int)K,
                                                                                  group, ep, stream, order);
    }
    return fail(SKYB200_EINVAL, "epi must be 0..4");
}

extern "C" int skyb200_dw_group_plan(int64_t rows, int64_t H, int64_t V) { return dw_group_n(rows, H, V); }

extern "C" int skyb200_head_tile_plan(int64_t rows, int64_t H, int64_t V, int32_t* bn_out) {
    SKYB200_REQUIRE(bn_out != nullptr, "null pointer");
    SKYB200_REQUIRE(rows > 0 && H > 0 && V > 0, "bad shape");
    bn_out[0] = kHeadBN;
    bn_out[1] = use_wide(1, (rows + kTileM - 1) / kTileM, H) ? kWideBN : kHeadBN;
    bn_out[2] = use_wide(2, (V + kTileM - 1) / kTileM, H) ? kWideBN : kHeadBN;
    bn_out[3] = dw_group_n(rows, H, V, bn_out[2]);
    return SKYB200_OK;
}

// Tuning hook: ONE launch of a backward GEMM of the fused step exactly as the step issues it (operand majors, tile
// order, L2 hints, lockstep), on caller-provided buffers: which = 1: d_hidden[rows, H] = S[rows, V] * W[V, H];
// which = 2: d_weight[V, H] (+)= S^T * hidden[rows, H].  S is the bf16 d(logits) store with pitch round_up(V, 64).
extern "C" int skyb200_head_gemm_bench(int32_t which, const void* hidden, const void* weight, const void* S,
                                       void* d_hidden, float* d_weight, int32_t accumulate, int64_t rows, int64_t H,
                                       int64_t V, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SKYB200_REQUIRE(S != nullptr && rows > 0 && H % 8 == 0, "bad arguments");
    int rc = skyb200_check_device();
    if (rc) return rc;
    const int64_t ldS = round_up(V, 64);
    if (which == 1) {
        SKYB200_REQUIRE(weight && d_hidden, "null pointer");
        EpiStoreBF16::Params eh{static_cast<__nv_bfloat16*>(d_hidden), H};
        return launch_d_hidden(S, ldS, weight, (int)rows, H, V, eh, stream, nullptr);
    }
    if (which == 2) {
        SKYB200_REQUIRE(hidden && d_weight, "null pointer");
        EpiTmaF32::Params ew;
        if ((rc = make_tmap_f32_out(&ew.tmap_out, d_weight, (uint64_t)H, (uint64_t)V, (uint64_t)H))) return rc;
        ew.accumulate = accumulate;
        ew.hint = hint_dw_c();
        return launch_d_weight<EpiTmaF32>(S, ldS, hidden, (int)rows, H, V, ew, stream, nullptr);
    }
    return fail(SKYB200_EINVAL, "which must be 1 (d_hidden) or 2 (d_weight)");
}

extern "C" const uint32_t* skyb200_gemm_debug_record(void) {
    debug_record_dev();
    return reinterpret_cast<const uint32_t*>(g_dbg_host);
}

// ------------------------------------------------------------------------------------------ peer mapping (CUDA IPC)
extern "C" int skyb200_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out) {
    SKYB200_REQUIRE(ptr && handle_out && offset_out, "null pointer");
    int rc = skyb200_check_device();
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUdeviceptr base = 0;
    size_t size = 0;
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range_fn = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
        return reinterpret_cast<RangeFn>(fn);
    }();
    if (range_fn == nullptr) return fail(SKYB200_ECUDA, "cuMemGetAddressRange entry point not available");
    const CUresult cr = range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
    if (cr != CUDA_SUCCESS) return fail(SKYB200_ECUDA, "cuMemGetAddressRange failed with CUresult %d", (int)cr);
    cudaIpcMemHandle_t h;
    SKYB200_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return SKYB200_OK;
}

namespace {
std::mutex g_ipc_mu;
std::vector<std::pair<void*, void*>> g_ipc_open;  // (pointer handed out, base returned by cudaIpcOpenMemHandle)
}  // namespace

extern "C" int skyb200_ipc_open(const void* handle, int64_t offset, void** ptr_out) {
    SKYB200_REQUIRE(handle && ptr_out && offset >= 0, "bad arguments");
    int rc = skyb200_check_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    SKYB200_CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = static_cast<char*>(base) + offset;
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    g_ipc_open.emplace_back(*ptr_out, base);
    return SKYB200_OK;
}

extern "C" int skyb200_ipc_close(void* ptr) {
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    for (size_t i = 0; i < g_ipc_open.size(); ++i) {
        if (g_ipc_open[i].first == ptr) {
            void* base = g_ipc_open[i].second;
            g_ipc_open.erase(g_ipc_open.begin() + i);
            SKYB200_CUDA_CHECK(cudaIpcCloseMemHandle(base));
            return SKYB200_OK;
        }
    }
    return fail(SKYB200_EINVAL, "skyb200_ipc_close: pointer was not returned by skyb200_ipc_open");
}
