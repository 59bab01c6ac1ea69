// Persistent warp-specialised bf16 GEMM for sm_100a with pluggable TMEM epilogues.
//
//   C[M,N] = A[M,K] * B[N,K]^T        (fp32 accumulation in tensor memory)
//
// Operands are staged by TMA into 128B-swizzled shared memory, multiplied by tcgen05.mma (one issuing
// thread), accumulated in TMEM (double-buffered, 2 x BN columns) and drained by four epilogue warps
// (thread i <-> accumulator row i) through tcgen05.ld.  kCtas == 2 pairs two SMs on one 256-row tile
// (cta_group::2): each CTA loads its own 128 A rows and half of the B tile.
//
// BN == 512 is the WIDE form for deep-K GEMMs: one 256 x 512 tile per CTA pair fills all 512 TMEM columns (a single
// accumulator: the epilogue no longer overlaps the next tile's mainloop, which costs nothing when a tile spends
// hundreds of K blocks in the mainloop) and every K block issues two N = 256 MMA groups against the same A stage.
// Operand bytes fetched per flop fall by 25 % ((256 + 512) / (2 * 256 * 512) vs (256 + 256) / (2 * 256 * 256)): the
// 256 x 256 form needs ~12 TB/s of L2 -> SM traffic at 1.5 PFLOP/s, which is what the chip delivers at best
// (profiles/README.md, r02o: all three GEMMs move 80.7 GB per launch and none finishes under 6.4 ms at any clock).
//
// Either operand may be K-major (rows of K contiguous elements) or MN-major (stored [K, MN] with MN
// contiguous); the backward GEMMs of the fused LM head use the MN-major forms so no transposed copies of
// the weight / hidden / dlogits matrices are ever made.
//
// Warp roles (256 threads):  warps 0..3 = epilogue, warp 4 = TMA producer, warp 5 = MMA issuer (leader CTA
// only), warp 6 = TMEM allocator; warp 7 idles at the final barrier.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx.cuh"
#include "tile_order.cuh"

namespace skyb200 {

constexpr int kBM = 128;  // accumulator rows per CTA (= TMEM lanes)
constexpr int kBK = 64;   // bf16 elements per 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;
// Warp roles: epilogue warps 0..3 (TMEM lane quarter = warp id), then the three control warps.  The control
// loops run warp-converged (uniform registers) with one elected lane issuing TMA / tcgen05 instructions: the
// earlier single-lane form made ptxas emit divergent "waterfall" code around every uniform-datapath operand and
// cost ~5 % of the tensor pipe (measured, profiles/r01c_*).
constexpr int kEpiWarp0 = 0, kTmaWarp = 4, kMmaWarp = 5, kAllocWarp = 6;

// L2 eviction policies (the encodings cuTensorMap / createpolicy produce for fraction 1.0)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

template <int kCtas>
__device__ __forceinline__ void tma_load_2d_hint(const void* desc, uint64_t* bar, void* smem_dst, int32_t crd0,
                                                 int32_t crd1, uint64_t policy) {
    if constexpr (kCtas == 2) {
        const uint32_t bar_addr = ptx::smem_u32(bar) & 0xFEFFFFFFu;  // leader CTA's barrier
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;"
            :
            : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(crd0),
              "r"(crd1), "l"(policy)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;"
            :
            : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(ptx::smem_u32(bar)),
              "r"(crd0), "r"(crd1), "l"(policy)
            : "memory");
    }
}

// Debug record written to mapped host memory when a barrier wait times out (so a protocol bug shows up
// as a failed launch with a readable reason instead of a hung GPU).
struct GemmDebug {
    unsigned int flag;
    unsigned int role;
    unsigned int which;
    unsigned int stage;
    unsigned int parity;
    unsigned int tile;
    unsigned int block;
    unsigned int pad;
};

__device__ __forceinline__ void bounded_wait(uint64_t* bar, uint32_t parity, GemmDebug* dbg, uint32_t role,
                                             uint32_t which, uint32_t stage, uint32_t tile) {
    uint32_t spins = 0;
    // fast path first (the barrier is usually already complete for the MMA lane), then sleep-waits
    if (ptx::mbar_try_wait(bar, parity)) return;
    while (!ptx::mbar_try_wait_hint(bar, parity, 2000u)) {
        if (++spins > (1u << 22)) {  // each try is a hardware sleep of up to 2 us: this is seconds
            if (dbg != nullptr && atomicExch(&dbg->flag, 1u) == 0u) {
                dbg->role = role;
                dbg->which = which;
                dbg->stage = stage;
                dbg->parity = parity;
                dbg->tile = tile;
                dbg->block = blockIdx.x;
                __threadfence_system();
            }
            __trap();
        }
    }
}

// Tile geometry.  A tile of BN columns is computed as kHalves MMA groups of kMmaN <= 256 columns (the UMMA limit); the
// B stage of a CTA is kHalves sub-blocks of kSubN rows, sub-block h = columns h * kMmaN + cta_rank * kSubN + [0, kSubN)
// of the tile.  BN <= 256: two accumulator buffers (epilogue of tile t under the mainloop of tile t + 1); BN == 512: one.
template <int kCtas, int BN>
struct GemmTile {
    static constexpr int kMmaN = BN > 256 ? 256 : BN;
    static constexpr int kHalves = BN / kMmaN;
    static constexpr int kAcc = BN > 256 ? 1 : 2;
    static constexpr int kSubN = kMmaN / kCtas;
    static constexpr uint32_t kSubBytes = kSubN * kBK * 2;
};

// Shared memory: [A stages][B stages][epilogue staging: kEpiBytes, 1024-byte aligned][barriers]
template <int kCtas, int BN, int kStages, uint32_t kEpiBytes = 0>
struct GemmSmem {
    static constexpr uint32_t kABytes = kBM * kBK * 2;
    static constexpr uint32_t kLoadN = BN / kCtas;
    static constexpr uint32_t kBBytes = kLoadN * kBK * 2;
    static constexpr uint32_t kBarBytes = (2 * kStages + 4) * 8 + 16;
    static constexpr uint32_t kEpiOffset = kStages * (kABytes + kBBytes);
    static constexpr uint32_t kTotal = 1024 + kStages * (kABytes + kBBytes) + kEpiBytes + kBarBytes;
    static_assert(kEpiBytes % 1024 == 0, "epilogue staging must keep the 1024-byte swizzle alignment");
    static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

// Optional device-side compaction of the iteration space (all pointers may be null):
//   m_map / m_count : the kernel visits only the M tiles m_map[0 .. *m_count) (skipped tiles are not written);
//   k_map / k_count : the reduction visits only the K blocks k_map[0 .. *k_count) (in that order).
// Used by the fused policy step to skip token rows whose d(logits) is identically zero (masked positions).
// With *k_count == 0 the launch does nothing (callers use it only with an accumulating epilogue).
struct GemmMaps {
    const int* m_map;
    const int* m_count;
    const int* k_map;
    const int* k_count;
};

template <int kCtas, int BN, int kStages, bool kAMn, bool kBMn, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M,
                  int N, int K, int group_m, int tile_order, uint64_t hint_a, uint64_t hint_b,
                  const __grid_constant__ typename Epi::Params ep, GemmDebug* dbg, unsigned int* sync_ctr,
                  int sync_every, int sync_slack, GemmMaps maps) {
    using S = GemmSmem<kCtas, BN, kStages, Epi::kSmemBytes>;
    using T = GemmTile<kCtas, BN>;
    constexpr int TILE_M = kBM * kCtas;
    constexpr int kMmaN = T::kMmaN, kHalves = T::kHalves, kAcc = T::kAcc, kSubN = T::kSubN;
    constexpr uint32_t kSubBytes = T::kSubBytes;
    constexpr uint32_t kTmemCols = kAcc * BN;
    static_assert(BN == 128 || BN == 256 || BN == 512, "BN must be 128, 256 or 512");
    static_assert(kTmemCols <= 512, "TMEM overflow");

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (required by the 128B swizzle atoms); identical in both CTAs of a pair.
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * S::kABytes;
    uint8_t* smem_epi = smem + S::kEpiOffset;  // Epi::kSmemBytes of staging for TMA-store epilogues (4 equal warp slices)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + Epi::kSmemBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (kCtas == 2) ? ptx::cluster_ctarank() : 0u;

    if (warp == kTmaWarp && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], kCtas * 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kAllocWarp) {
        ptx::tmem_alloc<kCtas>(tmem_slot, kTmemCols);
    }
    ptx::tcgen05_fence_before();
    if constexpr (kCtas == 2) {
        ptx::cluster_sync();
    } else {
        __syncthreads();
    }
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const int num_m_tiles = maps.m_count ? *maps.m_count : (M + TILE_M - 1) / TILE_M;
    const int num_n_tiles = (N + BN - 1) / BN;
    const int num_kb = maps.k_count ? *maps.k_count : (K + kBK - 1) / kBK;
    const int total_tiles = num_kb > 0 ? num_m_tiles * num_n_tiles : 0;
    const int cluster_id = blockIdx.x / kCtas;
    const int num_clusters = gridDim.x / kCtas;

    if (warp == kTmaWarp) {
        // ------------------------------------------------------------ TMA producer
        // The whole warp runs the (warp-uniform) loop so that stage / phase / addresses live in uniform registers;
        // one elected lane issues the expect_tx and the TMA copies.
        {
            const bool issuer = ptx::elect_one();
            uint32_t stage = 0, phase = 0;
            uint32_t sync_step = 0, sync_epoch = 0, pend_epoch = 0, pend_active = 0;
            bool pending = false;
            for (int t = cluster_id; t < total_tiles; t += num_clusters) {
                const TileCoord tc = decode_tile(t, num_m_tiles, num_n_tiles, group_m, tile_order);
                const int m_tile = maps.m_map ? maps.m_map[tc.m_tile] : tc.m_tile;
                const int m0 = m_tile * TILE_M + static_cast<int>(cta_rank) * kBM;
                const int n0 = tc.n_tile * BN + static_cast<int>(cta_rank) * kSubN;
                int k_window = 0;  // lane l holds k_map[32 * (kb / 32) + l]: one coalesced load per 32 K blocks
                // CTAs that are resident at this tile iteration (the last iteration may be partial)
                const uint32_t active_ctas =
                    (t - cluster_id + num_clusters <= total_tiles ? num_clusters : total_tiles % num_clusters) * kCtas;
                for (int kb = 0; kb < num_kb; ++kb) {
                    int kidx = kb;
                    if (maps.k_map != nullptr) {
                        if ((kb & 31) == 0) k_window = (kb + lane < num_kb) ? maps.k_map[kb + lane] : 0;
                        kidx = __shfl_sync(0xffffffffu, k_window, kb & 31);
                    }
                    // Best-effort lockstep: CTAs that share operand panels through L2 re-align every `sync_every`
                    // K blocks (bounded wait: a straggler can never dead-lock the others).  Split phase: a CTA
                    // ARRIVES at the epoch counter at the period boundary but only WAITS for the others
                    // `sync_slack` K blocks later - the drift the CTAs accumulate inside a period is absorbed while
                    // the producer keeps the ring full, instead of starving the tensor pipe at the barrier.
                    if (sync_ctr != nullptr) {
                        const uint32_t ph = sync_step++ % static_cast<uint32_t>(sync_every);
                        if (ph == 0) {
                            if (issuer) atomicAdd(sync_ctr + (sync_epoch & 8191u), 1u);
                            pend_epoch = sync_epoch++;
                            pend_active = active_ctas;
                            pending = true;
                        }
                        if (pending && ph == static_cast<uint32_t>(sync_slack)) {
                            if (issuer) {
                                unsigned int* c = sync_ctr + (pend_epoch & 8191u);
                                for (int spin = 0; spin < 400; ++spin) {
                                    unsigned int v;
                                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c));
                                    if (v >= pend_active) break;
                                    __nanosleep(50);
                                }
                            }
                            pending = false;
                            __syncwarp();
                        }
                    }
                    bounded_wait(&empty_bar[stage], phase ^ 1u, dbg, 0, 0, stage, t);
                    if (issuer) {
                    if (cta_rank == 0) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], (S::kABytes + S::kBBytes) * kCtas);
                    }
                    uint8_t* a_dst = smem_a + stage * S::kABytes;
                    uint8_t* b_dst = smem_b + stage * S::kBBytes;
                    const int k0 = kidx * kBK;
                    if constexpr (kAMn) {
#pragma unroll
                        for (int j = 0; j < kBM / 64; ++j) {
                            tma_load_2d_hint<kCtas>(&tmap_a, &full_bar[stage], a_dst + j * 8192, m0 + j * 64, k0, hint_a);
                        }
                    } else {
                        tma_load_2d_hint<kCtas>(&tmap_a, &full_bar[stage], a_dst, k0, m0, hint_a);
                    }
#pragma unroll
                    for (int h = 0; h < kHalves; ++h) {
                        if constexpr (kBMn) {
#pragma unroll
                            for (int j = 0; j < kSubN / 64; ++j) {
                                tma_load_2d_hint<kCtas>(&tmap_b, &full_bar[stage], b_dst + h * kSubBytes + j * 8192,
                                                        n0 + h * kMmaN + j * 64, k0, hint_b);
                            }
                        } else {
                            tma_load_2d_hint<kCtas>(&tmap_b, &full_bar[stage], b_dst + h * kSubBytes, k0, n0 + h * kMmaN,
                                                    hint_b);
                        }
                    }
                    }  // issuer
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        // Warp-converged loop (uniform registers for stage / phase / descriptors, no divergent waterfall code);
        // the same elected lane issues every tcgen05.mma and tcgen05.commit.
        if (cta_rank == 0) {
            const bool issuer = ptx::elect_one();
            constexpr uint32_t idesc = ptx::make_idesc_bf16(TILE_M, kMmaN, kAMn ? 1u : 0u, kBMn ? 1u : 0u);
            // K-major: 8-row groups are 1024 B apart, K advances 32 B inside the swizzle atom.
            // MN-major: 64-element MN groups are one [64 x 64] box (8192 B) apart, 8-row K groups 1024 B.
            constexpr uint32_t a_lbo = kAMn ? 8192u : 16u, a_sbo = 1024u, a_kstep = kAMn ? 2048u : 32u;
            constexpr uint32_t b_lbo = kBMn ? 8192u : 16u, b_sbo = 1024u, b_kstep = kBMn ? 2048u : 32u;
            uint32_t stage = 0, phase = 0, accum = 0, aphase = 0;
            for (int t = cluster_id; t < total_tiles; t += num_clusters) {
                bounded_wait(&tempty_bar[accum], aphase ^ 1u, dbg, 1, 1, accum, t);
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + accum * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    bounded_wait(&full_bar[stage], phase, dbg, 1, 0, stage, t);
                    ptx::tcgen05_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem_a + stage * S::kABytes);
                    const uint32_t b_addr = ptx::smem_u32(smem_b + stage * S::kBBytes);
                    if (issuer) {
#pragma unroll
                        for (int j = 0; j < kBK / kUmmaK; ++j) {
                            const uint64_t da = ptx::make_smem_desc_sw128(a_addr + j * a_kstep, a_lbo, a_sbo);
#pragma unroll
                            for (int h = 0; h < kHalves; ++h) {  // the MMA groups of one K step share the A operand
                                const uint64_t db =
                                    ptx::make_smem_desc_sw128(b_addr + h * kSubBytes + j * b_kstep, b_lbo, b_sbo);
                                ptx::umma_bf16<kCtas>(tmem_d + h * kMmaN, da, db, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                            }
                        }
                        ptx::umma_commit<kCtas>(&empty_bar[stage]);
                        if (kb == num_kb - 1) ptx::umma_commit<kCtas>(&tfull_bar[accum]);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++accum == kAcc) {
                    accum = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        // ------------------------------------------------------------ epilogue (4 warps, row per thread)
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        uint8_t* stage_w = smem_epi + q * (Epi::kSmemBytes / 4);
        uint32_t accum = 0, aphase = 0;
        for (int t = cluster_id; t < total_tiles; t += num_clusters) {
            const TileCoord tc = decode_tile(t, num_m_tiles, num_n_tiles, group_m, tile_order);
            bounded_wait(&tfull_bar[accum], aphase, dbg, 2, 2, accum, t);
            ptx::tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + accum * BN;
            const int m_tile = maps.m_map ? maps.m_map[tc.m_tile] : tc.m_tile;
            const int row = m_tile * TILE_M + static_cast<int>(cta_rank) * kBM + q * 32 + lane;
            Epi::template run<BN>(ep, taddr, row, tc.n_tile, M, N, stage_w);
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kCtas == 2)
                    ptx::mbar_arrive_cluster(&tempty_bar[accum], 0);
                else
                    ptx::mbar_arrive(&tempty_bar[accum]);
            }
            if (++accum == kAcc) {
                accum = 0;
                aphase ^= 1u;
            }
        }
        Epi::finish();  // TMA-store epilogues: this warp's outstanding bulk stores complete before the CTA exits
    }

    ptx::tcgen05_fence_before();
    if constexpr (kCtas == 2) {
        ptx::cluster_sync();
    } else {
        __syncthreads();
    }
    if (warp == kAllocWarp) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc<kCtas>(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------ epilogues

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// 256-bit global accesses (sm_100): one instruction covers a full 32-byte sector, so partial-sector
// read-for-fill traffic at L2 disappears for row-per-thread epilogue stores.
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                              uint32_t a5, uint32_t a6, uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2),
                 "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7)
                 : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p)
                 : "memory");
}
// Same, with an L2 evict-first policy: for data that is written (or read) exactly once by this kernel, so it does
// not push the operand panels the other CTAs are about to reuse out of L2.
__device__ __forceinline__ void st_global_256_stream(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                     uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p), "r"(a0),
                 "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7), "l"(kEvictFirst)
                 : "memory");
}
__device__ __forceinline__ void ld_global_256_stream(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.L2::cache_hint.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p), "l"(kEvictFirst)
                 : "memory");
}

// 32 fp32 values -> 32 bf16 -> two 32-byte stores (dst must be 32-byte aligned)
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const float (&d)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        st_global_256_stream(dst + 16 * h, pack_bf16x2(d[16 * h + 0], d[16 * h + 1]), pack_bf16x2(d[16 * h + 2], d[16 * h + 3]),
                      pack_bf16x2(d[16 * h + 4], d[16 * h + 5]), pack_bf16x2(d[16 * h + 6], d[16 * h + 7]),
                      pack_bf16x2(d[16 * h + 8], d[16 * h + 9]), pack_bf16x2(d[16 * h + 10], d[16 * h + 11]),
                      pack_bf16x2(d[16 * h + 12], d[16 * h + 13]), pack_bf16x2(d[16 * h + 14], d[16 * h + 15]));
    }
}

// C (fp32, row-major, leading dim ldc) = or += accumulator.   Used for dW and for GEMM self-tests.
struct EpiStoreF32 {
    static constexpr uint32_t kSmemBytes = 0;
    __device__ static void finish() {}
    struct Params {
        float* C;
        int64_t ldc;
        int accumulate;
    };
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t*) {
        const int n0 = n_tile * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (row < M && col0 < N) {
                float* dst = p.C + static_cast<int64_t>(row) * p.ldc + col0;
                if (col0 + 32 <= N && (p.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 31u) == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t o[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) o[k] = v[8 * i + k];
                        if (p.accumulate) {
                            uint32_t old[8];
                            ld_global_256_stream(dst + 8 * i, old);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                o[k] = __float_as_uint(__uint_as_float(o[k]) + __uint_as_float(old[k]));
                        }
                        st_global_256_stream(dst + 8 * i, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
                    }
                } else {
                    for (int i = 0; i < 32 && col0 + i < N; ++i) {
                        const float a = __uint_as_float(v[i]);
                        dst[i] = p.accumulate ? dst[i] + a : a;
                    }
                }
            }
        }
    }
};

// C (fp32 [M, N], pitch in the tensor map) = or += accumulator, through shared memory and TMA: each epilogue warp stages
// 32 rows x 32 columns (4 KB, 128-byte swizzle, two buffers) and one lane issues cp.async.bulk.tensor (store) or
// cp.reduce.async.bulk.tensor ... .add (UTMAREDG: the fp32 add is performed at L2 - the SM neither reads the old value
// nor waits for it).  Used for d_weight: the per-thread read-modify-write of EpiStoreF32 moved 4.4 GB per chunk
// through the LSU in 32-byte row-strided pieces.  Rows >= M / columns >= N are clipped by the tensor map.
struct EpiTmaF32 {
    static constexpr uint32_t kSmemBytes = 4 * 2 * 4096;
    __device__ static void finish() {
        if ((threadIdx.x & 31) == 0) ptx::bulk_wait_group<0>();
    }
    struct Params {
        CUtensorMap tmap_out;  // fp32 [M, N], box 32 x 32, SWIZZLE_128B
        int accumulate;
        uint64_t hint;  // L2 eviction policy of the output traffic (kEvictNormal / kEvictFirst)
    };
    // One 32 x 32 piece: registers -> swizzled staging buffer -> bulk store / reduce-add.
    __device__ static __forceinline__ void emit(const Params& p, const uint32_t (&v)[32], uint8_t* stage, uint32_t srow,
                                                uint32_t sw, int c, int col0, int row, int M, int lane) {
        if (lane == 0) {  // the buffer about to be rewritten must have been read by its previous bulk operation
            if (c == 0)
                ptx::bulk_wait_group_read<0>();
            else
                ptx::bulk_wait_group_read<1>();
        }
        __syncwarp();
        const uint32_t sbuf = srow + static_cast<uint32_t>(c & 1) * 4096u;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {  // 16-byte chunk j of the 128-byte row goes to slot j ^ (row & 7)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + ((j ^ sw) << 4)), "r"(v[4 * j]),
                         "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                         : "memory");
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row < M) {  // lane 0 holds the warp's first row
            if (p.accumulate)
                ptx::tma_reduce_add_2d_hint(&p.tmap_out, stage + (c & 1) * 4096, col0, row, p.hint);
            else
                ptx::tma_store_2d_hint(&p.tmap_out, stage + (c & 1) * 4096, col0, row, p.hint);
            ptx::bulk_commit_group();
        }
    }
    // Software-pipelined over pairs of pieces: the tcgen05.ld of piece c + 1 is in flight while piece c is staged and
    // handed to the copy engine (with the wide tiles the epilogue is no longer hidden behind the next tile's mainloop).
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t* stage) {
        const int n0 = n_tile * BN;
        const int lane = threadIdx.x & 31;
        const uint32_t srow = ptx::smem_u32(stage) + static_cast<uint32_t>(lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        const int pieces = min(BN / 32, (N - n0 + 31) / 32);  // uniform across the whole warp; >= 1
        uint32_t va[32], vb[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < pieces; c += 2) {
            ptx::tmem_ld_wait();
            __syncwarp();
            if (c + 1 < pieces) ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
            emit(p, va, stage, srow, sw, c, n0 + c * 32, row, M, lane);
            if (c + 1 < pieces) {
                ptx::tmem_ld_wait();
                __syncwarp();
                if (c + 2 < pieces) ptx::tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
                emit(p, vb, stage, srow, sw, c + 1, n0 + (c + 1) * 32, row, M, lane);
            }
        }
    }
};

// d_weight GEMM fused with the data-parallel REDUCE-SCATTER of the gradient (FSDP semantics: rank r keeps rows
// [r * rows_per_rank, (r + 1) * rows_per_rank) of the rank-averaged [M, N] fp32 gradient).  Every accumulator piece
// (32 rows x 32 columns, staged in shared memory exactly as in EpiTmaF32) is scaled by 1 / world and ADDED straight
// into its owner's shard while the tensor cores work on the next tile, through TMA reduce-add (UTMAREDG) on a tensor
// map over the owner's shard - the rank's own buffer or a peer-mapped one: NVLink carries the reduction to the owner's
// L2, which performs the add.  (remote_red = 1 sends the peers' pieces as 16-byte fp32 vector reductions,
// red.relaxed.sys.global.add.v4.f32, instead - a warp instruction covers 4 rows x 128 contiguous bytes; measured 14 %
// slower per launch at 2 ranks, kept as the fallback.)  No full-size local gradient buffer exists and no
// collective follows the step: the transfer is spread over every chunk's GEMM (1.9 GB per 9472-token chunk at 8 ranks,
// ~8 % of one NVLink direction), the only cross-rank synchronisation is the caller's zero / done barriers.
// The order in which the ranks' contributions reach an element is not fixed: results are reproducible to fp32
// rounding, not bit for bit (callers that need that use the local buffer + NCCL).
struct EpiF32Scatter {
    static constexpr uint32_t kSmemBytes = 4 * 2 * 4096;
    static constexpr int kMaxWorld = 8;
    __device__ static void finish() {
        if ((threadIdx.x & 31) == 0) ptx::bulk_wait_group<0>();
    }
    struct Params {
        CUtensorMap tmap[kMaxWorld];  // tmap[r]: rank r's shard, fp32 [rows_per_rank, N], box 32 x 32, SWIZZLE_128B
        float* shard[kMaxWorld];  // shard[r] = rank r's shard as mapped into this process (peer memory for r != rank)
        int remote_red;           // 0: peers' pieces through TMA reduce-add as well; 1: red.v4.f32 from the threads
        int64_t ld;              // row pitch of a shard in elements
        int rows_per_rank;       // multiple of 32: a warp's 32 rows never straddle two owners
        int rank, world;
        float scale;
        uint64_t hint;
    };
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t* stage) {
        const int n0 = n_tile * BN;
        const int lane = threadIdx.x & 31;
        const int row0 = row - lane;  // the warp's first row (a multiple of 32)
        const int owner = row0 / p.rows_per_rank;
        const int lrow0 = row0 - owner * p.rows_per_rank;  // first row inside the owner's shard
        const bool local = (owner == p.rank) || (p.remote_red == 0);  // "local" = leaves through TMA reduce-add
        const uint32_t sbase = ptx::smem_u32(stage);
        const uint32_t srow = sbase + static_cast<uint32_t>(lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        float* const dst = (row0 < M && owner < p.world) ? p.shard[owner] : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= N) break;  // uniform across the whole warp
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            if (lane == 0) {  // the buffer about to be rewritten must have been read by its previous bulk operation
                if (c == 0)
                    ptx::bulk_wait_group_read<0>();
                else
                    ptx::bulk_wait_group_read<1>();
            }
            __syncwarp();
            const uint32_t boff = static_cast<uint32_t>(c & 1) * 4096u;
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {  // 16-byte chunk j of the 128-byte row goes to slot j ^ (row & 7)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + boff + ((j ^ sw) << 4)),
                             "f"(__uint_as_float(v[4 * j]) * p.scale), "f"(__uint_as_float(v[4 * j + 1]) * p.scale),
                             "f"(__uint_as_float(v[4 * j + 2]) * p.scale), "f"(__uint_as_float(v[4 * j + 3]) * p.scale)
                             : "memory");
            }
            if (local) {
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0 && dst != nullptr) {
                    ptx::tma_reduce_add_2d_hint(&p.tmap[owner], stage + (c & 1) * 4096, col0, lrow0, p.hint);
                    ptx::bulk_commit_group();
                }
            } else {
                __syncwarp();
                if (dst != nullptr) {
                    const int k = lane & 7;          // 16-byte chunk inside the 128-byte row
                    const int col = col0 + 4 * k;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = 4 * j + (lane >> 3);
                        float4 q;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                                     : "r"(sbase + boff + static_cast<uint32_t>(r) * 128u +
                                           ((static_cast<uint32_t>(k) ^ static_cast<uint32_t>(r & 7)) << 4))
                                     : "memory");
                        if (row0 + r < M && col < N) {
                            float* g = dst + static_cast<int64_t>(lrow0 + r) * p.ld + col;
                            asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(q.x),
                                         "f"(q.y), "f"(q.z), "f"(q.w)
                                         : "memory");
                        }
                    }
                }
            }
        }
    }
};

// Drains the accumulator and discards it (mainloop-only timing in scripts/gemm_tune.py).
struct EpiNull {
    static constexpr uint32_t kSmemBytes = 0;
    __device__ static void finish() {}
    struct Params {
        float* sink;
    };
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t*) {
        float acc = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        }
        if (acc == 1.2345e30f && p.sink != nullptr) p.sink[0] = acc + row + n_tile + M + N;
    }
};

// Does not even read the accumulator (isolates the cost of the TMEM drain itself in scripts/gemm_tune.py).
struct EpiSkip {
    static constexpr uint32_t kSmemBytes = 0;
    __device__ static void finish() {}
    struct Params {
        float* sink;
    };
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t*) {
        if (row == -12345 && p.sink != nullptr) p.sink[0] = static_cast<float>(taddr + n_tile + M + N);
    }
};

// C (bf16, row-major) = accumulator.   Used for d_hidden.
struct EpiStoreBF16 {
    static constexpr uint32_t kSmemBytes = 0;
    __device__ static void finish() {}
    struct Params {
        __nv_bfloat16* C;
        int64_t ldc;
    };
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t*) {
        const int n0 = n_tile * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (row < M && col0 < N) {
                __nv_bfloat16* dst = p.C + static_cast<int64_t>(row) * p.ldc + col0;
                if (col0 + 32 <= N && (p.ldc & 15) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 31u) == 0) {
                    float d[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(v[i]);
                    store_row32_bf16(dst, d);
                } else {
                    for (int i = 0; i < 32 && col0 + i < N; ++i) dst[i] = __float2bfloat16_rn(__uint_as_float(v[i]));
                }
            }
        }
    }
};

// Forward LM-head epilogue: per (row, vocab-slab) online-softmax partials in log2 units.
//   y = acc * scale_log2  (scale_log2 = log2(e) / temperature)
//   part_m = max_j y_j ; part_s = sum_j 2^(y_j - part_m) ; part_t = sum_j 2^(y_j - part_m) * acc_j
// and the raw accumulator at the label column (written by the one slab that contains it).
struct LsePartialsParams {
    float scale_log2;
    float* part_m;  // [num_n_tiles][ld]
    float* part_s;
    float* part_t;
    int64_t ld;
    const int64_t* labels;  // [M]
    float* label_acc;       // [M]
    __nv_bfloat16* logits;  // EpiLseStoreRel only: [M, ld_logits] bf16 slab-relative log2-logits (see below)
    int64_t ld_logits;
};

// Plain forward (no backward follows, or the backward recomputes): one pass over the accumulator tile with a running
// maximum that rescales the sums when it moves.
struct EpiLsePartials {
    static constexpr uint32_t kSmemBytes = 0;
    __device__ static void finish() {}
    using Params = LsePartialsParams;
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t*) {
        const int n0 = n_tile * BN;
        const bool row_ok = row < M;
        const int lab = row_ok ? static_cast<int>(p.labels[row]) : -1;
        float m2 = -INFINITY, s = 0.f, tt = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (col0 >= N) continue;  // uniform across the whole warp
            const int nvalid = min(32, N - col0);
            float cm = -INFINITY;
            if (nvalid == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(v[i]));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nvalid) cm = fmaxf(cm, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m2, cm * p.scale_log2);
            const float corr = ex2_approx(m2 - m_new);  // 0 on the first chunk (m2 = -inf)
            s *= corr;
            tt *= corr;
            m2 = m_new;
            if (nvalid == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float a = __uint_as_float(v[i]);
                    const float e = ex2_approx(fmaf(a, p.scale_log2, -m2));
                    s += e;
                    tt = fmaf(e, a, tt);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i < nvalid) {
                        const float a = __uint_as_float(v[i]);
                        const float e = ex2_approx(fmaf(a, p.scale_log2, -m2));
                        s += e;
                        tt = fmaf(e, a, tt);
                    }
                }
            }
            if (lab >= col0 && lab < col0 + nvalid) {
                float z = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (col0 + i == lab) z = __uint_as_float(v[i]);
                p.label_acc[row] = z;
            }
        }
        if (row_ok) {
            const int64_t o = static_cast<int64_t>(n_tile) * p.ld + row;
            p.part_m[o] = m2;
            p.part_s[o] = s;
            p.part_t[o] = tt;
        }
    }
};

// Forward of the fused step: the same partials, plus a bf16 copy of the tile for the backward GEMMs (which therefore
// do not recompute the projection).  What is stored is NOT the raw logit but the SLAB-RELATIVE log2-logit
//     u_j = acc_j * scale_log2 - part_m            (<= 0, and ~0 exactly where the probability mass is)
// so that the bf16 rounding error |u| * 2^-9 vanishes on the entries that carry the gradient: the backward forms
// p_j = 2^(u_j + part_m - lse2) with a relative error of 2^-9 * ln2 * log2(p_max / p_j) - independent of the logit
// SCALE - whereas rounding the raw logit z costs |z| * 2^-9 on every entry (28 % of p at |z| = 70).  The slab
// maximum must be known before the first store, so the tile is read from TMEM twice: pass 1 takes the row maximum
// over the slab, pass 2 exponentiates against it (no running-maximum rescale any more), accumulates the two sums and
// writes u.  The ex2 argument IS u: the store costs one cvt.bf16x2 per two columns and no extra MUFU.
//
// The tile leaves the SM through TMA: each epilogue warp stages 32 rows x 64 columns (4 KB, 128-byte swizzle, two
// buffers) in shared memory and one lane issues cp.async.bulk.tensor (UTMASTG) - full 128-byte lines written by the
// copy engine instead of 32 row-strided 32-byte LSU stores per instruction, which cost the forward GEMM 12 % of its
// tensor-pipe time (profiles/README.md, r02b).  Rows >= M and columns >= N are clipped by the tensor map.
struct LseStoreParams : LsePartialsParams {
    CUtensorMap tmap_out;  // bf16 [M, N] with pitch ld_logits, box 64 x 32, SWIZZLE_128B
};
struct EpiLseStoreRel {
    static constexpr uint32_t kSmemBytes = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)
    __device__ static void finish() {
        if ((threadIdx.x & 31) == 0) ptx::bulk_wait_group<0>();
    }
    using Params = LseStoreParams;
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t* stage) {
        const int n0 = n_tile * BN;
        const int lane = threadIdx.x & 31;
        const bool row_ok = row < M;
        const int lab = row_ok ? static_cast<int>(p.labels[row]) : -1;
        // ---- pass 1: slab maximum of the raw accumulators (scale_log2 > 0: the maximum commutes with the scaling)
        float cm = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (col0 >= N) continue;  // uniform across the whole warp
            const int nvalid = min(32, N - col0);
            if (nvalid == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(v[i]));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nvalid) cm = fmaxf(cm, __uint_as_float(v[i]));
            }
        }
        const float m2 = cm * p.scale_log2;
        // ---- pass 2: sums against the slab maximum + the slab-relative bf16 tile through shared memory and TMA
        float s = 0.f, tt = 0.f;
        const uint32_t srow = ptx::smem_u32(stage) + static_cast<uint32_t>(lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll 1
        for (int c2 = 0; c2 < BN / 64; ++c2) {
            const int col0 = n0 + c2 * 64;
            if (col0 >= N) break;  // uniform across the whole warp
            uint32_t pk[32];  // 64 bf16 of this row
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                __syncwarp();
                ptx::tmem_ld_32x32b_x32(taddr + c2 * 64 + h * 32, v);
                ptx::tmem_ld_wait();
                const int c0 = col0 + h * 32;
                const int nvalid = min(32, N - c0);  // <= 0: the whole half lies beyond N
                if (lab >= c0 && lab < c0 + nvalid) {
                    float z = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i == lab) z = __uint_as_float(v[i]);
                    p.label_acc[row] = z;
                }
                if (nvalid >= 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float a = __uint_as_float(v[i]);
                        const float u = fmaf(a, p.scale_log2, -m2);
                        const float e = ex2_approx(u);
                        s += e;
                        tt = fmaf(e, a, tt);
                        v[i] = __float_as_uint(u);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float u = 0.f;  // columns >= N: clipped by the tensor map, never stored
                        if (i < nvalid) {
                            const float a = __uint_as_float(v[i]);
                            u = fmaf(a, p.scale_log2, -m2);
                            const float e = ex2_approx(u);
                            s += e;
                            tt = fmaf(e, a, tt);
                        }
                        v[i] = __float_as_uint(u);
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    pk[16 * h + i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            }
            // the buffer about to be rewritten was the source of the store issued two stores ago (one store ago at a
            // tile start after a ragged tile: wait for all there); only lane 0 issues, so only lane 0 has groups
            if (lane == 0) {
                if (c2 == 0)
                    ptx::bulk_wait_group_read<0>();
                else
                    ptx::bulk_wait_group_read<1>();
            }
            __syncwarp();
            const uint32_t sbuf = srow + static_cast<uint32_t>(c2 & 1) * 4096u;
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {  // 16-byte chunk j of the 128-byte row goes to slot j ^ (row & 7)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + ((j ^ sw) << 4)), "r"(pk[4 * j]),
                             "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                             : "memory");
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0 && row < M) {  // lane 0 holds the warp's first row
                ptx::tma_store_2d(&p.tmap_out, stage + (c2 & 1) * 4096, col0, row);
                ptx::bulk_commit_group();
            }
        }
        if (row_ok) {
            const int64_t o = static_cast<int64_t>(n_tile) * p.ld + row;
            p.part_m[o] = m2;
            p.part_s[o] = s;
            p.part_t[o] = tt;
        }
    }
};

// Backward LM-head epilogue: recompute the logits tile and turn it into d(raw logits) in bf16:
//   u = acc*scale_log2 - lse2[row]         (log2 p)
//   p = 2^u
//   d = inv_T * ( g_lp * 1[j == label] - p * (g_lp + g_ent * (u*ln2 + H)) )
struct EpiDLogits {
    static constexpr uint32_t kSmemBytes = 0;
    __device__ static void finish() {}
    struct Params {
        float scale_log2;
        float inv_temperature;
        const float* lse2;     // [M] log2-sum-exp of y
        const float* g_lp;     // [M]
        const float* g_ent;    // [M] or nullptr
        const float* entropy;  // [M] (needed iff g_ent)
        const int64_t* labels;
        __nv_bfloat16* dS;  // [M, ld]
        int64_t ld;
    };
    template <int BN>
    __device__ static void run(const Params& p, uint32_t taddr, int row, int n_tile, int M, int N, uint8_t*) {
        const int n0 = n_tile * BN;
        const bool row_ok = row < M;
        float lse2 = 0.f, glp = 0.f, gent = 0.f, H = 0.f;
        int lab = -1;
        if (row_ok) {
            lse2 = p.lse2[row];
            glp = p.g_lp[row];
            lab = static_cast<int>(p.labels[row]);
            if (p.g_ent != nullptr) {
                gent = p.g_ent[row];
                H = p.entropy[row];
            }
        }
        const float a0 = (glp + gent * H) * (-p.inv_temperature);
        const float b0 = gent * 0.6931471805599453f * (-p.inv_temperature);
        const float lab_add = glp * p.inv_temperature;
        const bool all_zero = (glp == 0.f && gent == 0.f);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
            ptx::tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (col0 >= N || !row_ok) continue;  // lanes reconverge at the __syncwarp above
            float d[32];
            if (all_zero) {
#pragma unroll
                for (int i = 0; i < 32; ++i) d[i] = 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float u = fmaf(__uint_as_float(v[i]), p.scale_log2, -lse2);
                    const float pr = ex2_approx(u);
                    d[i] = pr * fmaf(u, b0, a0);
                }
                if (lab >= col0 && lab < col0 + 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i == lab) d[i] += lab_add;
                }
            }
            __nv_bfloat16* dst = p.dS + static_cast<int64_t>(row) * p.ld + col0;
            if (col0 + 32 <= N && (p.ld & 15) == 0 && (reinterpret_cast<uintptr_t>(p.dS) & 31u) == 0) {
                store_row32_bf16(dst, d);
            } else {
                for (int i = 0; i < 32 && col0 + i < N; ++i) dst[i] = __float2bfloat16_rn(d[i]);
            }
        }
    }
};

}  // namespace skyb200
