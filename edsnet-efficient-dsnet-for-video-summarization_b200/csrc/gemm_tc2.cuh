// CTA-pair (cta_group::2) version of the three-pass split-fp16 tcgen05 GEMM of gemm_tc.cuh.
//
// Why: the one-CTA kernel streams A_hi, A_lo, B_hi, B_lo tiles (64 KB per 64-wide K block per 128 x 128 output tile)
// from L2 and is bound by exactly that feed (measured 11.7 TB/s of L2 -> SM traffic at 1.12 PFLOP/s of MMA work, the
// LTS ceiling of the chip), not by the tensor pipe.  Two CTAs of a cluster that work on ONE 256 x 128 tile need each
// other's B rows: every CTA loads its own 128 rows of A but only HALF of the B tile (64 rows), and tcgen05.mma
// .cta_group::2 (M = 256) reads the two halves from both shared memories.  48 KB instead of 64 KB per CTA and K block
// = 0.75 x the L2 traffic for the same flops.
//
// Roles per CTA (320 threads): warp 0 = TMA producer (own A rows, own half of B; completes on the LEADER's full
// barrier), warp 1 = MMA issuer (leader CTA only; commits multicast to both CTAs' empty / tmem-full barriers),
// warps 2..9 = epilogue (drain the CTA's own 128 accumulator lanes, then arrive on the leader's tmem-empty barrier).
// Accumulators: the same four-accumulator scheme as gemm_tc.cuh (512 TMEM columns per CTA, allocated with
// tcgen05.alloc.cta_group::2).
#pragma once
#include "gemm_tc.cuh"

namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
}

template <int STAGES>
struct Tc2Cfg {
    static constexpr int BM = 128, BN = 128, BK = 64;      // per-CTA output tile; the pair covers 256 x 128
    static constexpr int kATile = BM * BK * 2;             // 16 KB: one plane of this CTA's A rows
    static constexpr int kBHalf = (BN / 2) * BK * 2;       // 8 KB: one plane of this CTA's half of the B tile
    static constexpr int kStageBytes = 2 * (kATile + kBHalf);          // 48 KB per CTA
    static constexpr int kTmemCols = 512;
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreads = 320;
    static constexpr int kScratchOff = STAGES * kStageBytes + 256;
    static constexpr int kSmemBytes = kScratchOff + kEpiWarps * kEpiScratchWarp + 1024;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                float* __restrict__ C, int M, int N, int K, GemmEpiArgs ep) {
    using Cfg = Tc2Cfg<STAGES>;
    constexpr int BN = Cfg::BN, BK = Cfg::BK;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // same offset in both CTAs of the pair
    const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // used in the leader only
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tempty_bar = bar_base + 8u * (2 * STAGES + 1);             // used in the leader only
    constexpr int kSlotOff = 8 * (2 * STAGES + 2);
    const uint32_t tmem_slot = bar_base + kSlotOff;
    unsigned char* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + STAGES * Cfg::kStageBytes + kSlotOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nk = K / BK;
    const int tiles_n = N / BN;
    const int n_tiles = tiles_n * ((M + 2 * Cfg::BM - 1) / (2 * Cfg::BM));    // 256 x 128 pair tiles

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 2 * Cfg::kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                    // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer (both CTAs) ----
            int it = 0;
            bool ok = true;
            for (int tile = pair; tile < n_tiles && ok; tile += n_pairs) {
                const int m0 = (tile / tiles_n) * (2 * Cfg::BM) + (int)rank * Cfg::BM;
                const int nb0 = (tile % tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    if (!mbar_wait(empty_bar(s), ph ^ 1u)) { ok = false; break; }
                    const uint32_t st = smem_base + s * Cfg::kStageBytes;
                    const uint32_t lead_full = mapa_u32(full_bar(s), 0);
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);     // both CTAs' bytes
                    // stage layout: A_hi | A_lo | B_hi half | B_lo half
                    tma_load_2d_pair(st, &mapA, lead_full, kb * BK, m0);
                    tma_load_2d_pair(st + Cfg::kATile, &mapA, lead_full, kb * BK, M + m0);
                    tma_load_2d_pair(st + 2 * Cfg::kATile, &mapB, lead_full, kb * BK, nb0);
                    tma_load_2d_pair(st + 2 * Cfg::kATile + Cfg::kBHalf, &mapB, lead_full, kb * BK, N + nb0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ---- MMA issuer (leader CTA): D[256 x 128] over both CTAs' TMEM ----
            constexpr uint32_t idesc = make_idesc(2 * Cfg::BM, BN);
            bool ok = true;
            int it = 0, t = 0;
            for (int tile = pair; tile < n_tiles && ok; tile += n_pairs, ++t) {
                ok = mbar_wait(tempty_bar, ((uint32_t)t & 1u) ^ 1u);      // both epilogues have drained TMEM
                tc_fence_after();
                for (int kb = 0; kb < nk && ok; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    ok = mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * Cfg::kStageBytes;
                    const uint32_t a_hi = st, a_lo = st + Cfg::kATile;
                    const uint32_t b_hi = st + 2 * Cfg::kATile, b_lo = b_hi + Cfg::kBHalf;
                    const uint32_t acc_main = tmem_base + (uint32_t)((kb % 3) * BN);
                    const uint32_t acc_lo = tmem_base + 3u * BN;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t koff = k * 32;
                        const uint64_t dah = make_smem_desc<BK>(a_hi + koff), dbh = make_smem_desc<BK>(b_hi + koff);
                        const uint64_t dal = make_smem_desc<BK>(a_lo + koff), dbl = make_smem_desc<BK>(b_lo + koff);
                        umma_f16_pair(acc_main, dah, dbh, idesc, (kb < 3 && k == 0) ? 0u : 1u);
                        umma_f16_pair(acc_lo, dah, dbl, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_f16_pair(acc_lo, dal, dbh, idesc, 1u);
                    }
                    umma_commit_pair(empty_bar(s));
                }
                umma_commit_pair(tfull_bar);
            }
        }
    } else {
        // ---- epilogue (both CTAs): this CTA's 128 rows of the pair tile ----
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int n_main = nk < 3 ? nk : 3;
        const uint32_t lead_tempty = mapa_u32(tempty_bar, 0);
        float* scratch = reinterpret_cast<float*>(gen_base + Cfg::kScratchOff + (warp - 2) * kEpiScratchWarp);
        bool ok = true;
        int t = 0;
        for (int tile = pair; tile < n_tiles && ok; tile += n_pairs, ++t) {
            const int m0 = (tile / tiles_n) * (2 * Cfg::BM) + (int)rank * Cfg::BM;
            const int n0 = (tile % tiles_n) * BN;
            ok = mbar_wait(tfull_bar, (uint32_t)t & 1u);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64);
            const int nc0 = n0 + chalf * 64;
            float v[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t r0[16], r1[16], r2[16], r3[16];
                tmem_ld16_nowait(t0 + (uint32_t)c0, r0);
                tmem_ld16_nowait(t0 + 3u * BN + (uint32_t)c0, r3);
                if (n_main > 1) tmem_ld16_nowait(t0 + 1u * BN + (uint32_t)c0, r1);
                if (n_main > 2) tmem_ld16_nowait(t0 + 2u * BN + (uint32_t)c0, r2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float a0 = __uint_as_float(r0[j]);
                    const float a01 = n_main > 1 ? __fadd_rn(a0, __uint_as_float(r1[j])) : a0;
                    const float a23 = n_main > 2 ? __fadd_rn(__uint_as_float(r2[j]), __uint_as_float(r3[j]))
                                                 : __uint_as_float(r3[j]);
                    v[c0 + j] = __fadd_rn(a01, a23);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty);
            if (EPI == EPI_QKV_PLANES) epi_store_planes(v, m0 + quarter * 32, lane, nc0, M, N, C, ep, scratch);
            else epi_store_f32<EPI>(v, m0 + quarter * 32, lane, nc0, M, N, C, ep, scratch);
        }
    }
    // neither CTA may leave (or free TMEM) while the other still reads its shared memory / signals its barriers
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
}

template <int STAGES, int EPI>
cudaError_t launch_pair(const __half* A16, const __half* B16, float* C, int M, int N, int K, GemmEpiArgs ep,
                        cudaStream_t st, std::string* msg) {
    using Cfg = Tc2Cfg<STAGES>;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A16, 2ull * M, K, Cfg::BK, Cfg::BM, msg)) return cudaErrorUnknown;
    if (!make_map(&mapB, B16, 2ull * N, K, Cfg::BK, Cfg::BN / 2, msg)) return cudaErrorUnknown;
    auto kern = gemm_tc2_kernel<STAGES, EPI>;
    static const char tag = 0;                      // one per template instantiation
    if (DeviceOnce once_{&tag}) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
    }
    const int n_tiles = (N / Cfg::BN) * ((M + 2 * Cfg::BM - 1) / (2 * Cfg::BM));
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(mapA, mapB, C, M, N, K, ep);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Two-pass CTA-pair kernel (round 2).  The one-CTA two-pass kernel moves 96 KB through an SM's shared memory per 64-wide K
// block (48 KB of TMA fill + 48 KB of operand reads) for 4 x 128 tensor cycles: at 128 B per cycle that is 768 cycles, and
// 715-760 is what to_qkv measures.  In a pair the N = 256 right operand [B_hi | B_lo] splits exactly along its planes:
// the leader holds the 128 rows of B_hi, its peer the 128 rows of B_lo, each CTA its own 128 rows of A_hi; ONE
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16: 128 cycles per SM, so the ~90-cycle issue interval that capped the
// three-instruction kernel above stays hidden) leaves every CTA with the same (main | cross) accumulator pair as the
// one-CTA kernel, so the epilogues are shared.  32 KB per CTA and K block (64 KB through shared memory), six stages,
// accumulator pair double-buffered in TMEM.
// ---------------------------------------------------------------------------------------------------------------
template <int STAGES>
struct Tc2pCfg {
    static constexpr int BM = 128, BN = 128, BK = 64;      // per-CTA output tile; the pair covers 256 x 128
    static constexpr int kATile = BM * BK * 2;             // 16 KB: A_hi rows of this CTA
    static constexpr int kBTile = BN * BK * 2;             // 16 KB: B_hi (leader) or B_lo (peer) rows of the n tile
    static constexpr int kStageBytes = kATile + kBTile;    // 32 KB per CTA
    static constexpr int kTmemCols = 512;                  // two (main | cross) pairs
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreads = 320;
    static constexpr int kScratchOff = STAGES * kStageBytes + 256;
    static constexpr int kSmemBytes = kScratchOff + kEpiWarps * kEpiScratchWarp + 1024;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm_tc2p_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 float* __restrict__ C, int M, int N, int K, GemmEpiArgs ep) {
    using Cfg = Tc2pCfg<STAGES>;
    constexpr int BN = Cfg::BN, BK = Cfg::BK;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // same offset in both CTAs of the pair
    const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // used in the leader only
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };   // used in the leader only
    constexpr int kSlotOff = 8 * (2 * STAGES + 4);
    const uint32_t tmem_slot = bar_base + kSlotOff;
    unsigned char* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + STAGES * Cfg::kStageBytes + kSlotOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nk = K / BK;
    const int tiles_n = N / BN;
    const int n_tiles = tiles_n * ((M + 2 * Cfg::BM - 1) / (2 * Cfg::BM));    // 256 x 128 pair tiles

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * Cfg::kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                    // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer (both CTAs): own A_hi rows; the leader B_hi, the peer B_lo ----
            int it = 0;
            bool ok = true;
            for (int tile = pair; tile < n_tiles && ok; tile += n_pairs) {
                const int m0 = (tile / tiles_n) * (2 * Cfg::BM) + (int)rank * Cfg::BM;
                const int nb0 = (tile % tiles_n) * BN + (int)rank * N;        // plane p of B starts at row p * N
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    if (!mbar_wait(empty_bar(s), ph ^ 1u)) { ok = false; break; }
                    const uint32_t st = smem_base + s * Cfg::kStageBytes;
                    const uint32_t lead_full = mapa_u32(full_bar(s), 0);
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);     // both CTAs' bytes
                    tma_load_2d_pair(st, &mapA, lead_full, kb * BK, m0);
                    tma_load_2d_pair(st + Cfg::kATile, &mapB, lead_full, kb * BK, nb0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ---- MMA issuer (leader CTA): D[256 x 256] = A_hi [B_hi | B_lo]^T over both CTAs' TMEM ----
            constexpr uint32_t idesc = make_idesc(2 * Cfg::BM, 2 * BN);
            bool ok = true;
            int it = 0, t = 0;
            for (int tile = pair; tile < n_tiles && ok; tile += n_pairs, ++t) {
                const int buf = t & 1;
                ok = mbar_wait(tempty_bar(buf), ((uint32_t)(t >> 1) & 1u) ^ 1u);      // both epilogues have drained it
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * 2 * BN);
                for (int kb = 0; kb < nk && ok; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    ok = mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = smem_base + s * Cfg::kStageBytes;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t koff = k * 32;
                        umma_f16_pair(acc, make_smem_desc<BK>(st + koff), make_smem_desc<BK>(st + Cfg::kATile + koff), idesc,
                                      (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_pair(empty_bar(s));
                }
                umma_commit_pair(tfull_bar(buf));
            }
        }
    } else {
        // ---- epilogue (both CTAs): this CTA's 128 rows of the pair tile ----
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        float* scratch = reinterpret_cast<float*>(gen_base + Cfg::kScratchOff + (warp - 2) * kEpiScratchWarp);
        bool ok = true;
        int t = 0;
        for (int tile = pair; tile < n_tiles && ok; tile += n_pairs, ++t) {
            const int buf = t & 1;
            const int m0 = (tile / tiles_n) * (2 * Cfg::BM) + (int)rank * Cfg::BM;
            const int n0 = (tile % tiles_n) * BN;
            ok = mbar_wait(tfull_bar(buf), (uint32_t)(t >> 1) & 1u);
            tc_fence_after();
            const uint32_t t0 = tmem_base + (uint32_t)(buf * 2 * BN) + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64);
            const int nc0 = n0 + chalf * 64;
            float v[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(t0 + (uint32_t)c0, r0);
                tmem_ld16_nowait(t0 + (uint32_t)BN + (uint32_t)c0, r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[c0 + j] = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty_bar(buf), 0));
            if (EPI == EPI_QKV_PLANES) epi_store_planes(v, m0 + quarter * 32, lane, nc0, M, N, C, ep, scratch);
            else epi_store_f32<EPI>(v, m0 + quarter * 32, lane, nc0, M, N, C, ep, scratch);
        }
    }
    // neither CTA may leave (or free TMEM) while the other still reads its shared memory / signals its barriers
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
}

template <int STAGES, int EPI>
cudaError_t launch_pair2(const __half* A16, const __half* B16, float* C, int M, int N, int K, GemmEpiArgs ep,
                         cudaStream_t st, std::string* msg) {
    using Cfg = Tc2pCfg<STAGES>;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A16, 2ull * M, K, Cfg::BK, Cfg::BM, msg)) return cudaErrorUnknown;
    if (!make_map(&mapB, B16, 2ull * N, K, Cfg::BK, Cfg::BN, msg)) return cudaErrorUnknown;
    auto kern = gemm_tc2p_kernel<STAGES, EPI>;
    static const char tag = 0;                      // one per template instantiation
    if (DeviceOnce once_{&tag}) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
    }
    const int n_tiles = (N / Cfg::BN) * ((M + 2 * Cfg::BM - 1) / (2 * Cfg::BM));
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(mapA, mapB, C, M, N, K, ep);
    return cudaGetLastError();
}

}  // namespace tc
