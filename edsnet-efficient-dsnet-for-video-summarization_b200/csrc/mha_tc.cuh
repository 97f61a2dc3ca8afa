// Full multi-head attention base (modules/models.py:12-74; BASELINE.json config 4) on tcgen05 tensor cores:
// softmax(Q K^T / sqrt(128)) V per head, d_k = 128, 8 heads, eval mode -- flash style, the (T x T) score matrix the
// reference materialises (134 MB at T = 2048) never exists.  Same operand format and the same three split-fp16 passes
// (fp32-grade) as the Nystrom attention kernels of attn_tc.cuh:
//   mha_planes_kernel   fp32 Q | K | V [R][3072] (output of the projection GEMM) -> row-scaled fp16 hi / lo planes with one
//                       power-of-two scale per (row, part, head) = per 128 columns; q's scale carries the 1 / sqrt(d_k)
//   mha_tc_kernel       one CTA per (128-row query tile of one video, head); keys stream in 64-row tiles:
//                       S = Q K^T (M 128, N 64, K 128) and O_tile = P V (M 128, N 128, K 64) on tcgen05, TMA-fed, the
//                       running max / sum / output row in registers (two threads per query row)
#pragma once
#include "attn_tc.cuh"
#include "mha.cuh"

namespace tc {

// one warp per row: 24 slots (part x head) of 128 columns, lane <-> 4 consecutive columns of a slot
__global__ void __launch_bounds__(256)
mha_planes_kernel(const float* __restrict__ qkv, __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv,
                  int rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* src = qkv + (size_t)row * kMhaQkvCols;
    const float qmul = 0.08838834764831845f;                               // 1 / sqrt(128)
#pragma unroll 4
    for (int slot = 0; slot < 24; ++slot) {
        const float4 x = ldg4(src + slot * 128 + lane * 4);
        const float mx = warp_max(fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
        const int e = scale_exp(mx);
        const float sc = ldexpf(1.f, e);
        if (lane == 0) inv[(size_t)row * 24 + slot] = ldexpf(1.f, -e) * (slot < 8 ? qmul : 1.f);
        const float v[4] = {x.x * sc, x.y * sc, x.z * sc, x.w * sc};
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2half_rn(v[j]);
            l[j] = __float2half_rn(v[j] - __half2float(h[j]));
        }
        __half2 hh[2] = {__halves2half2(h[0], h[1]), __halves2half2(h[2], h[3])};
        __half2 ll[2] = {__halves2half2(l[0], l[1]), __halves2half2(l[2], l[3])};
        const size_t o = (size_t)row * kMhaQkvCols + slot * 128 + lane * 4;
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(hh);
        *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(ll);
    }
}

// shared memory: Q [2 K blocks][hi | lo][128 rows][128 B] 64 KB | K ring 2 x ([2 K blocks][hi | lo][64 keys][128 B]) 64 KB |
// V [hi | lo][2 column halves][64 keys][128 B] 32 KB | P [hi | lo][128 rows][128 B] 32 KB | scales, barriers
constexpr int kMtQ = 65536, kMtKStage = 32768, kMtV = 32768, kMtP = 32768;
constexpr int kMtVecBytes = 2 * 2 * 64 * 4 + 2 * 128 * 4 + 32;              // key scales [2 stages][k | v][64], pair exchange, v maxima
constexpr int kMtSmemBytes = kMtQ + 2 * kMtKStage + kMtV + kMtP + kMtVecBytes + 128 + 1024;

// 320 threads: warps 0..7 = query rows (thread t and t + 128 share row t & 127: keys 0..31 | 32..63 of a score tile,
// output columns 0..63 | 64..127), warp 8 = TMA producer, warp 9 = MMA issuer.
// TMEM 512 columns: S main | S cross (64 + 64), O main | O cross (128 + 128).
__global__ void __launch_bounds__(320, 1)
mha_tc_kernel(const __grid_constant__ CUtensorMap mapq_hi, const __grid_constant__ CUtensorMap mapq_lo,
              const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
              const float* __restrict__ inv, const int* __restrict__ cu_rows, const int2* __restrict__ tiles,
              float* __restrict__ y) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    constexpr int oQ = 0, oK = kMtQ, oV = oK + 2 * kMtKStage, oP = oV + kMtV, oVec = oP + kMtP;
    float* sc_vec = reinterpret_cast<float*>(g + oVec);                     // [stage][k | v][64]
    float* s_pair = sc_vec + 2 * 2 * 64;                                    // [2 halves][128 rows]
    float* s_vmx = s_pair + 2 * 128;                                        // [stage][warp 2 | warp 3]
    // barriers: K full[2] +0, K empty[2] +16, S done +32, P.V done +40, V full +48, V empty +56, S read out +64,
    // P in place +72, Q full +80; TMEM slot +96
    const uint32_t bars = base + oVec + kMtVecBytes;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + oVec + kMtVecBytes + 96);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y;
    const int2 tile = tiles[blockIdx.x];
    const VidInfo vi = vid_info(cu_rows, tile.x);
    const int q0 = tile.y;
    const int n_tiles = (vi.T + 63) / 64;

    if (tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 8, 1);                         // K full
        mbar_init(bars + 16, 1); mbar_init(bars + 24, 1);                   // K empty
        mbar_init(bars + 32, 1);                                            // S product done
        mbar_init(bars + 40, 1);                                            // P.V product done
        mbar_init(bars + 48, 1);                                            // V full
        mbar_init(bars + 56, 1);                                            // V empty
        mbar_init(bars + 64, 256);                                          // every row thread has read S
        mbar_init(bars + 72, 256);                                          // every row thread has stored P
        mbar_init(bars + 80, 1);                                            // Q tile has landed
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(bars + 96, 512);
    if (tid < 128) {
        // scales of key tile 0: thread t < 64 -> k scale of key t, 64 <= t < 128 -> v scale of key t - 64
        const int key = tid & 63, part = 1 + (tid >> 6);
        const float sc = key < vi.T ? __ldg(inv + (size_t)(vi.row0 + key) * 24 + part * 8 + h) : 0.f;
        sc_vec[(tid >> 6) * 64 + key] = sc;
        if (tid >= 64) {
            const float m = warp_max(sc);
            if (lane == 0) s_vmx[warp - 2] = m;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 8) {
        if (lane == 0) {
            // ---- TMA producer ----
            mbar_expect_tx(bars + 80, kMtQ);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                const int col = h * kMhaDk + kb * 64;
                tma_load_2d(base + oQ + kb * 32768, &mapq_hi, bars + 80, col, vi.row0 + q0);
                tma_load_2d(base + oQ + kb * 32768 + 16384, &mapq_lo, bars + 80, col, vi.row0 + q0);
            }
            bool pok = true;
            for (int i = 0; i < n_tiles && pok; ++i) {
                const int s = i & 1, row = vi.row0 + i * 64;
                pok = mbar_wait(bars + 16 + 8 * s, ((uint32_t)(i >> 1) & 1u) ^ 1u);
                if (!pok) break;
                const uint32_t kst = base + oK + s * kMtKStage;
                mbar_expect_tx(bars + 8 * s, kMtKStage);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const int col = kMhaFeat + h * kMhaDk + kb * 64;
                    tma_load_2d(kst + kb * 16384, &map_hi, bars + 8 * s, col, row);
                    tma_load_2d(kst + kb * 16384 + 8192, &map_lo, bars + 8 * s, col, row);
                }
                pok = mbar_wait(bars + 56, ((uint32_t)i & 1u) ^ 1u);        // P.V(i-1) has read the V tile
                if (!pok) break;
                mbar_expect_tx(bars + 48, kMtV);
#pragma unroll
                for (int dh = 0; dh < 2; ++dh) {
                    const int col = 2 * kMhaFeat + h * kMhaDk + dh * 64;
                    tma_load_2d(base + oV + dh * 8192, &map_hi, bars + 48, col, row);
                    tma_load_2d(base + oV + 16384 + dh * 8192, &map_lo, bars + 48, col, row);
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ---- MMA issuer: S(i+1) as soon as S(i) has been read out, P.V(i) as soon as P(i) is in place ----
            bool mok = true;
            auto issue_s = [&](int i) {
                const int s1 = i & 1;
                mok = mbar_wait(bars + 8 * s1, (uint32_t)(i >> 1) & 1u) && mok;          // K of that tile has landed
                tc_fence_after();
                const uint32_t kst = base + oK + s1 * kMtKStage;
                // K's hi and lo planes of a 64-column block are contiguous (8 KB each) and the accumulators adjacent
                // (main | cross): Q_hi [K_hi | K_lo]^T is one N = 128 instruction, Q_lo K_hi^T follows into the cross half
                constexpr uint32_t idesc = make_idesc(128, 64), idesc2 = make_idesc(128, 128);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t a0 = base + oQ + kb * 32768 + k * 32, b0 = kst + kb * 16384 + k * 32;
                        const uint64_t dah = make_smem_desc<64>(a0), dal = make_smem_desc<64>(a0 + 16384);
                        const uint64_t dbh = make_smem_desc<64>(b0);
                        umma_f16(tmem_base, dah, dbh, idesc2, (kb | k) != 0 ? 1u : 0u);
                        umma_f16(tmem_base + 64u, dal, dbh, idesc, 1u);
                    }
                umma_commit(bars + 32);
                umma_commit(bars + 16 + 8 * s1);                            // K stage free once the S product has read it
            };
            if (n_tiles > 0) {
                mok = mbar_wait(bars + 80, 0u);
                issue_s(0);
            }
            for (int i = 0; i < n_tiles && mok; ++i) {
                if (i + 1 < n_tiles) {
                    mok = mbar_wait(bars + 64, (uint32_t)i & 1u) && mok;    // S(i) is in registers everywhere
                    issue_s(i + 1);
                }
                mok = mbar_wait(bars + 48, (uint32_t)i & 1u) && mok;        // V of this tile has landed
                mok = mbar_wait(bars + 72, (uint32_t)i & 1u) && mok;        // P(i) stored, O(i-1) read by everyone
                tc_fence_after();
                constexpr uint32_t idesc_o = make_idesc_bmn(128, 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t a0 = base + oP + k * 32, b0 = base + oV + k * 2048;
                    const uint64_t dah = make_smem_desc<64>(a0), dal = make_smem_desc<64>(a0 + 16384);
                    // V_hi (two MN atoms) and V_lo (two more) are four atoms 8 KB apart: P_hi [V_hi | V_lo] is one N = 256
                    // instruction over the adjacent (main | cross) output accumulators
                    const uint64_t dbh = make_smem_desc_mn2(b0);
                    umma_f16(tmem_base + 128u, dah, dbh, make_idesc_bmn(128, 256), k != 0 ? 1u : 0u);
                    umma_f16(tmem_base + 256u, dal, dbh, idesc_o, 1u);
                }
                umma_commit(bars + 40);
                umma_commit(bars + 56);                                     // V tile free after P.V
            }
        }
    } else {
        const int row = tid & 127, half = tid >> 7;
        const int qrow = q0 + row;
        const float inv_q = qrow < vi.T ? __ldg(inv + (size_t)(vi.row0 + qrow) * 24 + h) : 0.f;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + (uint32_t)(half * 32);
        const uint32_t tO = tmem_base + lane_addr + 128u + (uint32_t)(half * 64);
        float o[64];
#pragma unroll
        for (int d = 0; d < 64; ++d) o[d] = 0.f;
        float run_max = -INFINITY, run_sum = 0.f;
        uint32_t s_phase = 0, pv_phase = 0;
        float alpha_prev = 1.f, inv_p_prev = 1.f;
        bool ok = true;
        auto collect_pv = [&]() {
            ok = mbar_wait(bars + 40, pv_phase) && ok;
            pv_phase ^= 1u;
            tc_fence_after();
            float pv[64];
            tmem_read64_sum(tO, tO + 128u, pv);
#pragma unroll
            for (int d = 0; d < 64; ++d) o[d] = fmaf(o[d], alpha_prev, pv[d] * inv_p_prev);
        };
        for (int i = 0; i < n_tiles && ok; ++i) {
            const int s = i & 1;
            // next tile's scales (stored into the other stage's slot below)
            float nsc = 0.f;
            if (tid < 128) {
                const int key = (i + 1) * 64 + (tid & 63), part = 1 + (tid >> 6);
                if (key < vi.T) nsc = __ldg(inv + (size_t)(vi.row0 + key) * 24 + part * 8 + h);
            }
            ok = mbar_wait(bars + 32, s_phase) && ok;
            s_phase ^= 1u;
            tc_fence_after();
            float p[32];
            tmem_read32_sum(tS, tS + 64u, p);
            const float* isk = sc_vec + (s * 2 + 0) * 64 + half * 32;
            const float* isv = sc_vec + (s * 2 + 1) * 64 + half * 32;
            const int kvalid = vi.T - i * 64 - half * 32;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                p[j] = j < kvalid ? p[j] * (inv_q * isk[j]) : -INFINITY;
                mx = fmaxf(mx, p[j]);
            }
            const float vmx = fmaxf(s_vmx[s * 2], s_vmx[s * 2 + 1]);        // largest v scale of the tile
            s_pair[half * 128 + row] = mx;
            tc_fence_before();
            mbar_arrive(bars + 64);                                         // this thread holds its S(i) values
            named_bar_sync(1, 256);
            if (tid < 128) {
                sc_vec[(((i + 1) & 1) * 2 + (tid >> 6)) * 64 + (tid & 63)] = nsc;
                if (tid >= 64) {
                    const float m = warp_max(nsc);
                    if (lane == 0) s_vmx[((i + 1) & 1) * 2 + (warp - 2)] = m;
                }
            }
            const float new_max = fmaxf(run_max, fmaxf(mx, s_pair[(half ^ 1) * 128 + row]));
            const float alpha = expf(run_max - new_max);                    // exp(-inf) = 0 on the first tile
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e = expf(p[j] - new_max);                       // 0 for masked keys
                ps += e;
                p[j] = e * isv[j];                                          // v's per-key plane scale folded into P
            }
            run_sum = run_sum * alpha + ps;
            run_max = new_max;
            if (i > 0) collect_pv();                                        // P tile and O accumulator are about to be reused
            const int ep = scale_exp(vmx);
            store_row32(g + oP, g + oP + 16384, row, half * 4, p, ldexpf(1.f, ep));
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bars + 72);                                         // P(i) stored, O(i-1) read
            named_bar_sync(1, 256);                                         // scale slots / s_pair reusable
            alpha_prev = alpha;
            inv_p_prev = ldexpf(1.f, -ep);
        }
        if (n_tiles > 0 && ok) collect_pv();
        s_pair[half * 128 + row] = run_sum;
        named_bar_sync(1, 256);
        const float rs = 1.f / (run_sum + s_pair[(half ^ 1) * 128 + row]);
        if (ok && qrow < vi.T) {
            float* dst = y + (size_t)(vi.row0 + qrow) * kMhaFeat + h * kMhaDk + half * 64;
#pragma unroll
            for (int d = 0; d < 64; d += 4) st4(dst + d, make_float4(o[d] * rs, o[d + 1] * rs, o[d + 2] * rs, o[d + 3] * rs));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
