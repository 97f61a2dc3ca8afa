// Tensor-core (tcgen05, fp16 hi/lo split, fp32-grade) landmark attention (transformer/nystroformer.py:95-142).
//
// Input format: the to_qkv GEMM (gemm_tc.cuh, EPI_QKV_PLANES) leaves q | k | v as two fp16 planes [R][1536]
// (hi = fp16(x 2^s), lo = fp16(x 2^s - hi)) plus inverse scales inv[R][24] (slot = part * 8 + head; q's slot carries
// the 1/8 of nystroformer.py:91).  A 64-column head slice of a plane row is exactly one 128-byte swizzle row, so TMA
// drops K-major / MN-major UMMA operands straight into shared memory -- no conversion, no register staging.
//
//   landmarks_planes_kernel  segment means of q, k                              (HBM bound, CUDA cores)
//   a3v_tc_kernel            softmax_keys(q_land k^T) v, streamed over 64-key tiles, two heads per CTA
//   attn_out_tc_kernel       softmax(q k_land^T) W per (video, head), streamed over 128-row tiles
//   value_conv_kernel        + depth-wise 33-tap FIR of v, thread <-> column, register sliding window
//
// Thread <-> accumulator row everywhere, so every softmax is a register-only row reduction.  Right operands that are
// naturally [k][n] row-major (v, W) are passed MN-major: same bytes, different descriptor.  An MN-major operand must
// not be scaled along K, so v's per-key scale is folded into the probabilities and W uses one scale per matrix.
#pragma once
#include "fc_stack_tc.cuh"

namespace tc {

// ---------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int scale_exp(float mx) {
    int e = 0;
    if (mx > 0.f && mx < INFINITY) e = 14 - ilogbf(mx);
    return max(-100, min(100, e));
}
// byte offset of 16-byte chunk c (0..7) of row r in a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void store_row64(unsigned char* hi, unsigned char* lo, int r, const float (&v)[64], float sc) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        __half2 hh[4], ll[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float v0 = v[c * 8 + 2 * q] * sc, v1 = v[c * 8 + 2 * q + 1] * sc;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            hh[q] = __halves2half2(h0, h1);
            ll[q] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
        }
        const uint32_t off = sw128_off(r, c);
        *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(hh);
        *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(ll);
    }
}
__device__ __forceinline__ void load_row64(float (&v)[64], const float* __restrict__ src) {
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
        const float4 x = ldg4(src + j);
        v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
    }
}
__device__ __forceinline__ float absmax64(const float (&v)[64]) {
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) mx = fmaxf(mx, fabsf(v[j]));
    return mx;
}
// half a row (32 values, columns c0..c0+31 of the head's 64) -> 4 chunks of the hi and lo planes of row r
__device__ __forceinline__ void store_row32(unsigned char* hi, unsigned char* lo, int r, int chunk0, const float (&v)[32],
                                            float mul) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        __half2 hh[4], ll[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float v0 = v[c * 8 + 2 * q] * mul, v1 = v[c * 8 + 2 * q + 1] * mul;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            hh[q] = __halves2half2(h0, h1);
            ll[q] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
        }
        const uint32_t off = sw128_off(r, chunk0 + c);
        *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(hh);
        *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(ll);
    }
}
__device__ __forceinline__ float absmax32(const float (&v)[32]) {
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(v[j]));
    return mx;
}

// MN-major right operand stored as [k rows][128 B = 64 n values], SWIZZLE_128B: 8-row atoms of 1024 B along K
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (64ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_bmn(int M, int N) { return make_idesc(M, N) | (1u << 16); }

// sum of the hi.hi and cross-term accumulators for 32 columns of this thread's row
__device__ __forceinline__ void tmem_read32_sum(uint32_t t_main, uint32_t t_lo, float (&v)[32]) {
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
        uint32_t r0[16], r1[16];
        tmem_ld16_nowait(t_main + (uint32_t)c0, r0);
        tmem_ld16_nowait(t_lo + (uint32_t)c0, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[c0 + j] = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
    }
}
// sum of the hi.hi and cross-term accumulators for 16 columns of this thread's row
__device__ __forceinline__ void tmem_read16_sum(uint32_t t_main, uint32_t t_lo, float (&v)[16]) {
    uint32_t r0[16], r1[16];
    tmem_ld16_nowait(t_main, r0);
    tmem_ld16_nowait(t_lo, r1);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
}
// a quarter row (16 values, two 16-byte chunks chunk0, chunk0 + 1) -> the hi and lo planes of row r
__device__ __forceinline__ void store_row16(unsigned char* hi, unsigned char* lo, int r, int chunk0, const float (&v)[16],
                                            float mul) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        __half2 hh[4], ll[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float v0 = v[c * 8 + 2 * q] * mul, v1 = v[c * 8 + 2 * q + 1] * mul;
            const __half2 h = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(h);
            hh[q] = h;
            ll[q] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
        }
        const uint32_t off = sw128_off(r, chunk0 + c);
        *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(hh);
        *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(ll);
    }
}
// sum of the hi.hi and cross-term accumulators for 64 columns of this thread's row
__device__ __forceinline__ void tmem_read64_sum(uint32_t t_main, uint32_t t_lo, float (&v)[64]) {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r0[16], r1[16];
        tmem_ld16_nowait(t_main + (uint32_t)c0, r0);
        tmem_ld16_nowait(t_lo + (uint32_t)c0, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[c0 + j] = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
    }
}
__device__ __forceinline__ uint64_t make_smem_desc_mn2(uint32_t addr) {      // two MN atoms (64 columns each), 8 KB apart
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (512ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// three-pass product of [M=128][K=64] (K-major A planes) with a [N=64][K=64] K-major or [K=64][N=64] MN-major B.
// B's hi and lo planes lie 8 KB apart (b_lo == b_hi + 8192) and the accumulators side by side (acc_lo == acc_main + 64),
// so A_hi [B_hi | B_lo] is ONE N = 128 instruction whose halves land in (main | cross); A_lo B_hi follows into the cross
// half: 8 instead of 12 tcgen05.mma per product, same accumulation order per accumulator (bit-identical results).  The
// issuing thread needs ~90 cycles per instruction (clock64 probe in the pinv chain), and in these latency-bound kernels
// that issue time is on the critical path.
template <bool B_MN>
__device__ __forceinline__ void issue_split_mma64(uint32_t acc_main, uint32_t acc_lo, uint32_t a_hi, uint32_t a_lo,
                                                  uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc2 = B_MN ? make_idesc_bmn(128, 128) : make_idesc(128, 128);
    constexpr uint32_t idesc1 = B_MN ? make_idesc_bmn(128, 64) : make_idesc(128, 64);
    (void)acc_lo; (void)b_lo;                                   // fixed by the layout: acc_main + 64, b_hi + 8192
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t ka = (uint32_t)k * 32u, kb = B_MN ? (uint32_t)k * 2048u : (uint32_t)k * 32u;
        const uint64_t dah = make_smem_desc<64>(a_hi + ka), dal = make_smem_desc<64>(a_lo + ka);
        const uint64_t db2 = B_MN ? make_smem_desc_mn2(b_hi + kb) : make_smem_desc<64>(b_hi + kb);
        const uint64_t db1 = B_MN ? make_smem_desc_mn(b_hi + kb) : make_smem_desc<64>(b_hi + kb);
        umma_f16(acc_main, dah, db2, idesc2, k != 0 ? 1u : 0u);
        umma_f16(acc_main + 64u, dal, db1, idesc1, 1u);
    }
}

// Both heads of a pair in one instruction: the left operand stacks the two heads along M (rows 0..63 | 64..127), the four
// right-operand planes (head 0 hi, head 0 lo, head 1 hi, head 1 lo) lie 8 KB apart and form ONE N = 256 operand, and the
// accumulators (head 0 main | cross | head 1 main | cross) are 256 adjacent TMEM columns.  A_hi B and A_lo B: two
// instructions per K step, 8 per product pair instead of 16.  Each accumulator half receives the products of both rows
// blocks (only its own head's block is read), and all four terms hi.hi + hi.lo + lo.hi + lo.lo are summed (the lo.lo term,
// 2^-22 relative, comes for free).  The tensor core takes an instruction every ~90 cycles from one SM, whatever its size
// (clock64 probes: two issuing warps did not issue faster than one), so the instruction count is what paces these
// small-tile kernels.
template <bool B_MN>
__device__ __forceinline__ void issue_pair_mma64(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t b) {
    constexpr uint32_t idesc = B_MN ? make_idesc_bmn(128, 256) : make_idesc(128, 256);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t ka = (uint32_t)k * 32u, kb = B_MN ? (uint32_t)k * 2048u : (uint32_t)k * 32u;
        const uint64_t db = B_MN ? make_smem_desc_mn2(b + kb) : make_smem_desc<64>(b + kb);
        umma_f16(acc, make_smem_desc<64>(a_hi + ka), db, idesc, k != 0 ? 1u : 0u);
        umma_f16(acc, make_smem_desc<64>(a_lo + ka), db, idesc, 1u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// landmarks from the planes: q_land / k_land [V][8][64][64] fp32 = mean over each block of `seg` padded rows.
// grid (64 landmarks, V), 256 threads; thread owns 4 consecutive columns of the 1024 q|k columns.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
landmarks_planes_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, const float* __restrict__ inv,
                        const int* __restrict__ cu_rows, float* __restrict__ q_land, float* __restrict__ k_land,
                        unsigned* __restrict__ ticket) {
    const int v = blockIdx.y, j = blockIdx.x;
    // the work-item counter of the (persistent) a3v kernel that follows on the same stream starts at zero
    if (ticket != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *ticket = 0u;
    const VidInfo vi = vid_info(cu_rows, v);
    const int c4 = threadIdx.x * 4;
    const int slot = c4 >> 6;                                   // part * 8 + head, part in {q, k}
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int r_lo = j * vi.seg - vi.pad, r_hi = r_lo + vi.seg;
    if (r_lo < 0) r_lo = 0;
    for (int r = r_lo; r < r_hi; ++r) {
        const size_t row = (size_t)(vi.row0 + r);
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + row * kQkvCols + c4));
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(lo + row * kQkvCols + c4));
        const float s = __ldg(inv + row * 24 + slot);
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
        const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
        acc.x += (h0.x + l0.x) * s; acc.y += (h0.y + l0.y) * s;
        acc.z += (h1.x + l1.x) * s; acc.w += (h1.y + l1.y) * s;
    }
    const float div = (float)vi.seg;
    acc.x /= div; acc.y /= div; acc.z /= div; acc.w /= div;
    const int is_k = c4 >= kInner;
    const int cc = c4 & (kInner - 1);
    const int hd = cc >> 6, d = cc & 63;
    float* dst = (is_k ? k_land : q_land) + ((((size_t)v * kHeads + hd) * kLandmark + j) * kDimHead + d);
    st4(dst, acc);
}

// ---------------------------------------------------------------------------------------------------------
// a3v = softmax_over_keys(q_land k^T) v for TWO heads of one video per CTA (nystroformer.py:118,130,133).
// Accumulator row t: head (t / 64) of the pair, landmark t % 64.  Both heads' landmarks form one 128-row A operand;
// S_hh = A k_hh^T and O_hh = P v_hh are issued for both heads and every thread reads the product of ITS head (the other
// half of each product is never read).  Keys stream in 64-row tiles through a 2-stage TMA ring; the running
// max / sum / output row live in registers (flash-attention style), the zero pad keys of the reference are folded
// into the start state (logit 0, value 0).
// TMEM 512 columns: S0 S1 O0 O1, each main | cross (64 + 64).
// ---------------------------------------------------------------------------------------------------------
constexpr int kA3Stage = 8 * 8192;                                          // k_h0 k_h1 v_h0 v_h1, hi and lo
constexpr int kA3VecBytes = 2 * 4 * 64 * 4 + 4 * 128 * 4 + 32 + 128 * 4 + 64;   // key scales [2 stages][4][64] + row exchange [4][128]
                                                                            // + largest v scale [2 stages][2 heads][2 warps]
                                                                            // + landmark-key scales [128] + attn2 reductions
constexpr int kA3SmemBytes = 32768 + 2 * kA3Stage + 32768 + kA3VecBytes + 128 + 1024;

// 576 threads: warps 0..15 = rows -- FOUR threads per accumulator row (t, t + 128, t + 256, t + 384 share row t & 127 and
// own 16 keys / output columns each: clock64 probes showed the per-row softmax chain of two threads with 32 columns
// each, 3 900 cycles per key tile, as what paces this kernel, not the 16 MMA instructions), warp 16 = TMA producer,
// warp 17 = MMA issuer (told by mbarriers when S has been read out / P is in place, so the row warps never wait for an
// issue loop).  Both heads' products are one instruction pair per K step (issue_pair_mma64).
// gridDim.z > 1 (few, long videos: one video would otherwise keep 4 of 148 SMs busy): CTA z streams the z-th contiguous
// range of key tiles and leaves its un-normalised output rows and (running max, sum) in `part` [V][8][Z][64][66];
// a3v_merge_kernel combines the ranges (flash-decoding style).  The zero pad keys belong to range 0.
//
// attn2 != nullptr: the CTA of range 0 also produces attn2 = softmax(q_land k_land^T) of its two heads
// (nystroformer.py:117,130) and the two magnitudes the pseudo-inverse start value needs (:16-18) before it starts on the
// key tiles: the landmark queries are already in place as the A operand, the landmark keys go through the (still unused)
// P tile as a K-major B operand, one extra product per head into the S accumulators.  The key / value ring fills
// meanwhile.  This replaces attn2_kernel (a launch of its own with 64^3 FFMA products per head) on the tcgen05 path.
constexpr int kA3PartLd = 66;
constexpr int kA3Threads = 576, kA3RowThreads = 512;
__global__ void __launch_bounds__(kA3Threads, 1)
a3v_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
              const float* __restrict__ inv, const int* __restrict__ cu_rows, const float* __restrict__ q_land,
              float* __restrict__ a3v, float* __restrict__ part_out, const float* __restrict__ k_land,
              float* __restrict__ attn2, float* __restrict__ stats, unsigned* __restrict__ ticket, int n_videos,
              int zsplit) {
    // ticket != nullptr: PERSISTENT -- one CTA per SM draws work items (head pair, video, key range) from a device-side
    // counter (zeroed by the landmarks kernel in front of this launch) until they run out.  A CTA of this size costs
    // 3-4.7 us to launch on an SM that has just been vacated (globaltimer stamps) and one item is only ~27 us of work;
    // every item starts from freshly initialised barriers, so nothing else about an item changes.  ticket == nullptr:
    // one item per CTA, item = blockIdx.x.
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    constexpr int oQl = 0, oKV = 32768, oP = 32768 + 2 * kA3Stage, oVec = oP + 32768;
    float* sc_vec = reinterpret_cast<float*>(g + oVec);                     // [stage][k_h0 k_h1 v_h0 v_h1][64]
    float* s_pair = sc_vec + 2 * 4 * 64;                                    // [4 parts][128 rows]
    float* s_vmx = s_pair + 4 * 128;                                        // [stage][head][warp 2 | warp 3]
    float* s_ikl = s_vmx + 8;                                               // [128] inverse plane scales of the landmark keys
    float* s_a2red = s_ikl + 128;                                           // [2 kinds][2 heads][2 warps] attn2 reductions
    // barriers: K full[2] +0, K empty[2] +16, S done +32, P.V done +40, V full[2] +48, V empty[2] +64, S read out +80,
    // P in place +88; TMEM slot +96; attn2 product done +104, read out +112
    const uint32_t bars = base + oVec + kA3VecBytes;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + oVec + kA3VecBytes + 96);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_items = 4 * n_videos * zsplit;
    __shared__ int s_item;
    if (warp == 16) tmem_alloc(bars + 96, 512);
    bool first_item = true;
    int next_item = 0;                                                      // thread 0: ticket drawn one item ahead
    uint32_t tmem_base = 0;
    for (;;) {
    // the ticket of the NEXT item is drawn while this one streams (the atomic's round trip was 7 % of this kernel's stall
    // samples when every item began with it, ncu r02w); tickets past the end are harmless
    if (tid == 0) s_item = ticket != nullptr ? (first_item ? (int)atomicAdd(ticket, 1u) : next_item)
                                             : (first_item ? (int)blockIdx.x : n_items);
    __syncthreads();                                                        // (also: everyone has left the previous item)
    const int item = s_item;
    if (item >= n_items) break;
    if (tid == 0 && ticket != nullptr) next_item = (int)atomicAdd(ticket, 1u);
    const int pair = item & 3, v = (item >> 2) % n_videos, zi = (item >> 2) / n_videos;
    const VidInfo vi = vid_info(cu_rows, v);
    const int tiles_all = (vi.T + 63) / 64;
    const int per = (tiles_all + zsplit - 1) / zsplit;
    const int tile0 = zi * per;                                             // first key tile of this CTA's range
    const int n_tiles = max(0, min(per, tiles_all - tile0));
    const int h0 = pair * 2;
    const bool do_attn2 = attn2 != nullptr && zi == 0;

    if (tid == 0) {
        if (!first_item) {
            for (int q = 0; q < 12; ++q) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bars + 8 * q) : "memory");
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bars + 104) : "memory");
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bars + 112) : "memory");
        }
        mbar_init(bars + 104, 1);                                           // attn2 products done
        mbar_init(bars + 112, kA3RowThreads);                               // attn2 logits read out
        mbar_init(bars, 1); mbar_init(bars + 8, 1);                         // K full
        mbar_init(bars + 16, 1); mbar_init(bars + 24, 1);                   // K empty
        mbar_init(bars + 32, 1);                                            // S products done
        mbar_init(bars + 40, 1);                                            // P.V products done
        mbar_init(bars + 48, 1); mbar_init(bars + 56, 1);                   // V full
        mbar_init(bars + 64, 1); mbar_init(bars + 72, 1);                   // V empty
        mbar_init(bars + 80, kA3RowThreads);                                // every row thread has read S
        mbar_init(bars + 88, kA3RowThreads);                                // every row thread has stored P
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float inv_ql = 1.f;
    // Prologue (a clock64 probe put it at 8 000 cycles with one thread per 64-value row, a third of an average CTA's
    // life, and nothing overlaps it at one CTA per SM): two threads per row, lanes 2r and 2r + 1 of the same warp, the row
    // maximum through one shuffle; all global loads issued before the first dependent use.
    if (tid < 256) {
        // landmark queries of both heads -> A operand planes (row r), per-row scale
        const int r = tid >> 1, hf = tid & 1;
        float q[32];
        const float* src = q_land + (((size_t)v * kHeads + h0 + (r >> 6)) * kLandmark + (r & 63)) * kDimHead + hf * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 x = ldg4(src + j);
            q[j] = x.x; q[j + 1] = x.y; q[j + 2] = x.z; q[j + 3] = x.w;
        }
        // scales of the first tile (threads 0..127): thread t < 64 -> key t: k scales of both heads; 64 <= t < 128 -> v scales
        float sc0 = 0.f, sc1 = 0.f;
        if (tid < 128) {
            const int key = tile0 * 64 + (tid & 63), qkv_part = 1 + (tid >> 6);
            if (key < vi.T) {
                const float* ip = inv + (size_t)(vi.row0 + key) * 24 + qkv_part * 8 + h0;
                sc0 = __ldg(ip); sc1 = __ldg(ip + 1);
            }
        }
        float mx = absmax32(q);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        const int e = scale_exp(mx);
        if (hf == 0) s_pair[r] = ldexpf(1.f, -e);
        store_row32(g + oQl, g + oQl + 16384, r, hf * 4, q, ldexpf(1.f, e));
        if (tid < 128) {
            sc_vec[((tid >> 6) * 2 + 0) * 64 + (tid & 63)] = sc0;
            sc_vec[((tid >> 6) * 2 + 1) * 64 + (tid & 63)] = sc1;
            if (tid >= 64) {                               // warps 2, 3 hold the v scales: their maxima per head
                const float m0 = warp_max(sc0), m1 = warp_max(sc1);
                if (lane == 0) { s_vmx[0 * 2 + (warp - 2)] = m0; s_vmx[1 * 2 + (warp - 2)] = m1; }
            }
        }
    } else if (tid < kA3RowThreads && do_attn2) {
        // landmark keys of both heads -> K-major B operand planes in the P tile: head hh at hh * 16 KB, hi then lo
        const int t = tid - 256, r = t >> 1, hf = t & 1, hh = r >> 6, j = r & 63;
        float kl[32];
        const float* src = k_land + (((size_t)v * kHeads + h0 + hh) * kLandmark + j) * kDimHead + hf * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            const float4 x = ldg4(src + c);
            kl[c] = x.x; kl[c + 1] = x.y; kl[c + 2] = x.z; kl[c + 3] = x.w;
        }
        float mx = absmax32(kl);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        const int e = scale_exp(mx);
        if (hf == 0) s_ikl[r] = ldexpf(1.f, -e);
        store_row32(g + oP + hh * 16384, g + oP + hh * 16384 + 8192, j, hf * 4, kl, ldexpf(1.f, e));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_base = *tmem_slot_ptr;

    if (warp == 16) {
        if (lane == 0) {
            // ---- TMA producer: 8 boxes of 64 rows x 64 columns per tile.  The K half of a stage is released by the
            // S products, the V half by the P.V products, so the K tiles run one tile further ahead than the V tiles ----
            bool pok = true;
            for (int i = 0; i < n_tiles && pok; ++i) {
                const int s = i & 1;
                const uint32_t ph = ((uint32_t)(i >> 1) & 1u) ^ 1u;
                const uint32_t st = base + oKV + s * kA3Stage;
                const int row = vi.row0 + (tile0 + i) * 64;
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {                            // 0: k_h0 k_h1, 1: v_h0 v_h1
                    const uint32_t full = bars + (kv ? 48 : 0) + 8 * s, empty = bars + (kv ? 64 : 16) + 8 * s;
                    if (!pok || !mbar_wait(empty, ph)) { pok = false; continue; }
                    mbar_expect_tx(full, kA3Stage / 2);
#pragma unroll
                    for (int b = 2 * kv; b < 2 * kv + 2; ++b) {
                        const int col = (1 + (b >> 1)) * kInner + (h0 + (b & 1)) * kDimHead;
                        tma_load_2d(st + b * 16384, &map_hi, full, col, row);
                        tma_load_2d(st + b * 16384 + 8192, &map_lo, full, col, row);
                    }
                }
            }
        }
    } else if (warp == 17) {
        if (lane == 0) {
            // ---- MMA issuer: S(i+1) as soon as S(i) has been read out, P.V(i) as soon as P(i) is in place ----
            bool mok = true;
            if (do_attn2) {
                tc_fence_after();
                issue_pair_mma64<false>(tmem_base, base + oQl, base + oQl + 16384, base + oP);
                umma_commit(bars + 104);
                mok = mbar_wait(bars + 112, 0u);                            // logits read out: the S accumulators are free
            }
            auto issue_s = [&](int i) {
                const int s1 = i & 1;
                mok = mbar_wait(bars + 8 * s1, (uint32_t)(i >> 1) & 1u) && mok;         // K of that tile has landed
                tc_fence_after();
                issue_pair_mma64<false>(tmem_base, base + oQl, base + oQl + 16384, base + oKV + s1 * kA3Stage);
                umma_commit(bars + 32);
                umma_commit(bars + 16 + 8 * s1);                            // K half of the stage free after the S products
            };
            if (n_tiles > 0) issue_s(0);
            for (int i = 0; i < n_tiles && mok; ++i) {
                const int s = i & 1;
                const uint32_t st = base + oKV + s * kA3Stage;
                if (i + 1 < n_tiles) {
                    mok = mbar_wait(bars + 80, (uint32_t)i & 1u) && mok;    // S(i) is in registers everywhere
                    issue_s(i + 1);
                }
                mok = mbar_wait(bars + 48 + 8 * s, (uint32_t)(i >> 1) & 1u) && mok;     // V of this tile has landed
                mok = mbar_wait(bars + 88, (uint32_t)i & 1u) && mok;        // P(i) stored, O(i-1) read by everyone
                tc_fence_after();
                issue_pair_mma64<true>(tmem_base + 256u, base + oP, base + oP + 16384, st + 32768);
                umma_commit(bars + 40);
                umma_commit(bars + 64 + 8 * s);                             // V half of the stage free after P.V
            }
        }
    } else {
        constexpr int NR = kA3RowThreads;
        const int row = tid & 127, part = tid >> 7;                         // part p owns keys / columns 16 p .. 16 p + 15
        const int hh = row >> 6;                                            // which head of the pair this row belongs to
        inv_ql = s_pair[row];
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + (uint32_t)(hh * 128 + part * 16), tO = tS + 256u;
        float o[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = 0.f;
        // the maximum of a value over the four threads of a row (exchange through s_pair; ends with everyone past the reads)
        auto row_max4 = [&](float x) -> float {
            s_pair[part * 128 + row] = x;
            named_bar_sync(1, NR);
            return fmaxf(fmaxf(s_pair[row], s_pair[128 + row]), fmaxf(s_pair[256 + row], s_pair[384 + row]));
        };
        auto row_sum4 = [&](float x) -> float {
            s_pair[part * 128 + row] = x;
            named_bar_sync(1, NR);
            return (s_pair[row] + s_pair[128 + row]) + (s_pair[256 + row] + s_pair[384 + row]);
        };
        // running max is shared by the four threads of a row; the running sum is per thread (its keys) and merged at the end
        const bool pads_here = vi.pad > 0 && zi == 0;                       // the zero pad keys: logit 0, value 0
        float run_max = pads_here ? 0.f : -INFINITY, run_sum = (part == 0 && pads_here) ? (float)vi.pad : 0.f;
        // Software pipeline over the key tiles: S(i+1) is issued as soon as every thread has pulled S(i) out of TMEM,
        // so it runs under the softmax of tile i; the P.V product of tile i is only collected in iteration i+1, after
        // that tile's softmax arithmetic, so it runs under the S read-out and the exponentials of tile i+1.
        uint32_t s_phase = 0, pv_phase = 0;
        float alpha_prev = 1.f, inv_p_prev = 1.f;
        bool ok = true;
        named_bar_sync(1, NR);                                              // everyone has read inv_ql out of s_pair
        if (do_attn2) {
            // ---- attn2 rows of this thread's head: softmax over the 64 landmark keys (16 per thread of the row) ----
            bool aok = mbar_wait(bars + 104, 0u);
            tc_fence_after();
            float p[16];
            tmem_read16_sum(tS, tS + 64u, p);
            tc_fence_before();
            mbar_arrive(bars + 112);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) { p[j] *= inv_ql * s_ikl[hh * 64 + part * 16 + j]; mx = fmaxf(mx, p[j]); }
            mx = row_max4(mx);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
            named_bar_sync(1, NR);                                          // maxima consumed: s_pair carries the sums now
            sum = row_sum4(sum);
            // the landmark-key planes are dead (the product has completed): the P tile holds the probabilities as fp32
            // [128 rows][64] for the column sums
            float* a2s = reinterpret_cast<float*>(g + oP);
            float rsum = 0.f;
            float* dst = attn2 + (((size_t)v * kHeads + h0 + hh) * kLandmark + (row & 63)) * kLandmark + part * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 q4 = make_float4(p[j] / sum, p[j + 1] / sum, p[j + 2] / sum, p[j + 3] / sum);   // (exact divisions: attn2 feeds the ill-conditioned chain)
                rsum += (q4.x + q4.y) + (q4.z + q4.w);
                st4(dst + j, q4);
                st4(a2s + row * 64 + part * 16 + j, q4);
            }
            named_bar_sync(1, NR);                                          // (also: everyone is done with the sums in s_pair)
            rsum = row_sum4(rsum);
            // column sums: thread (part 0, row) <-> column (row & 63) of head hh, the 64 rows in order
            float csum = 0.f;
            if (part == 0) {
                const float* col = a2s + hh * 64 * 64 + (row & 63);
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;                // four chains, combined in a fixed order
#pragma unroll 4
                for (int i = 0; i < 64; i += 4) {
                    c0 += col[i * 64]; c1 += col[(i + 1) * 64]; c2 += col[(i + 2) * 64]; c3 += col[(i + 3) * 64];
                }
                csum = (c0 + c1) + (c2 + c3);
                const float rmax = warp_max(rsum), cmax = warp_max(csum);
                if (lane == 0) { s_a2red[(0 * 2 + hh) * 2 + (warp & 1)] = rmax; s_a2red[(1 * 2 + hh) * 2 + (warp & 1)] = cmax; }
            }
            named_bar_sync(1, NR);
            if (aok && part == 0 && (row & 63) == 0) {
                float* sp = stats + ((size_t)v * kHeads + h0 + hh) * 2;
                sp[0] = fmaxf(s_a2red[(0 * 2 + hh) * 2], s_a2red[(0 * 2 + hh) * 2 + 1]);
                sp[1] = fmaxf(s_a2red[(1 * 2 + hh) * 2], s_a2red[(1 * 2 + hh) * 2 + 1]);
            }
            // (the first P store of the key loop comes after two more barriers of this group: the column sums are done)
        }
        auto collect_pv = [&]() {
            ok = mbar_wait(bars + 40, pv_phase) && ok;
            pv_phase ^= 1u;
            tc_fence_after();
            float pv[16];
            tmem_read16_sum(tO, tO + 64u, pv);
#pragma unroll
            for (int d = 0; d < 16; ++d) o[d] = fmaf(o[d], alpha_prev, pv[d] * inv_p_prev);
        };
        for (int i = 0; i < n_tiles && ok; ++i) {
            const int s = i & 1;
            // prefetch the next tile's scales (written to the other stage's slot below)
            float nsc0 = 0.f, nsc1 = 0.f;
            if (tid < 128) {
                const int key = (tile0 + i + 1) * 64 + (tid & 63), qkv_part = 1 + (tid >> 6);
                if (key < vi.T) {
                    const float* ip = inv + (size_t)(vi.row0 + key) * 24 + qkv_part * 8 + h0;
                    nsc0 = __ldg(ip); nsc1 = __ldg(ip + 1);
                }
            }
            ok = mbar_wait(bars + 32, s_phase) && ok;
            s_phase ^= 1u;
            tc_fence_after();
            // ---- this thread's 16 logits of its row ----
            float p[16];
            tmem_read16_sum(tS, tS + 64u, p);
            const float* isk = sc_vec + (s * 4 + hh) * 64 + part * 16;
            const float* isv = sc_vec + (s * 4 + 2 + hh) * 64 + part * 16;
            const int kvalid = vi.T - (tile0 + i) * 64 - part * 16;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                p[j] = j < kvalid ? p[j] * (inv_ql * isk[j]) : -INFINITY;
                mx = fmaxf(mx, p[j]);
            }
            const float vmx = fmaxf(s_vmx[(s * 2 + hh) * 2], s_vmx[(s * 2 + hh) * 2 + 1]);   // largest v scale of the tile
            tc_fence_before();
            mbar_arrive(bars + 80);                                         // this thread holds its S(i) values
            mx = row_max4(mx);
            // next tile's scales into the other stage's slot: its previous readers (tile i-1) are past their last barrier
            if (tid < 128) {
                sc_vec[(((i + 1) & 1) * 4 + (tid >> 6) * 2 + 0) * 64 + (tid & 63)] = nsc0;
                sc_vec[(((i + 1) & 1) * 4 + (tid >> 6) * 2 + 1) * 64 + (tid & 63)] = nsc1;
                if (tid >= 64) {
                    const float m0 = warp_max(nsc0), m1 = warp_max(nsc1);
                    if (lane == 0) {
                        s_vmx[(((i + 1) & 1) * 2 + 0) * 2 + (warp - 2)] = m0;
                        s_vmx[(((i + 1) & 1) * 2 + 1) * 2 + (warp - 2)] = m1;
                    }
                }
            }
            const float new_max = fmaxf(run_max, mx);
            const float alpha = expf(run_max - new_max);                    // exp(-inf) = 0 on a fresh start
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float e = expf(p[j] - new_max);                       // 0 for masked keys
                ps += e;
                p[j] = e * isv[j];                                          // fold v's per-key scale into P
            }
            run_sum = run_sum * alpha + ps;
            run_max = new_max;
            // the previous tile's P.V product: its P tile and O accumulator are about to be reused
            if (i > 0) collect_pv();
            // P' <= max_j isv[j]: one row scale for all four parts without another exchange
            const int ep = scale_exp(vmx);
            store_row16(g + oP, g + oP + 16384, row, part * 2, p, ldexpf(1.f, ep));
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bars + 88);                                         // P(i) stored, O(i-1) read
            named_bar_sync(1, NR);                                          // scale slots / s_pair reusable
            alpha_prev = alpha;
            inv_p_prev = ldexpf(1.f, -ep);
        }
        if (n_tiles > 0 && ok) collect_pv();
        // merge the four partial sums of the row
        const float tot = row_sum4(run_sum);
        if (ok && zsplit == 1) {
            const float rs = 1.f / tot;
            float* dst = a3v + (((size_t)v * kHeads + h0 + hh) * kLandmark + (row & 63)) * kDimHead + part * 16;
#pragma unroll
            for (int d = 0; d < 16; d += 4) st4(dst + d, make_float4(o[d] * rs, o[d + 1] * rs, o[d + 2] * rs, o[d + 3] * rs));
        } else if (ok) {
            float* dst = part_out + ((((size_t)v * kHeads + h0 + hh) * zsplit + zi) * kLandmark + (row & 63)) * kA3PartLd;
#pragma unroll
            for (int d = 0; d < 16; d += 2) *reinterpret_cast<float2*>(dst + part * 16 + d) = make_float2(o[d], o[d + 1]);
            if (part == 0) *reinterpret_cast<float2*>(dst + 64) = make_float2(run_max, tot);
        }
    }
    tc_fence_before();
    first_item = false;
    }   // work items
    tc_fence_before();
    __syncthreads();
    if (warp == 16 && !first_item) tmem_dealloc(tmem_base, 512);
    if (warp == 16 && first_item) {                                        // no item at all: the allocation is still ours
        tc_fence_after();
        tmem_dealloc(*tmem_slot_ptr, 512);
    }
}

// Combine the key ranges of a3v_tc_kernel: a3v[j][d] = sum_z o_z[j][d] e^(m_z - m) / sum_z l_z e^(m_z - m), m = max_z m_z
// (an empty range leaves m_z = -inf, l_z = 0).  grid (8, V), 256 threads: thread <-> (landmark row, 16 columns).
__global__ void __launch_bounds__(256)
a3v_merge_kernel(const float* __restrict__ part, int zsplit, float* __restrict__ a3v) {
    const int h = blockIdx.x, v = blockIdx.y;
    const int j = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16;
    const float* p0 = part + (((size_t)v * kHeads + h) * zsplit) * kLandmark * kA3PartLd + (size_t)j * kA3PartLd;
    float m = -INFINITY;
    for (int z = 0; z < zsplit; ++z) m = fmaxf(m, p0[(size_t)z * kLandmark * kA3PartLd + 64]);
    float acc[16], den = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] = 0.f;
    for (int z = 0; z < zsplit; ++z) {
        const float* pz = p0 + (size_t)z * kLandmark * kA3PartLd;
        const float w = expf(pz[64] - m);                                    // exp(-inf) = 0 for an empty range
        den = fmaf(pz[65], w, den);
#pragma unroll
        for (int d = 0; d < 16; d += 2) {
            const float2 o = *reinterpret_cast<const float2*>(pz + c0 + d);
            acc[d] = fmaf(o.x, w, acc[d]);
            acc[d + 1] = fmaf(o.y, w, acc[d + 1]);
        }
    }
    const float rs = 1.f / den;
    float* dst = a3v + (((size_t)v * kHeads + h) * kLandmark + j) * kDimHead + c0;
#pragma unroll
    for (int d = 0; d < 16; d += 4) st4(dst + d, make_float4(acc[d] * rs, acc[d + 1] * rs, acc[d + 2] * rs, acc[d + 3] * rs));
}

// ---------------------------------------------------------------------------------------------------------
// attn_out = softmax(q k_land^T) W for one (video, head) per CTA (nystroformer.py:115,130,133), q streamed in 128-row
// tiles through a 2-stage TMA ring.  k_land and W are converted to operand planes once per CTA.  The probabilities
// overwrite the q stage they were computed from (q is dead once S is in TMEM), so a CTA needs 96 KB: two per SM.
// Writes attn[R][512] (head-merged columns h*64..); value_conv_kernel adds the convolution afterwards.
// TMEM 256 columns: S main | S cross | O main | O cross.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAoVecBytes = 64 * 4 + 4 * 128 * 4;                            // inv_kl [64] + pair exchange max / sum [2][128] each
constexpr int kAoScratch = 0;                                               // the read-out borrows the dead q stage (collect)
constexpr int kAoSmemBytes = 2 * 8192 + 2 * 8192 + 2 * 32768 + kAoVecBytes + 128 + kAoScratch + 1024;

// 320 threads: warps 0..7 = rows (thread t and t + 128 share accumulator row t & 127: landmarks / output columns
// 0..31 and 32..63), warp 8 = TMA producer, warp 9 = MMA issuer.  The row tiles are independent, so S(i+1) is issued as
// soon as S(i) has been read out (if its q tile has landed, else right after P.V(i)), and the P.V product of tile i is
// collected and stored one iteration later, under the softmax of tile i+1.
__global__ void __launch_bounds__(320, 2)
attn_out_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                   const float* __restrict__ inv, const int* __restrict__ cu_rows, const float* __restrict__ k_land,
                   const float* __restrict__ w_mat, float* __restrict__ attn, float* __restrict__ w_max_out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    constexpr int oK = 0, oW = 16384, oQ = 32768, oVec = 32768 + 65536;
    float* inv_kl = reinterpret_cast<float*>(g + oVec);                     // [64]
    float* s_pmax = inv_kl + 64;                                            // [2][128]
    float* s_psum = s_pmax + 256;                                           // [2][128]
    unsigned* s_wmax = reinterpret_cast<unsigned*>(g + oVec + kAoVecBytes);
    // barriers: full[2] +0, empty[2] +16, S done +32, P.V done +40, S read out +48, P in place +56; TMEM slot +64
    const uint32_t bars = base + oVec + kAoVecBytes + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + oVec + kAoVecBytes + 16 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.x, v = blockIdx.y;
    const VidInfo vi = vid_info(cu_rows, v);
    // gridDim.z > 1 (few, long videos): CTA z takes the z-th contiguous range of 128-row tiles (rows are independent)
    const int tiles_all = (vi.T + 127) / 128;
    const int per = (tiles_all + (int)gridDim.z - 1) / (int)gridDim.z;
    const int tile0 = (int)blockIdx.z * per;
    const int n_tiles = max(0, min(per, tiles_all - tile0));
    const int rbase = tile0 * 128;                                          // first row of this CTA's range
    const size_t hoff = ((size_t)v * kHeads + h) * 4096;

    if (tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 8, 1);
        mbar_init(bars + 16, 1); mbar_init(bars + 24, 1);
        mbar_init(bars + 32, 1); mbar_init(bars + 40, 1);
        mbar_init(bars + 48, 256); mbar_init(bars + 56, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *s_wmax = 0u;
    }
    if (warp == 8) tmem_alloc(bars + 64, 256);
    __syncthreads();
    // Prologue: k_land and W (64 x 64 fp32 each, L2 resident) -> operand planes.  All 256 row threads take part and the
    // loads are coalesced (float4 number tid + 256 k: a matrix row = 16 consecutive lanes, its maximum one half-warp
    // reduction); one thread per row with 16 strided float4 loads each kept 192 threads idle behind 128 and was 9 % of
    // this kernel's stall samples (ncu r02w).
    float4 kv[4], wv[4];
    auto put4 = [&](unsigned char* hi, unsigned char* lo, int idx, const float4& x, float mul) {
        const int row = idx >> 4, c4 = idx & 15;
        const float v0 = x.x * mul, v1 = x.y * mul, v2 = x.z * mul, v3 = x.w * mul;
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1), h2 = __float2half_rn(v2), h3 = __float2half_rn(v3);
        __half2 hh[2] = {__halves2half2(h0, h1), __halves2half2(h2, h3)};
        __half2 ll[2] = {__halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1))),
                         __halves2half2(__float2half_rn(v2 - __half2float(h2)), __float2half_rn(v3 - __half2float(h3)))};
        const uint32_t off = sw128_off(row, c4 >> 1) + (uint32_t)((c4 & 1) * 8);
        *reinterpret_cast<uint2*>(hi + off) = *reinterpret_cast<uint2*>(hh);
        *reinterpret_cast<uint2*>(lo + off) = *reinterpret_cast<uint2*>(ll);
    };
    if (tid < 256) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            kv[k] = ldg4(k_land + hoff + (size_t)(tid + 256 * k) * 4);
            wv[k] = ldg4(w_mat + hoff + (size_t)(tid + 256 * k) * 4);
        }
        float wm = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + 256 * k;
            const float mx = half_warp_max(fmaxf(fmaxf(fabsf(kv[k].x), fabsf(kv[k].y)), fmaxf(fabsf(kv[k].z), fabsf(kv[k].w))));
            const int e = scale_exp(mx);
            if ((tid & 15) == 0) inv_kl[idx >> 4] = ldexpf(1.f, -e);
            put4(g + oK, g + oK + 8192, idx, kv[k], ldexpf(1.f, e));
            wm = fmaxf(wm, fmaxf(fmaxf(fabsf(wv[k].x), fabsf(wv[k].y)), fmaxf(fabsf(wv[k].z), fabsf(wv[k].w))));
        }
        wm = warp_max(wm);
        if (lane == 0) atomicMax(s_wmax, __float_as_uint(wm));
    }
    __syncthreads();
    const int ew = scale_exp(__uint_as_float(*s_wmax));
    if (tid < 256) {
        const float wmul = ldexpf(1.f, ew);
#pragma unroll
        for (int k = 0; k < 4; ++k) put4(g + oW, g + oW + 8192, tid + 256 * k, wv[k], wmul);
    }
    // max|W| of this (video, head): every output row is a convex combination of W's rows, so this bounds the
    // attention part of `merged` (value_conv_kernel derives the operand-plane scale from it)
    if (tid == 0 && w_max_out != nullptr && blockIdx.z == 0) w_max_out[((size_t)v * kHeads + h) * 2] = __uint_as_float(*s_wmax);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 8) {
        if (lane == 0) {
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i & 1;
                if (!mbar_wait(bars + 16 + 8 * s, ((uint32_t)(i >> 1) & 1u) ^ 1u)) break;
                const uint32_t st = base + oQ + s * 32768;
                mbar_expect_tx(bars + 8 * s, 32768);
                const int row = vi.row0 + rbase + i * 128, col = h * kDimHead;
                tma_load_2d(st, &map_hi, bars + 8 * s, col, row);
                tma_load_2d(st + 8192, &map_hi, bars + 8 * s, col, row + 64);
                tma_load_2d(st + 16384, &map_lo, bars + 8 * s, col, row);
                tma_load_2d(st + 24576, &map_lo, bars + 8 * s, col, row + 64);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ---- MMA issuer ----
            bool mok = true;
            auto issue_s = [&](int i) {
                const uint32_t st = base + oQ + (i & 1) * 32768;
                tc_fence_after();
                issue_split_mma64<false>(tmem_base, tmem_base + 64u, st, st + 16384, base + oK, base + oK + 8192);
                umma_commit(bars + 32);
            };
            if (n_tiles > 0) {
                mok = mbar_wait(bars, 0u);
                issue_s(0);
            }
            for (int i = 0; i < n_tiles && mok; ++i) {
                const int s = i & 1;
                const uint32_t st = base + oQ + s * 32768;
                bool next_issued = i + 1 >= n_tiles;
                const uint32_t full_next = bars + 8 * ((i + 1) & 1), ph_next = (uint32_t)((i + 1) >> 1) & 1u;
                mok = mbar_wait(bars + 48, (uint32_t)i & 1u) && mok;        // S(i) is in registers everywhere
                if (!next_issued && mbar_try_wait(full_next, ph_next)) {
                    issue_s(i + 1);
                    next_issued = true;
                }
                mok = mbar_wait(bars + 56, (uint32_t)i & 1u) && mok;        // P(i) in its q stage, O(i-1) read out
                tc_fence_after();
                issue_split_mma64<true>(tmem_base + 128u, tmem_base + 192u, st, st + 16384, base + oW, base + oW + 8192);
                umma_commit(bars + 40);
                umma_commit(bars + 16 + 8 * s);                             // stage free once P has been consumed
                if (!next_issued) {
                    mok = mbar_wait(full_next, ph_next) && mok;
                    issue_s(i + 1);
                }
            }
        }
    } else {
        const int trow = tid & 127, half = tid >> 7;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + (uint32_t)(half * 32), tO = tS + 128u;
        const float o_scale = ldexpf(1.f, -ew) * (1.f / 16384.f);
        uint32_t s_phase = 0, pv_phase = 0;
        bool ok = true;
        float inv_q = rbase + trow < vi.T ? __ldg(inv + (size_t)(vi.row0 + rbase + trow) * 24 + h) : 0.f;
        float rs_prev = 0.f;
        // P.V product of tile j: read out, scale by the row's 1 / sum, store
        // Stored straight from the accumulator layout a warp instruction would touch 32 rows x 16 bytes (32 L1 tag
        // look-ups).  The two warps of a TMEM lane quarter share a 32 x 64 fp32 region instead (as the value convolution
        // does): each dumps its 32 columns, a 64-thread named barrier, each reads 16 rows back with eight lanes per row
        // and stores 4 rows x 256 contiguous bytes per instruction.  The region needs no memory of its own: when
        // collect(i - 1) runs, S(i) has been read out, so the q tile of stage i & 1 is dead, and the bytes a pair
        // borrows (rows 32 quarter .. + 31 of both planes, 2 x 4 KB) are exactly the rows the SAME two warps fill with
        // P(i) right afterwards.
        const int quarter = warp & 3, bar_id = 2 + quarter;
        const int rsub = lane >> 3, c8 = lane & 7;
        auto reg_ptr = [&](unsigned char* stage, int r, int slot) -> float* {      // 16-byte slot `slot` of region row r
            return reinterpret_cast<float*>(stage + (slot >> 3) * 16384 + (quarter * 32 + r) * 128 +
                                            (((slot & 7) ^ (r & 7) ^ (slot >> 3)) & 7) * 16);
        };
        auto collect = [&](int j, float rs, unsigned char* stage) {
            ok = mbar_wait(bars + 40, pv_phase) && ok;
            pv_phase ^= 1u;
            tc_fence_after();
            float ov[32];
            tmem_read32_sum(tO, tO + 64u, ov);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                st4(reg_ptr(stage, lane, half * 8 + q),
                    make_float4(ov[4 * q] * rs, ov[4 * q + 1] * rs, ov[4 * q + 2] * rs, ov[4 * q + 3] * rs));
            named_bar_sync(bar_id, 64);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = half * 16 + 4 * i + rsub;
                const int row = rbase + j * 128 + quarter * 32 + rl;
                const float4 o0 = lds4(reg_ptr(stage, rl, 2 * c8)), o1 = lds4(reg_ptr(stage, rl, 2 * c8 + 1));
                if (row < vi.T) {
                    float* dst = attn + (size_t)(vi.row0 + row) * kInner + h * kDimHead + c8 * 8;
                    st4(dst, o0);
                    st4(dst + 4, o1);
                }
            }
            ok = named_bar_and(bar_id, 64, ok);                             // the rows are free for P again
        };
        for (int i = 0; i < n_tiles && ok; ++i) {
            const int s = i & 1;
            const int row = rbase + i * 128 + trow;
            const float inv_q_next = (row + 128 < vi.T) ? __ldg(inv + (size_t)(vi.row0 + row + 128) * 24 + h) : 0.f;
            unsigned char* stp = g + oQ + s * 32768;
            ok = mbar_wait(bars + 32, s_phase) && ok;
            s_phase ^= 1u;
            tc_fence_after();
            float p[32];
            tmem_read32_sum(tS, tS + 64u, p);
            tc_fence_before();
            mbar_arrive(bars + 48);                                         // this thread holds its S(i) values
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) { p[j] *= inv_q * inv_kl[half * 32 + j]; mx = fmaxf(mx, p[j]); }
            s_pmax[half * 128 + trow] = mx;
            named_bar_sync(1, 256);
            mx = fmaxf(mx, s_pmax[(half ^ 1) * 128 + trow]);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
            s_psum[half * 128 + trow] = sum;
            // the previous tile's output: its accumulator is about to be overwritten by P.V(i)
            if (i > 0) collect(i - 1, rs_prev, stp);
            // un-normalised probabilities (<= 1, fixed scale 2^14); the row sum divides the output instead
            store_row32(stp, stp + 16384, trow, half * 4, p, 16384.f);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bars + 56);                                         // P(i) stored, O(i-1) read out
            named_bar_sync(1, 256);                                         // partner's row sum visible
            rs_prev = o_scale / (sum + s_psum[(half ^ 1) * 128 + trow]);
            inv_q = inv_q_next;
        }
        if (n_tiles > 0 && ok) collect(n_tiles - 1, rs_prev, g + oQ + (n_tiles & 1) * 32768);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------
// Iterative Moore-Penrose pseudo-inverse of attn2 (nystroformer.py:13-28) and W = Z a3v on tensor cores, TWO heads of
// one video per CTA.  Every 64 x 64 matrix of the chain is kept as a [128 rows][128 B] plane tile (head 0 rows 0..63,
// head 1 rows 64..127) with ONE power-of-two scale per head.  The same bytes serve as left operand (K-major, the two
// heads stacked along M) and as right operand (MN-major, the two heads side by side along N, LBO = 8 KB): each
// 128 x 128 product holds the two wanted 64 x 64 products on its block diagonal, thread t reads row t of its own block.
// Per product: row -> registers -> (7I - ., 15I - ., ...) -> matrix max (shuffle + one barrier) -> planes -> MMA.
// grid (4 head pairs, V), 256 threads, TMEM 256 columns (main | cross).  Only THREE tiles live in shared memory: Z, a
// tile Q that is A during the first product of an iteration and XZ afterwards, and a tile P that is in turn T1, U, T2
// and a3v (a product's result may overwrite an operand once the MMAs have completed).  A's planes are parked in
// global memory by the prologue (in the video's zmat slot, 32 KB per head pair, overwritten by Z at the end) and
// copied back into Q under the third product of every iteration.  96 KB per CTA: TWO CTAs share an SM, and one
// chain's MMA / barrier latencies are covered by the other's.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPinvTile = 32768;
constexpr int kPinvTcSmemBytes = 3 * kPinvTile + 256 + 1024;

__device__ __forceinline__ void issue_pinv_product(uint32_t tmem_base, uint32_t left, uint32_t right) {
    constexpr uint32_t idesc = make_idesc_bmn(128, 128);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t ka = (uint32_t)k * 32u, kb = (uint32_t)k * 2048u;
        const uint64_t dah = make_smem_desc<64>(left + ka), dal = make_smem_desc<64>(left + 16384 + ka);
        const uint64_t dbh = make_smem_desc_mn2(right + kb), dbl = make_smem_desc_mn2(right + 16384 + kb);
        umma_f16(tmem_base, dah, dbh, idesc, k != 0 ? 1u : 0u);
        umma_f16(tmem_base + 128u, dah, dbl, idesc, k != 0 ? 1u : 0u);
        umma_f16(tmem_base + 128u, dal, dbh, idesc, 1u);
    }
}

// 256 threads: thread t and t + 128 share accumulator row (t & 127); the first owns columns 0..31 of its head's block,
// the second columns 32..63 (warps w and w + 4 may touch the same TMEM lane quarter).  Two warps per scheduler keep
// the dependent chain's ALU latency covered.
__global__ void __launch_bounds__(256, 2)
pinv_w_tc_kernel(const float* __restrict__ attn2, const float* __restrict__ stats, const float* __restrict__ a3v,
                 float* __restrict__ w_out, float* __restrict__ z_out, int iters) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    constexpr int oZ = 0, oQ = kPinvTile, oP = 2 * kPinvTile, oVec = 3 * kPinvTile;
    float* s_mx = reinterpret_cast<float*>(g + oVec);                       // [2 parities][2 slots][8 warps]
    const uint32_t bar = base + oVec + 128;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + oVec + 144);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = tid >> 7;                            // tile row, column half
    const int hh = row >> 6, i = row & 63, c0 = half * 32;
    const int v = blockIdx.y, h = blockIdx.x * 2 + hh;
    const size_t off = ((size_t)v * kHeads + h) * 4096;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(base + oVec + 144, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t t_main = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(hh * 64 + c0), t_lo = t_main + 128u;
    uint32_t phase = 0, par = 0;
    bool ok = true;

    // head-wide maximum of up to two per-thread values (one shuffle reduction + one barrier)
    auto head_max2 = [&](float m0, float m1, float& o0, float& o1) {
        m0 = warp_max(m0); m1 = warp_max(m1);
        float* slot = s_mx + par * 16;
        if (lane == 0) { slot[warp] = m0; slot[8 + warp] = m1; }
        tc_fence_before();
        __syncthreads();
        // rows of head hh live in warps 2hh, 2hh+1 (columns 0..31) and 4+2hh, 5+2hh (columns 32..63)
        o0 = fmaxf(fmaxf(slot[2 * hh], slot[2 * hh + 1]), fmaxf(slot[4 + 2 * hh], slot[5 + 2 * hh]));
        o1 = fmaxf(fmaxf(slot[8 + 2 * hh], slot[9 + 2 * hh]), fmaxf(slot[12 + 2 * hh], slot[13 + 2 * hh]));
        par ^= 1u;
    };
    // store this thread's half row (raw accumulator units, true value = raw * unscale) into tile `o` with the head-wide
    // scale 2^e chosen from the true maximum; returns 2^-e
    auto put = [&](int o, const float (&r)[32], float unscale, float head_mx_raw) -> float {
        const int e = scale_exp(head_mx_raw * fabsf(unscale));
        store_row32(g + o, g + o + 16384, row, half * 4, r, unscale * ldexpf(1.f, e));
        return ldexpf(1.f, -e);
    };
    // A's planes parked in global memory: this CTA's 32 KB slot, same byte layout as the shared-memory tile
    uint4* a_park = reinterpret_cast<uint4*>(z_out + ((size_t)v * kHeads + blockIdx.x * 2) * 4096);
    auto product = [&](int left, int right, float (&out)[32], bool reload_a) {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_pinv_product(tmem_base, base + left, base + right);
            umma_commit(bar);
        }
        uint4 a_copy[8];
        if (reload_a) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a_copy[k] = a_park[k * 256 + tid];       // in flight under the MMAs
        }
        ok = mbar_wait(bar, phase) && ok;
        phase ^= 1u;
        tc_fence_after();
        if (reload_a) {
            // the product that just completed was the last reader of XZ: tile Q takes A again for the next iteration
#pragma unroll
            for (int k = 0; k < 8; ++k) reinterpret_cast<uint4*>(g + oQ)[k * 256 + tid] = a_copy[k];
        }
#pragma unroll
        for (int q = 0; q < 32; q += 16) {
            uint32_t r0[16], r1[16];
            tmem_ld16_nowait(t_main + (uint32_t)q, r0);
            tmem_ld16_nowait(t_lo + (uint32_t)q, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) out[q + j] = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
        }
    };
    const bool has_diag = (i >= c0) && (i < c0 + 32);
    // t = d * I - t * s  for this thread's columns (s and d / s are powers of two times small integers: exact)
    auto diag_minus = [&](float (&t)[32], float s, float d) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = (has_diag && j == i - c0 ? d : 0.f) - t[j] * s;
    };

    // ---- start: A = attn2, Z0 = A^T / (max row sum * max column sum over ALL 8 heads of the video) ----
    float mrow = 0.f, mcol = 0.f;
#pragma unroll
    for (int q = 0; q < kHeads; ++q) {
        mrow = fmaxf(mrow, __ldg(stats + ((size_t)v * kHeads + q) * 2 + 0));
        mcol = fmaxf(mcol, __ldg(stats + ((size_t)v * kHeads + q) * 2 + 1));
    }
    const float denom = mrow * mcol;
    float z[32], t[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 x = ldg4(attn2 + off + i * 64 + c0 + j);               // row i of A
        t[j] = x.x; t[j + 1] = x.y; t[j + 2] = x.z; t[j + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = __ldg(attn2 + off + (c0 + j) * 64 + i) / denom;    // row i of A^T
    float mA, mZ;
    head_max2(absmax32(t), absmax32(z), mA, mZ);
    const float invA = put(oQ, t, 1.f, mA);
    float invZ = put(oZ, z, 1.f, mZ);
    __syncthreads();
    // park A's planes (the attn2 rows above were the last reads of anything this slot could alias)
#pragma unroll
    for (int k = 0; k < 8; ++k) a_park[k * 256 + tid] = reinterpret_cast<const uint4*>(g + oQ)[k * 256 + tid];

    for (int it = 0; it < iters && ok; ++it) {
        // XZ = A Z ; T1 = 7I - XZ
        float xz[32];
        product(oQ, oZ, xz, false);
        const float sXZ = invA * invZ;
#pragma unroll
        for (int j = 0; j < 32; ++j) { xz[j] *= sXZ; t[j] = (has_diag && j == i - c0 ? 7.f : 0.f) - xz[j]; }
        float mXZ, mT;
        head_max2(absmax32(xz), absmax32(t), mXZ, mT);
        const float invXZ = put(oQ, xz, 1.f, mXZ);
        float invT = put(oP, t, 1.f, mT);
        // U = 15I - XZ T1
        product(oQ, oP, t, false);
        diag_minus(t, invXZ * invT, 15.f);
        float mU, dummy;
        head_max2(absmax32(t), 0.f, mU, dummy);
        const float invU = put(oP, t, 1.f, mU);
        // T2 = 13I - XZ U   (XZ's last use: A comes back into its tile)
        product(oQ, oP, t, it + 1 < iters);
        diag_minus(t, invXZ * invU, 13.f);
        head_max2(absmax32(t), 0.f, mT, dummy);
        invT = put(oP, t, 1.f, mT);
        // Z' = 0.25 Z T2
        product(oZ, oP, z, false);
        const float sZ = 0.25f * invZ * invT;
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] *= sZ;
        head_max2(absmax32(z), 0.f, mZ, dummy);
        invZ = put(oZ, z, 1.f, mZ);
    }
    if (z_out != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) st4(z_out + off + i * 64 + c0 + j, make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]));
    }
    // ---- W = Z a3v ----
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 x = ldg4(a3v + off + i * 64 + c0 + j);
        t[j] = x.x; t[j + 1] = x.y; t[j + 2] = x.z; t[j + 3] = x.w;
    }
    float mV, dummy2;
    head_max2(absmax32(t), 0.f, mV, dummy2);
    const float invV = put(oP, t, 1.f, mV);
    product(oZ, oP, t, false);
    if (ok) {
        const float sW = invZ * invV;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            st4(w_out + off + i * 64 + c0 + j, make_float4(t[j] * sW, t[j + 1] * sW, t[j + 2] * sW, t[j + 3] * sW));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------
// value convolution + merge: merged[r][h*64+c] = attn[r][h*64+c] + sum_t w[h][t] * v[r + t - 16][h*64+c], rows
// outside the video are zero (nystroformer.py:61-65,137-138).
// grid (n_tiles128, 4): a CTA owns 128 rows x 128 value columns (two heads).  The 160 x 128 input window is rebuilt
// from the planes ((hi + lo) * inv, 128-bit loads) into shared memory once; then thread <-> (column, 64-row half) walks
// its rows in blocks of 8 with a 40-deep register window: 33 FMA per shared-memory load, attn rows prefetched.
// ---------------------------------------------------------------------------------------------------------
constexpr int kConvRowsIn = 128 + kTaps - 1;                                 // 160
constexpr int kConvSmemBytes = kConvRowsIn * 128 * 4;

__global__ void __launch_bounds__(256, 2)
value_conv_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, const float* __restrict__ inv,
                  const int* __restrict__ cu_rows, const int2* __restrict__ tiles, const float* __restrict__ conv_w,
                  float* __restrict__ merged) {
    extern __shared__ __align__(16) float vs[];                              // [160][128]
    const int2 tile = tiles[blockIdx.x];
    const VidInfo vi = vid_info(cu_rows, tile.x);
    const int r0 = tile.y, cb = blockIdx.y * 128, tid = threadIdx.x;
    for (int task = tid; task < kConvRowsIn * 16; task += 256) {
        const int rr = task >> 4, ch = task & 15;
        const int r = r0 - 16 + rr;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r >= 0 && r < vi.T) {
            const size_t row = (size_t)(vi.row0 + r);
            const size_t off = row * kQkvCols + 2 * kInner + cb + ch * 8;
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + off));
            const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + off));
            const float s = __ldg(inv + row * 24 + 16 + ((cb + ch * 8) >> 6));
            const __half2* hp = reinterpret_cast<const __half2*>(&h);
            const __half2* lp = reinterpret_cast<const __half2*>(&l);
            const float2 h0 = __half22float2(hp[0]), h1 = __half22float2(hp[1]), h2 = __half22float2(hp[2]), h3 = __half22float2(hp[3]);
            const float2 l0 = __half22float2(lp[0]), l1 = __half22float2(lp[1]), l2 = __half22float2(lp[2]), l3 = __half22float2(lp[3]);
            a = make_float4((h0.x + l0.x) * s, (h0.y + l0.y) * s, (h1.x + l1.x) * s, (h1.y + l1.y) * s);
            b = make_float4((h2.x + l2.x) * s, (h2.y + l2.y) * s, (h3.x + l3.x) * s, (h3.y + l3.y) * s);
        }
        st4(vs + rr * 128 + ch * 8, a);
        st4(vs + rr * 128 + ch * 8 + 4, b);
    }
    const int c = tid & 127, base = (tid >> 7) * 64;
    float w[kTaps];
#pragma unroll
    for (int t = 0; t < kTaps; ++t) w[t] = __ldg(conv_w + ((cb + c) >> 6) * kTaps + t);
    float* out = merged + (size_t)(vi.row0 + r0 + base) * kInner + cb + c;
    const int rows = min(64, vi.T - r0 - base);                              // output rows of this thread (may be <= 0)
    float an[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) an[j] = j < rows ? out[(size_t)j * kInner] : 0.f;
    __syncthreads();
    float win[40];                           // win[i] = input row (base + b + i) of the staged window
#pragma unroll
    for (int i = 0; i < 32; ++i) win[i] = vs[(base + i) * 128 + c];
    for (int b = 0; b < rows; b += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { win[32 + j] = vs[(base + b + 32 + j) * 128 + c]; acc[j] = an[j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) an[j] = (b + 8 + j < rows) ? out[(size_t)(b + 8 + j) * kInner] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int i = 0; i < kTaps; ++i) acc[j] = fmaf(w[i], win[j + i], acc[j]);
            if (b + j < rows) out[(size_t)(b + j) * kInner] = acc[j];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) win[i] = win[i + 8];
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same value convolution + merge on the tensor cores (tcgen05 precisions, to_out operand planes out).
// A 33-tap FIR along the rows is a banded Toeplitz product:  conv[r][c] = sum_j Band[r][j] V[j][c] with
// Band[r][j] = taps[j - r] for 0 <= j - r <= 32 over the 160-row window (rows r0-16 .. r0+143) of a 128-row tile.
//   * B operand = the v planes themselves: three 64-row TMA boxes per plane straight out of the q|k|v planes
//     (MN-major, 128-byte swizzle).  Window rows outside the video arrive as whatever neighbours them in the packed
//     buffer (or as TMA zero fill) and are switched off through the band.
//   * A operand = Band' = taps[j - r] * 2^-sv[j] * 2^c in shared memory (K-major, 3 K blocks of 64, hi / lo planes):
//     v's per-row plane scale 2^sv[j] varies along K, where a B operand cannot carry it, so its inverse is folded into
//     the band column j (a power of two: exact), together with the row mask (0 outside the video).  The band positions
//     are the same for every tile and head; only the 33 values per row are rewritten (the rest stays zero).
//   * 12 K steps x {hi.hi -> main; hi.lo, lo.hi -> cross} per head, accumulators double-buffered in TMEM: the MMAs of
//     head h+1 run under the epilogue of head h (TMEM -> + attention part -> plane scale -> hi / lo planes).
// One CTA per 128-row tile, all 8 heads; 320 threads: warps 0..7 = rows (two threads per row, 32 columns each),
// warp 8 = TMA producer, warp 9 = MMA issuer.
// Output: merged leaves as the to_out GEMM's operand planes (hi [R][512] | lo [R][512] | inverse row scale [R], the
// layout of edsnet_split_f16), which removes the separate split pass.  A plane scale only has to keep the row's values
// inside fp16 range with the low plane normal for every element that matters, so it is taken from a BOUND:
// |attention part| <= max|W| of the head (attn_out_tc_kernel leaves it in w_max), |convolution part| <= sum|taps| *
// max|v| over the window rows (from v's plane scales); one power of two per 128-row tile, maximum over the 8 heads.
// ---------------------------------------------------------------------------------------------------------
constexpr int kCvBandBytes = 3 * 32768;                                    // [kb][hi | lo][128 rows][128 B]
constexpr int kCvVPlane = kConvRowsIn * 128;                                // 160 window rows x 128 B = 20 KB
constexpr int kCvVStage = 2 * kCvVPlane;                                   // hi | lo
constexpr int kCvStages = 2;
constexpr int kCvScratch = 4 * 32 * 256;                                   // per TMEM lane quarter (a warp pair): 32 rows x 64 columns fp32
constexpr int kCvVecFloats = kHeads * 192 + kHeads * kTaps + 3 * kHeads + 4;   // inv window, taps, e_c / f-scale, scale
constexpr int kCvSmemBytes = kCvBandBytes + kCvStages * kCvVStage + kCvScratch + kCvVecFloats * 4 + 128 + 1024;

__global__ void __launch_bounds__(320, 1)
value_conv_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                     const float* __restrict__ inv, const int* __restrict__ cu_rows, const int2* __restrict__ tiles,
                     const float* __restrict__ conv_w, const float* __restrict__ attn, const float* __restrict__ w_max,
                     __half* __restrict__ m_hi, __half* __restrict__ m_lo, float* __restrict__ m_inv, int write_lo,
                     int n_tiles) {
    // write_lo == 0: the to_out product that follows runs two passes (A_hi only): the lo plane is not written
    // Persistent: one CTA per SM walks the 128-row tiles blockIdx.x, blockIdx.x + gridDim.x, ... (a tile is the same work
    // whatever its video).  A CTA of this size costs 3-4 us to launch on an SM that has just been vacated (globaltimer
    // probe on the a3v kernel), and the band zeroing, tap staging, barrier set-up and TMEM allocation are per CTA: all of it
    // is paid once per SM instead of once per tile.  Every barrier completes an even number of phases per tile, so the
    // parities below repeat from tile to tile.
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    constexpr int oBand = 0, oV = kCvBandBytes, oScr = oV + kCvStages * kCvVStage, oVec = oScr + kCvScratch;
    float* s_inv = reinterpret_cast<float*>(g + oVec);                     // [8 heads][192 window rows], 0 = row masked
    float* s_w = s_inv + kHeads * 192;                                      // [8][33]
    float* s_fc = s_w + kHeads * kTaps;                                     // [8] 2^e_c per head
    float* s_ic = s_fc + kHeads;                                            // [8] 2^-e_c
    unsigned* s_vmax = reinterpret_cast<unsigned*>(s_ic + kHeads);         // [8] bound on |merged| per head (float bits)
    // barriers: V full[kCvStages] +0, V empty[kCvStages] +24, band ready +48, MMA done +56; TMEM slot +64
    const uint32_t bars = base + oVec + kCvVecFloats * 4;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + oVec + kCvVecFloats * 4 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int q = 0; q < kCvStages; ++q) { mbar_init(bars + 8 * q, 1); mbar_init(bars + 24 + 8 * q, 1); }
        mbar_init(bars + 48, 256);
        mbar_init(bars + 56, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(bars + 64, 256);
    // zero the band once (only its 33 diagonals are ever rewritten, by every tile in full), stage the taps
    for (int i = tid; i < kCvBandBytes / 16; i += 320) reinterpret_cast<uint4*>(g + oBand)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < kHeads * kTaps; i += 320) s_w[i] = __ldg(conv_w + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    uint32_t phase = 0;                                                     // row threads: parity of the MMA-done barrier

    for (int ti = blockIdx.x; ti < n_tiles; ti += (int)gridDim.x) {
    const int2 tile = tiles[ti];
    const VidInfo vi = vid_info(cu_rows, tile.x);
    const int r0 = tile.y;
    const int win0 = vi.row0 + r0 - 16;                                     // packed row of window row 0 (may be < 0)
    __syncthreads();                                                        // (everyone has left the previous tile)
    // Per head (a warp each): inverse v plane scales of the window rows (0 outside the video: masks the band column),
    // their maximum, the band exponent e_c that puts the largest band entry into [2^14, 2^15), the band column factors
    // 2^(e_c - sv[j]), and the head's bound on |merged|; one barrier instead of four and no shared-memory atomics
    // (the four-phase version was 17 % of this kernel's stall samples, ncu r02w)
    if (warp < kHeads) {
        const int hd = warp;
        float f[6], vmi = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int j = lane + 32 * k, r = r0 - 16 + j;
            f[k] = (r >= 0 && r < vi.T && j < kConvRowsIn) ? __ldg(inv + (size_t)(vi.row0 + r) * 24 + 16 + hd) : 0.f;
            vmi = fmaxf(vmi, f[k]);
        }
        vmi = warp_max(vmi);
        float l1 = 0.f, wm = 0.f;
        for (int t = lane; t < kTaps; t += 32) { const float a = fabsf(s_w[hd * kTaps + t]); l1 += a; wm = fmaxf(wm, a); }
        l1 = warp_sum(l1);
        wm = warp_max(wm);
        const int ec = scale_exp(wm * vmi);
        const float fc = ldexpf(1.f, ec);
#pragma unroll
        for (int k = 0; k < 6; ++k) s_inv[hd * 192 + lane + 32 * k] = f[k] * fc;
        if (lane == 0) {
            s_fc[hd] = fc;
            s_ic[hd] = ldexpf(1.f, -ec);
            s_vmax[hd] = __float_as_uint(__ldg(w_max + ((size_t)tile.x * kHeads + hd) * 2) + l1 * (32768.f * vmi));
        }
    }
    __syncthreads();
    float bound = 0.f;
#pragma unroll
    for (int hd = 0; hd < kHeads; ++hd) bound = fmaxf(bound, __uint_as_float(s_vmax[hd]));
    const float sc = ldexpf(1.f, scale_exp(bound));

    if (warp == 8) {
        if (lane == 0) {
            // ---- TMA producer: the v planes of head hd, 192 window rows ----
            for (int hd = 0; hd < kHeads; ++hd) {
                const int s = hd % kCvStages;
                if (!mbar_wait(bars + 24 + 8 * s, ((uint32_t)(hd / kCvStages) & 1u) ^ 1u)) break;
                const uint32_t st = base + oV + s * kCvVStage;
                mbar_expect_tx(bars + 8 * s, kCvVStage);
                const int col = 2 * kInner + hd * kDimHead;
#pragma unroll
                for (int q = 0; q < kConvRowsIn / 32; ++q) {                // boxes of 32 rows x 64 columns
                    tma_load_2d(st + q * 4096, &map_hi, bars + 8 * s, col, win0 + q * 32);
                    tma_load_2d(st + kCvVPlane + q * 4096, &map_lo, bars + 8 * s, col, win0 + q * 32);
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ---- MMA issuer ----
            // A_hi [V_hi | V_lo] as one N = 128 instruction (the two value planes of a stage are two MN-major atoms
            // kCvVPlane bytes apart), A_lo V_hi into the cross half: 20 instead of 30 instructions per head
            constexpr uint32_t idesc = make_idesc_bmn(128, 64), idesc2 = make_idesc_bmn(128, 128);
            bool mok = true;
            for (int hd = 0; hd < kHeads && mok; ++hd) {
                const int s = hd % kCvStages;
                const uint32_t st = base + oV + s * kCvVStage;
                mok = mbar_wait(bars + 8 * s, (uint32_t)(hd / kCvStages) & 1u) && mok;     // v planes have landed
                mok = mbar_wait(bars + 48, (uint32_t)hd & 1u) && mok;                      // band of this head in place
                tc_fence_after();
                const uint32_t acc_main = tmem_base + (uint32_t)((hd & 1) * 128), acc_lo = acc_main + 64u;
#pragma unroll
                for (int ks = 0; ks < kConvRowsIn / 16; ++ks) {             // 10 K steps of 16 window rows
                    const uint32_t a0 = base + oBand + (ks >> 2) * 32768 + (ks & 3) * 32;
                    const uint32_t b0 = st + ks * 2048;
                    const uint64_t dah = make_smem_desc<64>(a0), dal = make_smem_desc<64>(a0 + 16384);
                    const uint64_t dbh = make_smem_desc_mn(b0);
                    const uint64_t db2 = (uint64_t)((b0 & 0x3FFFFu) >> 4) | ((uint64_t)(kCvVPlane >> 4) << 16) | (64ull << 32) |
                                         (1ull << 46) | (2ull << 61);
                    umma_f16(acc_main, dah, db2, idesc2, ks != 0 ? 1u : 0u);
                    umma_f16(acc_lo, dal, dbh, idesc, 1u);
                }
                umma_commit(bars + 56);
                umma_commit(bars + 24 + 8 * s);                             // v stage free once these MMAs are done
            }
        }
    } else {
        const int r = tid & 127, half = tid >> 7;
        const int row = r0 + r;
        const bool live = row < vi.T;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        // this thread's part of band row r: taps t0 .. t1-1 (columns r + t)
        const int t0 = half * 17, t1 = half ? kTaps : 17;
        auto build_band = [&](int hd) {
            const float* f = s_inv + hd * 192 + r + t0;                     // already times 2^e_c of the head
            const float* wv = s_w + hd * kTaps + t0;
#pragma unroll
            for (int t = 0; t < 17; ++t) {
                if (t0 + t < t1) {
                    const int j = r + t0 + t;
                    const float x = wv[t] * f[t];
                    const __half h = __float2half_rn(x);
                    const __half l = __float2half_rn(x - __half2float(h));
                    const uint32_t off = (uint32_t)((j >> 6) * 32768) + sw128_off(r, (j & 63) >> 3) + (uint32_t)((j & 7) * 2);
                    *reinterpret_cast<__half*>(g + oBand + off) = h;
                    *reinterpret_cast<__half*>(g + oBand + 16384 + off) = l;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bars + 48);
        };
        if (live && half == 0) m_inv[vi.row0 + row] = 1.f / sc;
        build_band(0);
        bool ok = true;
        // Epilogue layout: in TMEM a thread owns one row (32 of its columns), and stored that way a warp instruction
        // would touch 32 rows x 16 bytes.  The two warps of a TMEM lane quarter (column halves 0 and 1) therefore share a
        // 32 x 64 fp32 region (16-byte slots swizzled as in the to_out epilogue, gemm_tc.cuh wide2_slot): each dumps its
        // 32 columns, a 64-thread named barrier, then each reads 16 of the 32 rows with eight lanes per row: the
        // attention part arrives as ONE 32-byte load per lane (4 rows x 256 contiguous bytes per instruction) and a
        // plane leaves as 4 rows x 128 contiguous bytes per instruction (was 8 rows x 64 B in, 8 rows x 32 B out: the
        // L1 data pipe was 53 % busy, ncu r02v).
        float* reg = reinterpret_cast<float*>(g + oScr) + (warp & 3) * (32 * 64);
        const int bar_id = 2 + (warp & 3);
        const int wrow0 = r0 + (warp & 3) * 32 + half * 16;                 // first tile row this warp reads out
        const int rsub = lane >> 3, c8 = lane & 7;
        auto attn_ptr = [&](int hd, int i) -> const float* {
            const int rw = min(wrow0 + 4 * i + rsub, vi.T - 1);
            return attn + (size_t)(vi.row0 + rw) * kInner + hd * kDimHead + c8 * 8;
        };
        float an[4][8];                                                     // attention part, one head ahead
#pragma unroll
        for (int i = 0; i < 4; ++i) ldg8(an[i], attn_ptr(0, i));
        for (int hd = 0; hd < kHeads && ok; ++hd) {
            float a[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = an[i][j];
            if (hd + 1 < kHeads) {
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg8(an[i], attn_ptr(hd + 1, i));
            }
            ok = mbar_wait(bars + 56, phase) && ok;
            phase ^= 1u;
            tc_fence_after();
            if (hd + 1 < kHeads) build_band(hd + 1);                        // the band is free again: next head's MMAs
            const uint32_t t_main = tmem_base + lane_addr + (uint32_t)((hd & 1) * 128 + half * 32);
            float cv[32];
            tmem_read32_sum(t_main, t_main + 64u, cv);
            tc_fence_before();
            const float ic = s_ic[hd];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                st4(reg + lane * 64 + wide2_slot(half * 8 + j, lane) * 4,
                    make_float4(cv[4 * j], cv[4 * j + 1], cv[4 * j + 2], cv[4 * j + 3]));
            named_bar_sync(bar_id, 64);                                     // both column halves of the region are in place
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = half * 16 + 4 * i + rsub, rw = wrow0 + 4 * i + rsub;
                const float4 c0 = lds4(reg + rl * 64 + wide2_slot(2 * c8, rl) * 4);
                const float4 c1 = lds4(reg + rl * 64 + wide2_slot(2 * c8 + 1, rl) * 4);
                const float cvv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                __half2 hh[4], ll[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float v0 = fmaf(cvv[2 * t], ic, a[i][2 * t]) * sc, v1 = fmaf(cvv[2 * t + 1], ic, a[i][2 * t + 1]) * sc;
                    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                    hh[t] = __halves2half2(h0, h1);
                    ll[t] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
                }
                if (rw < vi.T) {
                    const size_t oo = (size_t)(vi.row0 + rw) * kInner + hd * kDimHead + c8 * 8;
                    *reinterpret_cast<uint4*>(m_hi + oo) = *reinterpret_cast<uint4*>(hh);
                    if (write_lo) *reinterpret_cast<uint4*>(m_lo + oo) = *reinterpret_cast<uint4*>(ll);
                }
            }
            ok = named_bar_and(bar_id, 64, ok);                             // region free for the next head
        }
    }
    }   // tiles of this CTA
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 256);
}

}  // namespace tc
