// D applications of the ONE shared fc block {Linear(128,128) + bias, ReLU, LayerNorm(128)} (anchor_based/dsnet.py:91-96,
// 107-108; eval mode) on tcgen05 tensor cores with fp16 hi/lo split operands (fp32-grade, see gemm_tc.cuh).
//
// One CTA = two warpgroups; each warpgroup owns a stream of 128-row tiles, thread <-> row.  Per layer a thread
//   * scales its 128-value row by a power of two, splits it into fp16 hi/lo and writes both planes into shared memory
//     in the K-major 128B-swizzled layout the UMMA descriptor expects (the A operand),
//   * (one elected thread) issues 8 K-steps x {hi.hi -> accumulator 0; hi.lo, lo.hi -> accumulator 1} against the
//     weight planes that stay resident in shared memory for the whole kernel (the B operand),
//   * reads its accumulator row back from TMEM, undoes the scales, adds bias, applies ReLU and the 128-wide LayerNorm
//     entirely in registers -- the row never leaves the thread between layers.
// While one warpgroup is in its register phase the other one's MMAs occupy the tensor pipe.
#pragma once
#include "gemm_tc.cuh"

namespace tc {

constexpr int kFcTile = 128;                                  // rows per warpgroup tile
constexpr int kFcPlane = 128 * 128 * 2;                       // bytes of one fp16 128x128 plane
// W (K half 0: hi | lo, K half 1: hi | lo) | A hi (wg0) | A lo (wg0) | A hi (wg1) | A lo (wg1) | vectors | barriers
constexpr int kFcSlots = 5;                                   // LayerNorm sums (2, by layer parity), row max, head dots (2)
constexpr int kFcXchBytes = 2 * kFcSlots * 2 * 128 * 2 * 4;    // [group][slot][half][row][2] floats
constexpr int kFcVecFloats = 7 * 128;                         // wscale bias gamma beta | w_cls w_loc0 w_loc1
constexpr int kFcSmemBytes = 6 * kFcPlane + kFcVecFloats * 4 + kFcXchBytes + 64 + 1024;

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// byte offset of 16-byte chunk `c16` (0..15 along K = 128 fp16) of row r inside a 128 x 128 fp16 operand plane stored as
// two K halves of [128 rows][128 B], each half in the canonical K-major SWIZZLE_128B layout (8-row / 1024-B atoms).
__device__ __forceinline__ uint32_t fc_plane_off(int r, int c16) {
    return (uint32_t)((c16 >> 3) * 16384 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}

// 512 threads = two groups of 256; a group owns a stream of 128-row tiles.  Thread t and t + 128 of a group share tile
// row (t & 127): the first holds columns 0..63 of the 128-value row, the second 64..127 (warps w and w + 4 touch the
// same TMEM lane quarter).  What the two halves of a row must agree on goes through a small exchange buffer and the
// group's named barrier: the row maximum that picks the plane scale of the first layer (the later layers' inputs are
// LayerNorm outputs, bounded by sqrt(127) max|gamma| + max|beta|: one fixed scale, no exchange) and the row sum /
// sum of squares of every LayerNorm.  Sixteen warps instead of eight keep the per-layer register arithmetic of one
// group under the other group's MMAs and barriers.
// SAVE (training forward, anchor_based/dsnet.py:91-95 in train() mode): Dropout(0.5) after the ReLU (Philox mask of
// common.cuh, survivors x 2; off when drop == 0) and h_save [depth][rows][128] = the rows BEFORE the LayerNorm of every
// application of the block -- all the backward needs (h > 0 <=> kept and active; the LayerNorm is recomputed).
template <bool SAVE>
__global__ void __launch_bounds__(512, 1)
fc_stack_tc_kernel(const float* __restrict__ u_in, const __half* __restrict__ w_planes, const float* __restrict__ w_inv_scale,
                   const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ u_out, int rows, int depth, const float* __restrict__ w_cls,
                   const float* __restrict__ w_loc, float4* __restrict__ heads_out, float* __restrict__ h_save, int drop,
                   unsigned long long seed, unsigned long long offset,
                   const unsigned long long* __restrict__ offset_dev) {
    // offset_dev (optional, device memory): added to `offset` -- lets a captured CUDA graph draw a fresh mask per replay
    // heads_out != nullptr: the three head projections of every output row (u . w_cls, u . w_loc[0], u . w_loc[1]; the
    // ROI pooling and the heads are linear, so the pooling windows then run over these 3 channels) are emitted, 16 bytes
    // per row; u_out may then be nullptr and the 512-byte rows are never written.
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_hi = base;                                             // K half h at h * 32 KB: hi rows, then lo rows
    float* vec = reinterpret_cast<float*>(gbase + 6 * kFcPlane);          // wscale[128] bias[128] gamma[128] beta[128]
    float* xch = vec + kFcVecFloats;                                        // [group][slot][half][row][2]
    const uint32_t bar0 = base + 6 * kFcPlane + kFcVecFloats * 4 + kFcXchBytes;   // mbarrier per group, then the TMEM slot
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gbase + 6 * kFcPlane + kFcVecFloats * 4 + kFcXchBytes + 16);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int grp = tid >> 8, gt = tid & 255;                              // group, thread within it
    const int trow = gt & 127, half = gt >> 7, cb = half * 64;             // tile row, column half, first column

    // ---- one-time setup: weight planes -> swizzled smem, vectors, barriers, TMEM ----
    for (int idx = tid; idx < 2 * 128 * 16; idx += 512) {
        const int plane = idx >> 11, r = (idx >> 4) & 127, c16 = idx & 15;
        const uint4 val = __ldg(reinterpret_cast<const uint4*>(w_planes + (size_t)plane * 128 * 128 + r * 128 + c16 * 8));
        // per K half: [W_hi rows | W_lo rows], 16 KB each, so that [W_hi | W_lo] is ONE 256-row right operand
        *reinterpret_cast<uint4*>(gbase + (c16 >> 3) * 32768 + plane * 16384 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4)) = val;
    }
    if (tid < 128) {
        vec[tid] = __ldg(w_inv_scale + tid);
        vec[128 + tid] = __ldg(bias + tid);
        vec[256 + tid] = __ldg(gamma + tid);
        vec[384 + tid] = __ldg(beta + tid);
        if (heads_out != nullptr) {
            vec[512 + tid] = __ldg(w_cls + tid);
            vec[640 + tid] = __ldg(w_loc + tid);
            vec[768 + tid] = __ldg(w_loc + 128 + tid);
        }
    }
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(bar0 + 16, 512);
    fence_proxy_async();                                                    // weight planes visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t acc_main = tmem_base + (uint32_t)(grp * 256), acc_lo = acc_main + 128u;
    const uint32_t a_hi = base + (2 + 2 * grp) * kFcPlane, a_lo = a_hi + kFcPlane;
    unsigned char* a_hi_ptr = gbase + (2 + 2 * grp) * kFcPlane;
    unsigned char* a_lo_ptr = a_hi_ptr + kFcPlane;
    const uint32_t my_bar = bar0 + 8u * grp;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    constexpr uint32_t idesc = make_idesc(128, 128), idesc2 = make_idesc(128, 256);
    const int n_tiles = (rows + kFcTile - 1) / kFcTile;
    // plane scale of every LayerNorm output: |LN(x)| <= sqrt(127) max|gamma| + max|beta|
    float gmax = 0.f, bmax = 0.f;
    for (int j = 0; j < 128; ++j) { gmax = fmaxf(gmax, fabsf(vec[256 + j])); bmax = fmaxf(bmax, fabsf(vec[384 + j])); }
    const float ln_bound = 11.27f * gmax + bmax;
    int e_ln = 0;
    if (ln_bound > 0.f && ln_bound < INFINITY) e_ln = 14 - ilogbf(ln_bound);
    e_ln = max(-100, min(100, e_ln));
    float* xg = xch + grp * (kFcSlots * 2 * 128 * 2);
    uint32_t phase = 0;
    bool ok = true;

    for (int tile = blockIdx.x * 2 + grp; tile < n_tiles && ok; tile += gridDim.x * 2) {
        const int row = tile * kFcTile + trow;
        float u[64];
        if (row < rows) {
            const float* src = u_in + (size_t)row * kHidden + cb;
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
                const float4 x = ldg4(src + j);
                u[j] = x.x; u[j + 1] = x.y; u[j + 2] = x.z; u[j + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) u[j] = 0.f;
        }
        for (int layer = 0; layer < depth && ok; ++layer) {
            // ---- A operand: half row -> scaled fp16 hi / lo planes ----
            int e = e_ln;
            if (layer == 0) {
                float mx = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) mx = fmaxf(mx, fabsf(u[j]));
                xg[(2 * 2 + half) * 256 + trow * 2] = mx;                   // slot 2: row maximum
                named_bar_sync(1 + grp, 256);
                mx = fmaxf(mx, xg[(2 * 2 + (half ^ 1)) * 256 + trow * 2]);
                e = 0;
                if (mx > 0.f && mx < INFINITY) e = 14 - ilogbf(mx);
                e = max(-100, min(100, e));
            }
            const float sc = ldexpf(1.f, e), inv_a = ldexpf(1.f, -e);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                __half2 hh[4], ll[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float v0 = u[c * 8 + 2 * q] * sc, v1 = u[c * 8 + 2 * q + 1] * sc;
                    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                    hh[q] = __halves2half2(h0, h1);
                    ll[q] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
                }
                const uint32_t off = fc_plane_off(trow, half * 8 + c);
                *reinterpret_cast<uint4*>(a_hi_ptr + off) = *reinterpret_cast<uint4*>(hh);
                *reinterpret_cast<uint4*>(a_lo_ptr + off) = *reinterpret_cast<uint4*>(ll);
            }
            fence_proxy_async();               // generic-proxy smem writes -> visible to the async (tensor core) proxy
            tc_fence_before();                 // orders this thread's earlier tcgen05.ld before the barrier
            named_bar_sync(1 + grp, 256);
            // ---- MMA: one elected thread of the group ----
            if (gt == 0) {
                tc_fence_after();
#pragma unroll
                // A_hi [W_hi | W_lo]^T as ONE N = 256 instruction into the adjacent (main | cross) accumulators, then
                // A_lo W_hi^T into the cross half: 16 instead of 24 instructions per layer (the tensor core takes one
                // instruction per ~90 cycles from an SM, DESIGN 5b, and an N = 128 instruction is only 64 cycles of work)
                for (int k = 0; k < 8; ++k) {
                    const uint32_t koff = (uint32_t)((k >> 2) * 16384 + (k & 3) * 32);
                    const uint32_t woff = (uint32_t)((k >> 2) * 32768 + (k & 3) * 32);
                    const uint64_t dah = make_smem_desc<64>(a_hi + koff), dal = make_smem_desc<64>(a_lo + koff);
                    const uint64_t dbw = make_smem_desc<64>(w_hi + woff);
                    umma_f16(acc_main, dah, dbw, idesc2, k != 0 ? 1u : 0u);
                    umma_f16(acc_lo, dal, dbw, idesc, 1u);
                }
                umma_commit(my_bar);
            }
            ok = mbar_wait(my_bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- epilogue in registers: (main + lo) * scales + bias, ReLU, LayerNorm(128) ----
            float sum = 0.f, sq = 0.f;
            uint32_t kw0 = 0xffffffffu, kw1 = 0xffffffffu;          // keep bits of this thread's 64 columns
            float keep_mul = 1.f;
            if (SAVE && drop) {
                const uint4 dw = dropout_words(seed, offset + (offset_dev != nullptr ? __ldg(offset_dev) : 0ull), row, layer);
                kw0 = half ? dw.z : dw.x;
                kw1 = half ? dw.w : dw.y;
                keep_mul = 2.f;
            }
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(acc_main + lane_addr + (uint32_t)(cb + c0), r0);
                tmem_ld16_nowait(acc_lo + lane_addr + (uint32_t)(cb + c0), r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float acc = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
                    const float lin = fmaf(acc, inv_a * vec[cb + c0 + j], vec[128 + cb + c0 + j]);
                    float val = fmaxf(lin, 0.f);
                    if (SAVE) val = (((c0 + j) < 32 ? kw0 : kw1) >> ((c0 + j) & 31)) & 1u ? val * keep_mul : 0.f;
                    u[c0 + j] = val;
                    sum += val;
                }
            }
            if (SAVE && row < rows) {
                float* hd = h_save + ((size_t)layer * rows + row) * kHidden + cb;
#pragma unroll
                for (int j = 0; j < 64; j += 4) st4(hd + j, make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]));
            }
            // mean over the whole row first, then the centred sum of squares (two exchanges, two-pass variance)
            float* slot = xg + ((layer & 1) * 2) * 256;
            slot[half * 256 + trow * 2] = sum;
            named_bar_sync(1 + grp, 256);
            const float mean = (sum + slot[(half ^ 1) * 256 + trow * 2]) * (1.f / 128.f);
#pragma unroll
            for (int j = 0; j < 64; ++j) { u[j] -= mean; sq = fmaf(u[j], u[j], sq); }
            slot[half * 256 + trow * 2 + 1] = sq;
            named_bar_sync(1 + grp, 256);
            const float rstd = 1.f / sqrtf((sq + slot[(half ^ 1) * 256 + trow * 2 + 1]) * (1.f / 128.f) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 64; ++j) u[j] = fmaf(u[j] * rstd, vec[256 + cb + j], vec[384 + cb + j]);
        }
        if (u_out != nullptr && row < rows) {
            float* dst = u_out + (size_t)row * kHidden + cb;
#pragma unroll
            for (int j = 0; j < 64; j += 4) st4(dst + j, make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]));
        }
        if (heads_out != nullptr) {
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                d0 = fmaf(u[j], vec[512 + cb + j], d0);
                d1 = fmaf(u[j], vec[640 + cb + j], d1);
                d2 = fmaf(u[j], vec[768 + cb + j], d2);
            }
            float* s3 = xg + 3 * 512;                      // slots 3 and 4: written once per tile
            s3[half * 256 + trow * 2] = d0;
            s3[half * 256 + trow * 2 + 1] = d1;
            s3[512 + half * 256 + trow * 2] = d2;
            named_bar_sync(1 + grp, 256);
            if (half == 0 && row < rows)
                heads_out[row] = make_float4(d0 + s3[256 + trow * 2], d1 + s3[256 + trow * 2 + 1],
                                             d2 + s3[512 + 256 + trow * 2], 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

static cudaError_t launch_fc_stack_tc(const float* u_in, const void* w_planes, const float* bias, const float* gamma,
                                      const float* beta, float* u_out, int rows, int depth, cudaStream_t st,
                                      const float* w_cls = nullptr, const float* w_loc = nullptr,
                                      float* heads_out = nullptr, float* h_save = nullptr, int drop = 0,
                                      unsigned long long seed = 0, unsigned long long offset = 0,
                                      const unsigned long long* offset_dev = nullptr) {
    static const char tag = 0;
    if (DeviceOnce once_{&tag}) {
        cudaError_t e = cudaFuncSetAttribute(tc::fc_stack_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             tc::kFcSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::fc_stack_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     tc::kFcSmemBytes);
        if (e != cudaSuccess) return e;
    }
    const int n_pairs = ((rows + tc::kFcTile - 1) / tc::kFcTile + 1) / 2;
    const int grid = n_pairs < tc::num_sms() ? n_pairs : tc::num_sms();
    if (h_save != nullptr)
        tc::fc_stack_tc_kernel<true><<<grid, 512, tc::kFcSmemBytes, st>>>(
            u_in, static_cast<const __half*>(w_planes), split_scales(w_planes, 128, 128), bias, gamma, beta, u_out, rows,
            depth, w_cls, w_loc, reinterpret_cast<float4*>(heads_out), h_save, drop, seed, offset, offset_dev);
    else
        tc::fc_stack_tc_kernel<false><<<grid, 512, tc::kFcSmemBytes, st>>>(
            u_in, static_cast<const __half*>(w_planes), split_scales(w_planes, 128, 128), bias, gamma, beta, u_out, rows,
            depth, w_cls, w_loc, reinterpret_cast<float4*>(heads_out), nullptr, 0, 0ull, 0ull, nullptr);
    return cudaGetLastError();
}
