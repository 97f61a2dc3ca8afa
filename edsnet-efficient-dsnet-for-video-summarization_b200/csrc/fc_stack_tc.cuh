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
// W hi | W lo | A hi (wg0) | A lo (wg0) | A hi (wg1) | A lo (wg1) | vectors | barriers
constexpr int kFcSmemBytes = 6 * kFcPlane + 4 * 128 * 4 + 64 + 1024;

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// byte offset of 16-byte chunk `c16` (0..15 along K = 128 fp16) of row r inside a 128 x 128 fp16 operand plane stored as
// two K halves of [128 rows][128 B], each half in the canonical K-major SWIZZLE_128B layout (8-row / 1024-B atoms).
__device__ __forceinline__ uint32_t fc_plane_off(int r, int c16) {
    return (uint32_t)((c16 >> 3) * 16384 + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(256, 1)
fc_stack_tc_kernel(const float* __restrict__ u_in, const __half* __restrict__ w_planes, const float* __restrict__ w_inv_scale,
                   const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ u_out, int rows, int depth) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_hi = base, w_lo = base + kFcPlane;
    float* vec = reinterpret_cast<float*>(gbase + 6 * kFcPlane);          // wscale[128] bias[128] gamma[128] beta[128]
    const uint32_t bar0 = base + 6 * kFcPlane + 4 * 128 * 4;               // mbarrier per warpgroup, then the TMEM slot
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 6 * kFcPlane + 4 * 128 * 4 + 16);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wg = tid >> 7, wtid = tid & 127;                             // warpgroup, thread within it (= row)

    // ---- one-time setup: weight planes -> swizzled smem, vectors, barriers, TMEM ----
    for (int idx = tid; idx < 2 * 128 * 16; idx += 256) {
        const int plane = idx >> 11, r = (idx >> 4) & 127, c16 = idx & 15;
        const uint4 val = __ldg(reinterpret_cast<const uint4*>(w_planes + (size_t)plane * 128 * 128 + r * 128 + c16 * 8));
        *reinterpret_cast<uint4*>(gbase + plane * kFcPlane + fc_plane_off(r, c16)) = val;
    }
    if (tid < 128) {
        vec[tid] = __ldg(w_inv_scale + tid);
        vec[128 + tid] = __ldg(bias + tid);
        vec[256 + tid] = __ldg(gamma + tid);
        vec[384 + tid] = __ldg(beta + tid);
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
    const uint32_t acc_main = tmem_base + (uint32_t)(wg * 256), acc_lo = acc_main + 128u;
    const uint32_t a_hi = base + (2 + 2 * wg) * kFcPlane, a_lo = a_hi + kFcPlane;
    unsigned char* a_hi_ptr = gbase + (2 + 2 * wg) * kFcPlane;
    unsigned char* a_lo_ptr = a_hi_ptr + kFcPlane;
    const uint32_t my_bar = bar0 + 8u * wg;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    constexpr uint32_t idesc = make_idesc(128, 128);
    const int n_tiles = (rows + kFcTile - 1) / kFcTile;
    uint32_t phase = 0;
    bool ok = true;

    for (int tile = blockIdx.x * 2 + wg; tile < n_tiles && ok; tile += gridDim.x * 2) {
        const int row = tile * kFcTile + wtid;
        float u[128];
        if (row < rows) {
            const float* src = u_in + (size_t)row * kHidden;
#pragma unroll
            for (int j = 0; j < 128; j += 4) {
                const float4 x = ldg4(src + j);
                u[j] = x.x; u[j + 1] = x.y; u[j + 2] = x.z; u[j + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 128; ++j) u[j] = 0.f;
        }
        for (int layer = 0; layer < depth && ok; ++layer) {
            // ---- A operand: row -> scaled fp16 hi / lo planes ----
            float mx = 0.f;
#pragma unroll
            for (int j = 0; j < 128; ++j) mx = fmaxf(mx, fabsf(u[j]));
            int e = 0;
            if (mx > 0.f && mx < INFINITY) e = 14 - ilogbf(mx);
            e = max(-100, min(100, e));
            const float sc = ldexpf(1.f, e), inv_a = ldexpf(1.f, -e);
#pragma unroll
            for (int c16 = 0; c16 < 16; ++c16) {
                __half2 hh[4], ll[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float v0 = u[c16 * 8 + 2 * q] * sc, v1 = u[c16 * 8 + 2 * q + 1] * sc;
                    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                    hh[q] = __halves2half2(h0, h1);
                    ll[q] = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
                }
                const uint32_t off = fc_plane_off(wtid, c16);
                *reinterpret_cast<uint4*>(a_hi_ptr + off) = *reinterpret_cast<uint4*>(hh);
                *reinterpret_cast<uint4*>(a_lo_ptr + off) = *reinterpret_cast<uint4*>(ll);
            }
            fence_proxy_async();               // generic-proxy smem writes -> visible to the async (tensor core) proxy
            tc_fence_before();                 // orders this thread's earlier tcgen05.ld before the barrier
            named_bar_sync(1 + wg, 128);
            // ---- MMA: one elected thread of the warpgroup ----
            if (wtid == 0) {
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t koff = (uint32_t)((k >> 2) * 16384 + (k & 3) * 32);
                    const uint64_t dah = make_smem_desc<64>(a_hi + koff), dal = make_smem_desc<64>(a_lo + koff);
                    const uint64_t dbh = make_smem_desc<64>(w_hi + koff), dbl = make_smem_desc<64>(w_lo + koff);
                    umma_f16(acc_main, dah, dbh, idesc, k != 0 ? 1u : 0u);
                    umma_f16(acc_lo, dah, dbl, idesc, k != 0 ? 1u : 0u);
                    umma_f16(acc_lo, dal, dbh, idesc, 1u);
                }
                umma_commit(my_bar);
            }
            ok = mbar_wait(my_bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- epilogue in registers: (main + lo) * scales + bias, ReLU, LayerNorm(128) ----
            float sum = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 16) {
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(acc_main + lane_addr + (uint32_t)c0, r0);
                tmem_ld16_nowait(acc_lo + lane_addr + (uint32_t)c0, r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float acc = __fadd_rn(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
                    const float lin = fmaf(acc, inv_a * vec[c0 + j], vec[128 + c0 + j]);
                    u[c0 + j] = fmaxf(lin, 0.f);
                    sum += u[c0 + j];
                }
            }
            const float mean = sum * (1.f / 128.f);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 128; ++j) { u[j] -= mean; q = fmaf(u[j], u[j], q); }
            const float rstd = 1.f / sqrtf(q * (1.f / 128.f) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 128; ++j) u[j] = fmaf(u[j] * rstd, vec[256 + j], vec[384 + j]);
        }
        if (row < rows) {
            float* dst = u_out + (size_t)row * kHidden;
#pragma unroll
            for (int j = 0; j < 128; j += 4) st4(dst + j, make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

static cudaError_t launch_fc_stack_tc(const float* u_in, const void* w_planes, const float* bias, const float* gamma,
                                      const float* beta, float* u_out, int rows, int depth, cudaStream_t st) {
    static bool opted = false;
    if (!opted) {
        cudaError_t e = cudaFuncSetAttribute(tc::fc_stack_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             tc::kFcSmemBytes);
        if (e != cudaSuccess) return e;
        opted = true;
    }
    const int n_pairs = ((rows + tc::kFcTile - 1) / tc::kFcTile + 1) / 2;
    const int grid = n_pairs < tc::num_sms() ? n_pairs : tc::num_sms();
    tc::fc_stack_tc_kernel<<<grid, 256, tc::kFcSmemBytes, st>>>(
        u_in, static_cast<const __half*>(w_planes), split_scales(w_planes, 128, 128), bias, gamma, beta, u_out, rows,
        depth);
    return cudaGetLastError();
}
