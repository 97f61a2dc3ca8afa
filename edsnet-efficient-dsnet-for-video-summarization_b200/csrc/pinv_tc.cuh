// Iterative Moore-Penrose pseudo-inverse of attn2 (transformer/nystroformer.py:13-28) and W = Z a3v on tcgen05, ONE head
// of one video per CTA, four CTAs per SM (round 2; pinv_w_tc_kernel of attn_tc.cuh is the round-1 kernel and stays as
// the cross-check behind EDSNET_PINV_VARIANT=1).
//
// What changed against the round-1 chain, and why (ncu, profiles/r02o: the old kernel issued 630 warp instructions per
// product and warp, 150 of them selects for the "d I - X" diagonal and 60 for matrix maxima, and its two co-resident
// CTAs overlapped by only 1.27x):
//   * XZ is carried along instead of being recomputed: with T2 = 13 I - XZ (15 I - XZ (7 I - XZ)) both
//     Z' = Z T2 / 4 and XZ' = A Z' = XZ T2 / 4, so ONE product of the stacked left operand [XZ ; Z] (128 rows) with T2
//     yields both: three dependent products per iteration instead of four (20 instead of 26 per chain), and the third
//     wastes no tensor work.  In fp32 the W = Z a3v this produces sits as close to the float64 result as the
//     reference's own fp32 iteration does (4.7e-7 .. 9.2e-7 against 2.2e-7 .. 1.5e-6 on the probe cases).
//   * attn2 is row-stochastic and the start value is A^T / (|A|_1 |A|_inf), so every XZ_k = A p_k(A^T A) A^T is symmetric
//     with eigenvalues in [0, 1]: |XZ| <= 1, |7 I - XZ| <= 7, |15 I - XZ T1| <= 15, |13 I - XZ U| <= 13 entry-wise.  The
//     plane scales of all four are compile-time powers of two (2x headroom); no matrix maximum is ever reduced.  Only Z
//     needs a data-dependent scale, and Z is only ever a LEFT (K-major) operand, which may carry one scale per row: the
//     two threads of a row agree on it through shared memory.
//   * the diagonal is patched in shared memory by the one thread that owns it (read back hi + lo, add d, re-split)
//     instead of 32 selects per thread and product.
//   * 256 threads per head, 48 KB and 128 TMEM columns per CTA: four chains per SM (the TMEM limit) on 32 warps.
//
// Tiles (SWIZZLE_128B, rows of 128 B = 64 fp16, hi plane then lo plane):
//   L [128 rows]: rows 0..63 = XZ (scale 2^14; A during the first product), rows 64..127 = Z (one scale per row)
//   R [ 64 rows]: the right operand, MN-major ([k][n] row-major): Z0, then T1 / U / T2 in turn, a3v at the end
// Accumulator row r of L R: rows 0..63 = XZ R, rows 64..127 = Z R.  Thread t: row t & 127, columns 32 (t >> 7) .. +31.
#pragma once
#include "attn_tc.cuh"

namespace tc {

#ifndef PINV2_CTAS
#define PINV2_CTAS 4
#endif
constexpr int kPinv2L = 32768, kPinv2R = 16384;
constexpr int kPinv2VecBytes = 2 * 128 * 4 + 64;                            // pair exchange [2][128], reduce slots
constexpr int kPinv2SmemBytes = kPinv2L + kPinv2R + kPinv2VecBytes + 64 + 1024;

// eight values (one 16-byte chunk of each plane) of row r -> hi / lo planes, times `mul` (a power of two)
__device__ __forceinline__ void pinv2_store8(unsigned char* hi, unsigned char* lo, int r, int chunk, const float (&v)[8],
                                             float mul) {
    __half2 hh[4], ll[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float v0 = v[2 * q] * mul, v1 = v[2 * q + 1] * mul;
        const __half2 h = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(h);
        hh[q] = h;
        ll[q] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    }
    const uint32_t off = sw128_off(r, chunk);
    *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(hh);
    *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(ll);
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// This thread's 32 accumulator columns (main + cross) in four chunks of eight; the next chunk's TMEM loads are in flight
// while f(chunk, values) works on the current one.  Nothing but one chunk is ever live: the kernel fits 64 registers
// (four CTAs per SM) without spills.
template <typename F>
__device__ __forceinline__ void pinv2_stream(uint32_t t_main, uint32_t t_lo, F&& f) {
    uint32_t a[8], b[8];
    tmem_ld8_nowait(t_main, a);
    tmem_ld8_nowait(t_lo, b);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __fadd_rn(__uint_as_float(a[j]), __uint_as_float(b[j]));
        if (c < 3) {
            tmem_ld8_nowait(t_main + (uint32_t)(8 * (c + 1)), a);
            tmem_ld8_nowait(t_lo + (uint32_t)(8 * (c + 1)), b);
        }
        f(c, v);
    }
}
// element (r, r) of a tile just written by THIS thread: value += d (in plane units), re-split
__device__ __forceinline__ void pinv2_patch_diag(unsigned char* hi, unsigned char* lo, int r, float d) {
    const uint32_t off = sw128_off(r, r >> 3) + (uint32_t)((r & 7) * 2);
    __half* ph = reinterpret_cast<__half*>(hi + off);
    __half* pl = reinterpret_cast<__half*>(lo + off);
    const float x = __half2float(*ph) + __half2float(*pl) + d;
    const __half h = __float2half_rn(x);
    *ph = h;
    *pl = __float2half_rn(x - __half2float(h));
}

__device__ int g_pinv_arrivals[256];                                        // CTAs that have started on each SM (parity only)

__global__ void __launch_bounds__(256, PINV2_CTAS)
pinv_w_tc2_kernel(const float* __restrict__ attn2, const float* __restrict__ stats, const float* __restrict__ a3v,
                  float* __restrict__ w_out, float* __restrict__ z_out, int iters) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    constexpr int oL = 0, oR = kPinv2L, oVec = kPinv2L + kPinv2R;
    unsigned char* Lhi = g + oL;
    unsigned char* Llo = g + oL + 16384;
    unsigned char* Rhi = g + oR;
    unsigned char* Rlo = g + oR + 8192;
    float* s_x = reinterpret_cast<float*>(g + oVec);                        // [2 halves][128 rows]
    float* s_red = s_x + 256;                                               // [8 warps] + result
    const uint32_t barA = base + oVec + kPinv2VecBytes, barB = barA + 8;    // every product / the products the Z rows read
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(g + oVec + kPinv2VecBytes + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = tid >> 7;
    // Which half of the 128 tile rows carries XZ alternates between the CTAs that arrive on an SM: a warp can only read
    // the TMEM lane quarter (warp % 4), and warp % 4 is also its scheduler, so with a fixed assignment the XZ-side work of
    // all four co-resident chains (two thirds of all instructions) would pile up on schedulers 0 and 1.
    __shared__ int s_flip;
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_flip = atomicAdd(&g_pinv_arrivals[smid & 255u], 1) & 1;
        mbar_init(barA, 1);
        mbar_init(barB, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(base + oVec + kPinv2VecBytes + 16, 128);
    __syncthreads();
    const int flip = s_flip;
    const bool top = (row >> 6) == flip;                                    // XZ side; else Z side
    const int i = row & 63, c0 = half * 32, ch0 = half * 4;
    const int h = blockIdx.x, v = blockIdx.y;
    const size_t off = ((size_t)v * kHeads + h) * 4096;
    const bool has_diag = (i >= c0) && (i < c0 + 32);

    // ---- start: A = attn2 -> the XZ rows of L; Z0 = A^T / (max row sum x max column sum over ALL 8 heads) -> R and the Z rows ----
    float mrow = 0.f, mcol = 0.f;
#pragma unroll
    for (int q = 0; q < kHeads; ++q) {
        mrow = fmaxf(mrow, __ldg(stats + ((size_t)v * kHeads + q) * 2 + 0));
        mcol = fmaxf(mcol, __ldg(stats + ((size_t)v * kHeads + q) * 2 + 1));
    }
    const float denom = mrow * mcol;
    const int eZ0 = scale_exp(1.f / denom);                                 // |Z0| <= max(A) / denom <= 1 / denom
    float inv_z = ldexpf(1.f, -eZ0);                                        // inverse plane scale of this thread's Z row
    if (top) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 x0 = ldg4(attn2 + off + i * 64 + c0 + c * 8), x1 = ldg4(attn2 + off + i * 64 + c0 + c * 8 + 4);
            const float t[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            pinv2_store8(Lhi, Llo, row, ch0 + c, t, 16384.f);               // probabilities: fixed scale 2^14
        }
    } else {
        const float mul = ldexpf(1.f, eZ0) / denom;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldg(attn2 + off + (c0 + c * 8 + j) * 64 + i);   // row i of A^T
            pinv2_store8(Lhi, Llo, row, ch0 + c, t, mul);
            pinv2_store8(Rhi, Rlo, i, ch0 + c, t, mul);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t t_main = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, t_lo = t_main + 64u;
    uint32_t phA = 0, phB = 0;
    bool ok = true;

    auto issue = [&](bool for_z) {
        if (tid == 0) {
            tc_fence_after();
            // R's hi and lo planes are 8 KB apart = two MN-major atoms of ONE 128-column right operand: L_hi [R_hi | R_lo]
            // is a single instruction whose halves land in (main | cross); L_lo R_hi follows into the cross half.  Eight
            // instead of twelve instructions per product: the issuing thread needs ~90 cycles per tcgen05.mma (clock64
            // probe), and that issue time sits on the chain's critical path.
            constexpr uint32_t idesc2 = make_idesc_bmn(128, 128), idesc1 = make_idesc_bmn(128, 64);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ka = (uint32_t)k * 32u, kb = (uint32_t)k * 2048u;
                umma_f16(tmem_base, make_smem_desc<64>(base + oL + ka), make_smem_desc_mn2(base + oR + kb), idesc2, k != 0 ? 1u : 0u);
                umma_f16(tmem_base + 64u, make_smem_desc<64>(base + oL + 16384 + ka), make_smem_desc_mn(base + oR + kb), idesc1, 1u);
            }
            umma_commit(barA);
            if (for_z) umma_commit(barB);
        }
    };
    auto wait_a = [&]() {
        ok = mbar_wait(barA, phA) && ok;
        phA ^= 1u;
        tc_fence_after();
    };
    auto wait_b = [&]() {
        ok = mbar_wait(barB, phB) && ok;
        phB ^= 1u;
        tc_fence_after();
    };
    // The XZ-side threads have stored the next right operand: make it visible to the tensor core, then issue.  The Z-side
    // warps sleep at this barrier meanwhile (a warp parked at a hardware barrier costs no issue slots).
    auto step_sync = [&]() {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    };
    // right operand <- d I - raw * |mul| (mul carries the minus sign and the plane scale)
    auto right_from_acc = [&](float mul, float d) {
        pinv2_stream(t_main, t_lo, [&](int c, const float (&x)[8]) { pinv2_store8(Rhi, Rlo, i, ch0 + c, x, mul); });
        if (has_diag) pinv2_patch_diag(Rhi, Rlo, i, d);
    };

    // ---- XZ0 = A Z0 ; T1 = 7 I - XZ0 ----
    issue(false);
    if (top) {
        wait_a();
        const float s = ldexpf(1.f, -14 - eZ0);                             // XZ = raw s (all factors are powers of two)
        pinv2_stream(t_main, t_lo, [&](int c, const float (&x)[8]) {
            pinv2_store8(Lhi, Llo, row, ch0 + c, x, s * 16384.f);           // XZ over A (the product has completed)
            pinv2_store8(Rhi, Rlo, i, ch0 + c, x, s * -4096.f);             // T1 = 7 I - XZ <= 7: scale 2^12
        });
        if (has_diag) pinv2_patch_diag(Rhi, Rlo, i, 7.f * 4096.f);
    }
    step_sync();

    float z_unscale = 1.f;                                                  // Z = accumulator x z_unscale (Z side)
    for (int it = 0; it < iters; ++it) {            // (a timed-out wait keeps walking: every barrier stays matched)
        // U = 15 I - XZ T1
        issue(false);
        if (top) {
            wait_a();
            right_from_acc(-3.0517578125e-05f, 15.f * 2048.f);              // -2^-26 (XZ 2^14, T1 2^12) x 2^11 (U <= 15)
        }
        step_sync();
        // T2 = 13 I - XZ U
        issue(false);
        if (top) {
            wait_a();
            right_from_acc(-6.103515625e-05f, 13.f * 2048.f);               // -2^-25 (XZ 2^14, U 2^11) x 2^11 (T2 <= 13)
        }
        step_sync();
        // [XZ' ; Z'] = [XZ ; Z] T2 / 4
        issue(true);
        if (top) {
            wait_a();
            pinv2_stream(t_main, t_lo, [&](int c, const float (&x)[8]) {
                pinv2_store8(Lhi, Llo, row, ch0 + c, x, 1.220703125e-04f);  // XZ' = raw 2^-27 (2^-14 2^-11 / 4), x 2^14
                pinv2_store8(Rhi, Rlo, i, ch0 + c, x, -3.0517578125e-05f);  // next iteration's T1: -XZ' x 2^12
            });
            if (has_diag) pinv2_patch_diag(Rhi, Rlo, i, 7.f * 4096.f);
        } else {
            wait_b();
            z_unscale = inv_z * 1.220703125e-04f;                           // Z' = raw z_unscale (2^-11 / 4 = 2^-13)
            // one scale per Z row: the two threads of the row exchange their maxima
            float mx = 0.f;
            pinv2_stream(t_main, t_lo, [&](int, const float (&x)[8]) {
#pragma unroll
                for (int j = 0; j < 8; ++j) mx = fmaxf(mx, fabsf(x[j]));
            });
            s_x[half * 128 + row] = mx * z_unscale;
            named_bar_sync(2, 128);
            const int e = scale_exp(fmaxf(s_x[row], s_x[128 + row]));
            const float zmul = z_unscale * ldexpf(1.f, e);
            pinv2_stream(t_main, t_lo, [&](int c, const float (&x)[8]) { pinv2_store8(Lhi, Llo, row, ch0 + c, x, zmul); });
            inv_z = ldexpf(1.f, -e);
        }
        step_sync();
    }
    // ---- Z out (still in the accumulator), a3v -> R ----
    if (top) {
        float mx = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 x = ldg4(a3v + off + i * 64 + c0 + c * 4);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        }
        mx = warp_max(mx);
        if (lane == 0) s_red[warp] = mx;
        named_bar_sync(1, 128);
        const int e = scale_exp(fmaxf(fmaxf(s_red[2 * flip], s_red[2 * flip + 1]), fmaxf(s_red[4 + 2 * flip], s_red[5 + 2 * flip])));
        const float mul = ldexpf(1.f, e);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 x0 = ldg4(a3v + off + i * 64 + c0 + c * 8), x1 = ldg4(a3v + off + i * 64 + c0 + c * 8 + 4);
            const float t[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            pinv2_store8(Rhi, Rlo, i, ch0 + c, t, mul);
        }
        if (i == 0 && half == 0) s_red[8] = ldexpf(1.f, -e);
    } else if (z_out != nullptr) {
        if (iters > 0) {
            pinv2_stream(t_main, t_lo, [&](int c, const float (&x)[8]) {
                float* dst = z_out + off + i * 64 + c0 + c * 8;
                st4(dst, make_float4(x[0] * z_unscale, x[1] * z_unscale, x[2] * z_unscale, x[3] * z_unscale));
                st4(dst + 4, make_float4(x[4] * z_unscale, x[5] * z_unscale, x[6] * z_unscale, x[7] * z_unscale));
            });
        } else {
#pragma unroll 4
            for (int j = 0; j < 32; ++j) z_out[off + i * 64 + c0 + j] = __ldg(attn2 + off + (c0 + j) * 64 + i) / denom;
        }
    }
    step_sync();
    // ---- W = Z a3v ----
    const float inv_v = s_red[8];
    issue(true);
    if (!top) {
        wait_b();
        const float s = inv_z * inv_v;
        pinv2_stream(t_main, t_lo, [&](int c, const float (&x)[8]) {
            float* dst = w_out + off + i * 64 + c0 + c * 8;
            st4(dst, make_float4(x[0] * s, x[1] * s, x[2] * s, x[3] * s));
            st4(dst + 4, make_float4(x[4] * s, x[5] * s, x[6] * s, x[7] * s));
        });
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace tc
