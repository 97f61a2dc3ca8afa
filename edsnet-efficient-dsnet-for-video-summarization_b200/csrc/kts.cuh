// Kernel temporal segmentation on the device (first half): the shot boundaries that infer.py derives
// from the sub-sampled features before it calls the scoring path.  kts/cpd_nonlin.py:4-92 (calc_scatters + dynamic
// programme), kts/cpd_auto.py:6-33 (number of change points by a penalised objective), called from
// helpers/video_helper.py:109-126 with ncp = T - 1, vmax = 1, lmin = 1, lmax = 100000.
//
// Arithmetic follows the reference operation by operation, so that the change points come out bit-identical for the
// same kernel matrix: the 2-D prefix sums of K run sequentially in float32 (np.cumsum over a float32 array, axis 0 then
// axis 1), everything else in float64; the DP adds J[t, l-1] + I[k-1, t] once and takes the FIRST minimum (np.argmin).
//
// Per video (n = T frames), in the caller's scratch: C [n][n] float32 (K, then its prefix sums in place),
// Jt [n][n] float64 (scatters, TRANSPOSED: Jt[j][i] = J[i][j], the DP then streams rows), P [n][n+1] uint16 (previous
// change point per (k, l)), K1 [n+1] float64, scores [n] float64.
#pragma once
#include "common.cuh"

struct KtsVideo {
    long long off;     // byte offset of this video's scratch block
    int row0;          // first packed feature row
    int n;             // frames
};

__host__ __device__ inline size_t kts_align(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline size_t kts_off_c(int n) { (void)n; return 0; }
__host__ __device__ inline size_t kts_off_jt(int n) { return kts_align((size_t)n * n * 4); }
__host__ __device__ inline size_t kts_off_p(int n) { return kts_off_jt(n) + kts_align((size_t)n * n * 8); }
__host__ __device__ inline size_t kts_off_k1(int n) { return kts_off_p(n) + kts_align((size_t)n * (n + 1) * 2); }
__host__ __device__ inline size_t kts_off_sc(int n) { return kts_off_k1(n) + kts_align((size_t)(n + 1) * 8); }
__host__ __device__ inline size_t kts_video_bytes(int n) { return kts_off_sc(n) + kts_align((size_t)(n + 1) * 8); }

// K = X X^T in float32 (np.matmul(features, features.T), video_helper.py:117), 16 x 16 output tile per CTA.
// (The reference's BLAS sums in an unspecified order, so K itself is only reproducible to float32 rounding.)
__global__ void __launch_bounds__(256)
kts_gram_kernel(const float* __restrict__ x, const KtsVideo* __restrict__ vids, unsigned char* __restrict__ scratch) {
    __shared__ float sa[16][33], sb[16][33];
    const KtsVideo v = vids[blockIdx.z];
    const int n = v.n, i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    if (i0 >= n || j0 >= n) return;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc = 0.f;
    for (int k0 = 0; k0 < kFeat; k0 += 32) {
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            sa[r][c] = i0 + r < n ? __ldg(x + (size_t)(v.row0 + i0 + r) * kFeat + k0 + c) : 0.f;
            sb[r][c] = j0 + r < n ? __ldg(x + (size_t)(v.row0 + j0 + r) * kFeat + k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc = fmaf(sa[ty][c], sb[tx][c], acc);
        __syncthreads();
    }
    if (i0 + ty < n && j0 + tx < n)
        reinterpret_cast<float*>(scratch + v.off + kts_off_c(n))[(size_t)(i0 + ty) * n + j0 + tx] = acc;
}

// K1 = cumsum([0] + diag(K)) in float64, then C = cumsum(cumsum(K, axis 0), axis 1) in float32, sequentially.
// One CTA per video: phase 1 thread <-> column (rows in order), phase 2 thread <-> row (columns in order).
__global__ void __launch_bounds__(256)
kts_prefix_kernel(const KtsVideo* __restrict__ vids, unsigned char* __restrict__ scratch) {
    const KtsVideo v = vids[blockIdx.x];
    const int n = v.n;
    float* C = reinterpret_cast<float*>(scratch + v.off + kts_off_c(n));
    double* K1 = reinterpret_cast<double*>(scratch + v.off + kts_off_k1(n));
    if (threadIdx.x == 0) {
        double s = 0.0;
        K1[0] = 0.0;
        for (int i = 0; i < n; ++i) { s = __dadd_rn(s, (double)C[(size_t)i * n + i]); K1[i + 1] = s; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += 256) {
        float s = 0.f;
        for (int r = 0; r < n; ++r) { s = __fadd_rn(s, C[(size_t)r * n + c]); C[(size_t)r * n + c] = s; }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += 256) {
        float s = 0.f;
        for (int c = 0; c < n; ++c) { s = __fadd_rn(s, C[(size_t)r * n + c]); C[(size_t)r * n + c] = s; }
    }
}

// scatters (cpd_nonlin.py:17-25), stored transposed.  K2[a][b] = a && b ? C[a-1][b-1] : 0.
__global__ void __launch_bounds__(256)
kts_scatter_kernel(const KtsVideo* __restrict__ vids, unsigned char* __restrict__ scratch) {
    const KtsVideo v = vids[blockIdx.y];
    const int n = v.n;
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= (long long)n * n) return;
    const int j = (int)(e / n), i = (int)(e % n);           // consecutive threads: consecutive i (the Jt row index j)
    const float* C = reinterpret_cast<const float*>(scratch + v.off + kts_off_c(n));
    const double* K1 = reinterpret_cast<const double*>(scratch + v.off + kts_off_k1(n));
    double* Jt = reinterpret_cast<double*>(scratch + v.off + kts_off_jt(n));
    double val = 0.0;
    if (j >= i) {
        auto K2 = [&](int a, int b) -> double { return (a > 0 && b > 0) ? (double)C[(size_t)(a - 1) * n + (b - 1)] : 0.0; };
        const double d = __dsub_rn(__dsub_rn(__dadd_rn(K2(j + 1, j + 1), K2(i, i)), K2(j + 1, i)), K2(i, j + 1));
        const float len = __fadd_rn((float)(j - i + 1), (j == i - 1) ? 1.f : 0.f);
        val = __dsub_rn(__dsub_rn(K1[j + 1], K1[i]), __ddiv_rn(d, (double)len));
    }
    Jt[(size_t)j * n + i] = val;
}

// Dynamic programme over k = 0 .. m change points (cpd_nonlin.py:61-78) with the previous-change table kept for every
// k, then cpd_auto's choice of m_best (cpd_auto.py:21-30) and the back-tracking from (m_best, n).  One CTA per video,
// 1024 threads: a warp takes one end position l at a time, lanes stride over the candidate t (coalesced along the
// transposed scatter row), first-minimum reduction.  m_fixed < 0: cpd_auto; otherwise exactly m_fixed change points.
__global__ void __launch_bounds__(1024)
kts_dp_kernel(const KtsVideo* __restrict__ vids, unsigned char* __restrict__ scratch, int ncp_cap, int m_fixed,
              double vmax, int desc_rate, int lmin, int lmax, const int* __restrict__ cu_rows,
              int* __restrict__ n_cps, int* __restrict__ cps_out, double* __restrict__ obj_out) {
    extern __shared__ double s_I[];                          // [2][n + 1]
    const int vi = blockIdx.x;
    const KtsVideo v = vids[vi];
    const int n = v.n;
    const double* Jt = reinterpret_cast<const double*>(scratch + v.off + kts_off_jt(n));
    unsigned short* P = reinterpret_cast<unsigned short*>(scratch + v.off + kts_off_p(n));
    double* scores = reinterpret_cast<double*>(scratch + v.off + kts_off_sc(n));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int m = m_fixed >= 0 ? m_fixed : (ncp_cap >= 0 ? min(ncp_cap, n - 1) : n - 1);
    // feasibility of the reference's assertion (m + 1) * lmin <= n <= (m + 1) * lmax is checked by the host wrapper
    const double BIG = 1e101;
    double* Ia = s_I;
    double* Ib = s_I + (n + 1);
    for (int l = tid; l <= n; l += 1024) {
        const bool in = l >= lmin && l < lmax;               // I[0, lmin:lmax] = J[0, lmin-1:lmax-1]
        Ia[l] = in ? Jt[(size_t)(l - 1) * n + 0] : BIG;
    }
    __syncthreads();
    if (tid == 0) scores[0] = Ia[n];
    for (int k = 1; k <= m; ++k) {
        for (int l = tid; l <= n; l += 1024) Ib[l] = BIG;
        __syncthreads();
        for (int l = (k + 1) * lmin + warp; l <= n; l += 32) {
            const int t0 = max(k * lmin, l - lmax), t1 = l - lmin + 1;
            const double* jrow = Jt + (size_t)(l - 1) * n;
            double best = INFINITY;
            int bi = 0x7fffffff;
            for (int t = t0 + lane; t < t1; t += 32) {
                const double c = __dadd_rn(jrow[t], Ia[t]);
                if (c < best) { best = c; bi = t; }          // ascending t: the first minimum wins
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                Ib[l] = best;
                P[(size_t)k * (n + 1) + l] = (unsigned short)bi;
            }
        }
        __syncthreads();
        if (tid == 0) scores[k] = Ib[n];
        double* t = Ia; Ia = Ib; Ib = t;
        __syncthreads();
    }
    if (tid == 0) {
        int m_best = m;
        if (m_fixed < 0) {
            // costs = scores / N + penalties, first minimum (cpd_auto.py:21-30); scores above 1e99 count as infinity
            const double N = (double)n, N2 = (double)(n * desc_rate);
            double best = INFINITY;
            m_best = 0;
            for (int q = 0; q <= m; ++q) {
                double sc = scores[q];
                if (sc > 1e99) sc = INFINITY;
                double pen = 0.0;
                if (q > 0) pen = __dmul_rn(__ddiv_rn(__dmul_rn(vmax, (double)q), __dmul_rn(2.0, N2)),
                                           __dadd_rn(log(__ddiv_rn(N2, (double)q)), 1.0));
                const double cost = __dadd_rn(__ddiv_rn(sc, N), pen);
                if (cost < best) { best = cost; m_best = q; }
            }
        }
        n_cps[vi] = m_best;
        int cur = n;
        int* out = cps_out + cu_rows[vi];                    // at most n - 1 entries per video
        for (int k = m_best; k >= 1; --k) {
            cur = (int)P[(size_t)k * (n + 1) + cur];
            out[k - 1] = cur;
        }
        if (obj_out) {                                       // objective for 0 .. m_best change points (scores2)
            double* oo = obj_out + cu_rows[vi];
            for (int q = 0; q <= m_best && q < n; ++q) oo[q] = scores[q] > 1e99 ? INFINITY : scores[q];
        }
    }
}
