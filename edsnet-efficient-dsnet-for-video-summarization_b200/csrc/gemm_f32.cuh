// FP32 CUDA-core GEMM  C[M,N] = A[M,K] * B[N,K]^T (+ epilogue)  -- the exact ("fp32") precision mode of the
// projections (to_qkv transformer/nystroformer.py:82, to_out :143, fc1 anchor_based/dsnet.py:106).
// Both operands are K-major exactly as nn.Linear stores them.  128x128x16 CTA tile, 256 threads, 8x8 per thread
// (two 4-wide strips per dimension so that every smem read is a conflict-free float4).
#pragma once
#include "common.cuh"

enum GemmEpi : int {
    EPI_NONE = 0,
    EPI_QSCALE = 1,        // columns < qcols multiplied by 1/8 (q = q * dim_head^-0.5, nystroformer.py:91)
    EPI_BIAS = 2,          // + bias[n]
    EPI_BIAS_RES = 3,      // + bias[n] + res[m, n]   (to_out bias + residual x, dsnet.py:105)
    EPI_QKV_PLANES = 4,    // tcgen05 path only: q|k|v written as row-scaled fp16 hi/lo planes per (row, head), the
                           // operand format of the tensor-core attention kernels (nystrom_tc.cuh); C = hi plane base
    // tcgen05 path only, the pair that folds LayerNorm(1024) into fc1 (dsnet.py:105-106):
    //   fc1(LN(y)) = rstd (z . (W o gamma)^T - mean(z) rowsum(W o gamma)) + (W beta + b),   z = y - c (any row constant)
    EPI_RES_LNPLANES = 5,  // to_out: z = acc + bias + res - c[m] written as fp16 hi/lo planes (edsnet_split_f16 layout at C,
                           // power-of-two row scale from an a-priori bound) + per (row, 64-column slot) sum z, sum z^2
    EPI_LN_FOLD = 6        // fc1 on those planes with B = W o gamma: rstd[m] (acc - mean[m] wgsum[n]) + bias[n]
};

struct GemmEpiArgs {
    const float* bias;     // [N]
    const float* res;      // [M, ldr]
    int ldr;
    int qcols;
    const float* a_scale;  // tcgen05 path: [M] power-of-two factor undoing the row scaling of A's fp16 planes
    const float* b_scale;  // tcgen05 path: [N] same for B
    float* aux;            // EPI_QKV_PLANES: [M][24] inverse scales, index part*8 + head (q's include the 1/8)
                           // EPI_RES_LNPLANES (written) / EPI_LN_FOLD (read): [M][16][2] partial (sum z, sum z^2)
    const float* aux2;     // EPI_RES_LNPLANES: [M][2] (mean, max|.|) of the residual rows; EPI_LN_FOLD: wgsum [N]
    const float* aux3;     // EPI_RES_LNPLANES: [2] = { max_n sum_k |B[n,k]|, max_n |bias[n]| }
};

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmLd = 132;

template <int EPI>
__global__ void __launch_bounds__(256, 2)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                float* __restrict__ C, int ldc, int M, int N, int K, GemmEpiArgs ep) {
    __shared__ __align__(16) float As[2][kGemmBK][kGemmLd];
    __shared__ __align__(16) float Bs[2][kGemmBK][kGemmLd];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;

    // global -> register staging: each thread moves 2 float4 of A and 2 of B per k-tile
    const int lrow = tid >> 2;          // 0..63 (+64 for the second)
    const int lk4 = (tid & 3) * 4;      // 0,4,8,12
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r = m0 + lrow + h * 64;
            ra[h] = (r < M) ? ldg4(A + (size_t)r * lda + k0 + lk4) : make_float4(0.f, 0.f, 0.f, 0.f);
            int c = n0 + lrow + h * 64;
            rb[h] = (c < N) ? ldg4(B + (size_t)c * ldb + k0 + lk4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r = lrow + h * 64;
            As[buf][lk4 + 0][r] = ra[h].x; As[buf][lk4 + 1][r] = ra[h].y;
            As[buf][lk4 + 2][r] = ra[h].z; As[buf][lk4 + 3][r] = ra[h].w;
            Bs[buf][lk4 + 0][r] = rb[h].x; Bs[buf][lk4 + 1][r] = rb[h].y;
            Bs[buf][lk4 + 2][r] = rb[h].z; Bs[buf][lk4 + 3][r] = rb[h].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    gload(0);
    sstore(0);
    __syncthreads();
    const int nk = K / kGemmBK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * kGemmBK);
#pragma unroll
        for (int k = 0; k < kGemmBK; ++k) {
            float4 a0 = lds4(&As[buf][k][ty * 4]), a1 = lds4(&As[buf][k][64 + ty * 4]);
            float4 b0 = lds4(&Bs[buf][k][tx * 4]), b1 = lds4(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);      // other buffer: last read two iterations ago, fenced by the barrier below
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int c = n0 + jh * 64 + tx * 4;
            if (c >= N) continue;
            float4 o = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            if (EPI == EPI_QSCALE) {
                if (c < ep.qcols) { o.x *= 0.125f; o.y *= 0.125f; o.z *= 0.125f; o.w *= 0.125f; }
            }
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_RES) {
                float4 b = ldg4(ep.bias + c);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (EPI == EPI_BIAS_RES) {
                float4 x = ldg4(ep.res + (size_t)r * ep.ldr + c);
                o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
            }
            st4(C + (size_t)r * ldc + c, o);
        }
    }
}

template <int EPI>
static cudaError_t launch_sgemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                                   int M, int N, int K, GemmEpiArgs ep, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    dim3 grid((N + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM);
    sgemm_nt_kernel<EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, ep);
    return cudaGetLastError();
}
