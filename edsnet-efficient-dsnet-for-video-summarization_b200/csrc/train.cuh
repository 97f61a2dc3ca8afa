// Training step of the anchor-based path (BASELINE.json config 3): what anchor_based/train.py:110-128 runs per video --
// model(seq) in train() mode, calc_cls_loss + calc_loc_loss (anchor_based/losses.py), loss.backward(), Adam -- as CUDA
// kernels.  The forward re-uses the inference kernels (the attention block has dropout 0) up to fc1; the shared fc block
// runs with Dropout(0.5) (fc_stack_tc.cuh, SAVE variant).  This file holds the backward:
//
//   loss_grad_kernel        d loss / d logits, d loss / d pred_loc in closed form (no division by p or 1 - p)
//   roi_heads_bwd_kernel    transposed multi-scale window sums: d(u . w_head) per feature row, head bias gradients
//   fc_stack_bwd_kernel     D shared blocks in reverse (LayerNorm, dropout / ReLU mask, d input = da W on CUDA cores),
//                           head weight / LayerNorm / bias gradients; leaves da and the block inputs for the dW GEMM
//   ln1024_bwd_kernel       LayerNorm(1024) backward + column sums (d gamma, d beta, d to_out.bias)
//   split_t_kernel          fp32 [rows][cols] -> TRANSPOSED row-scaled fp16 hi/lo planes [cols][rows padded to 64]: the
//                           operand format of the tcgen05 GEMM (gemm_tc.cuh) for every dW = dY^T X (contraction over the
//                           rows) and, applied to a weight, for every dX = dY W
//   attention backward      a3_stats / attn_bwd_rows / pinv_bwd / attn2_bwd / attn_bwd_keys / dqkv_finish: the Nystrom
//                           block (transformer/nystroformer.py:95-142) in reverse, fp32 on CUDA cores, 64 x 64 tiles
//   adam_kernel             torch.optim.Adam(lr, weight_decay) on flat parameter / gradient / moment buffers
//
// All dense dW / dX products of the three projections and of the fc block run on tcgen05 through gemm_dispatch with three
// split-fp16 passes (fp32-grade: the gradients are held to 1e-4 against the reference's autograd).
// oracle/backward_model.py restates every kernel below in torch ops; tests/test_backward_model.py pins that model to
// torch.autograd, the GPU tests compare these kernels with it stage by stage.
#pragma once
#include "common.cuh"
#include "attn_tc.cuh"
#include "tail.cuh"

__global__ void __launch_bounds__(256)
dropout_mask_kernel(unsigned long long seed, unsigned long long offset, int rows, int depth, uint8_t* __restrict__ out) {
    const int idx = blockIdx.x * 256 + threadIdx.x;                 // one thread per (layer, row)
    if (idx >= rows * depth) return;
    const int layer = idx / rows, row = idx - layer * rows;
    const uint4 w = dropout_words(seed, offset, row, layer);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    uint8_t* dst = out + (size_t)idx * 128;
    for (int c = 0; c < 128; ++c) dst[c] = (uint8_t)((ww[c >> 5] >> (c & 31)) & 1u);
}

// ---------------------------------------------------------------------------------------------------------
// 64 x 64 x 64 products on smem tiles (row stride kLd64), 256 threads, tid = ty * 16 + tx.
//   NN: C[i][j] += sum_k A[i][k] B[k][j]      thread owns rows ty*4+i, columns tx*4+j        (common.cuh mm64_acc)
//   TN: C[i][j] += sum_k A[k][i] B[k][j]      same ownership
//   NT: C[i][j] += sum_k A[i][k] B[j][k]      thread owns rows ty*4+i, columns tx+16*j  (rows tx, tx+16, ... of B: the
//                                             float4 reads of a quarter warp then fall into 8 different bank groups)
// ---------------------------------------------------------------------------------------------------------
enum MmMode : int { MM_NN = 0, MM_TN = 1, MM_NT = 2 };

template <int MODE>
__device__ __forceinline__ int mm_col(int tx, int j) { return MODE == MM_NT ? tx + 16 * j : tx * 4 + j; }

// RT = rows per thread: 4 (256 threads, ty = tid >> 4 in 0..15) or 2 (512 threads, ty in 0..31); thread owns rows ty*RT+i
template <int MODE, int RT = 4>
__device__ __forceinline__ void mm64(float (&acc)[RT][4], const float* __restrict__ A, const float* __restrict__ B,
                                     int ty, int tx) {
    if (MODE == MM_NN) {
#pragma unroll 4
        for (int k4 = 0; k4 < 64; k4 += 4) {
            float4 a[RT], b[4];
#pragma unroll
            for (int i = 0; i < RT; ++i) a[i] = lds4(A + (ty * RT + i) * kLd64 + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) b[kk] = lds4(B + (k4 + kk) * kLd64 + tx * 4);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    acc[i][0] = fmaf(av[kk], b[kk].x, acc[i][0]); acc[i][1] = fmaf(av[kk], b[kk].y, acc[i][1]);
                    acc[i][2] = fmaf(av[kk], b[kk].z, acc[i][2]); acc[i][3] = fmaf(av[kk], b[kk].w, acc[i][3]);
                }
            }
        }
    } else if (MODE == MM_TN) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            float av[RT];
            if (RT == 4) {
                const float4 a = lds4(A + k * kLd64 + ty * 4);
                av[0] = a.x; av[1] = a.y; av[RT - 2] = a.z; av[RT - 1] = a.w;
            } else {
                const float2 a = *reinterpret_cast<const float2*>(A + k * kLd64 + ty * 2);
                av[0] = a.x; av[1] = a.y;
            }
            const float4 b = lds4(B + k * kLd64 + tx * 4);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
            }
        }
    } else {
#pragma unroll 4
        for (int k4 = 0; k4 < 64; k4 += 4) {
            float4 a[RT], b[4];
#pragma unroll
            for (int i = 0; i < RT; ++i) a[i] = lds4(A + (ty * RT + i) * kLd64 + k4);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = lds4(B + (tx + 16 * j) * kLd64 + k4);
#pragma unroll
            for (int i = 0; i < RT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] = fmaf(a[i].x, b[j].x, fmaf(a[i].y, b[j].y, fmaf(a[i].z, b[j].z, fmaf(a[i].w, b[j].w, acc[i][j]))));
        }
    }
}
template <int RT>
__device__ __forceinline__ void zero_acc(float (&acc)[RT][4]) {
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}
// C = alpha * acc + beta * C (+ diag on the diagonal), in the ownership of MODE
template <int MODE, int RT = 4>
__device__ __forceinline__ void mm64_store(float* __restrict__ C, const float (&acc)[RT][4], float alpha, float beta,
                                           float diag, int ty, int tx) {
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int r = ty * RT + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = mm_col<MODE>(tx, j);
            float v = alpha * acc[i][j];
            if (beta != 0.f) v = fmaf(beta, C[r * kLd64 + c], v);
            if (diag != 0.f && r == c) v += diag;
            C[r * kLd64 + c] = v;
        }
    }
}
// C = alpha * op(A, B) + beta * C + diag I; the caller places the barriers
template <int MODE, int RT = 4>
__device__ __forceinline__ void mm64_to(float* __restrict__ C, const float* __restrict__ A, const float* __restrict__ B,
                                        float alpha, float beta, float diag, int ty, int tx) {
    float acc[RT][4];
    zero_acc<RT>(acc);
    mm64<MODE, RT>(acc, A, B, ty, tx);
    mm64_store<MODE, RT>(C, acc, alpha, beta, diag, ty, tx);
}
// dense 64 x 64 fp32 matrix in global memory <-> smem tile, any block size
__device__ __forceinline__ void tile_load(float* __restrict__ S, const float* __restrict__ G, int tid, int nthr) {
    for (int idx = tid; idx < 1024; idx += nthr) st4(S + (idx >> 4) * kLd64 + (idx & 15) * 4, ldg4(G + idx * 4));
}
__device__ __forceinline__ void tile_store(float* __restrict__ G, const float* __restrict__ S, int tid, int nthr) {
    for (int idx = tid; idx < 1024; idx += nthr) st4(G + idx * 4, lds4(S + (idx >> 4) * kLd64 + (idx & 15) * 4));
}
// 64 x 64 global tile (row stride ld_g, rows >= n_valid read as zero) -> smem
__device__ __forceinline__ void load64_rows(float* __restrict__ S, const float* __restrict__ G, int ld_g, int n_valid, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * 256;
        const int r = idx >> 4, c4 = (idx & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_valid) v = ldg4(G + (size_t)r * ld_g + c4);
        st4(S + r * kLd64 + c4, v);
    }
}
// smem tile -> atomicAdd into a dense 64 x 64 global matrix
__device__ __forceinline__ void atomic_add64(float* __restrict__ G, const float* __restrict__ S, int tid) {
    for (int idx = tid; idx < 4096; idx += 256) atomicAdd(G + idx, S[(idx >> 6) * kLd64 + (idx & 63)]);
}

// ---------------------------------------------------------------------------------------------------------
// split_t: fp32 src [rows][cols] (row stride ld) -> planes of the TRANSPOSE, hi [cols][kp] | lo [cols][kp] fp16 and
// inv [cols] fp32 (kp = rows rounded up to 64, the padding is zero; layout of edsnet_split_f16 with rows <-> cols): every
// output row (= source column) is scaled by a power of two that puts its largest magnitude into [2^14, 2^15).
// Up to four independent jobs per launch, two passes (column maxima, then 32-column x 64-row tiles), both spread over the
// whole GPU whatever the shape (a [rows][128] operand has only four column blocks).
// ---------------------------------------------------------------------------------------------------------
struct SplitTJob {
    const float* src;
    __half* hi;
    __half* lo;
    float* inv;
    unsigned* cmax;                // [cols] bit patterns of max|column| (non-negative floats order like unsigned), ZERO on entry
    int rows, cols, ld, kp;
    int cta0;                      // first column block (32 columns) of this job in the launch
};
struct SplitTJobs { int n; SplitTJob j[4]; };

__device__ __forceinline__ const SplitTJob& split_t_pick(const SplitTJobs& jobs, int block) {
    int ji = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q)
        if (q < jobs.n && block >= jobs.j[q].cta0) ji = q;
    return jobs.j[ji];
}

// pass 1: column maxima.  grid (column blocks of all jobs, row chunks of 256), 256 threads.
__global__ void __launch_bounds__(256)
split_t_colmax_kernel(const SplitTJobs jobs) {
    __shared__ float cmax[8][32];
    const SplitTJob& jb = split_t_pick(jobs, (int)blockIdx.x);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = ((int)blockIdx.x - jb.cta0) * 32 + lane;
    const int r0 = (int)blockIdx.y * 256, r1 = min(jb.rows, r0 + 256);
    if (r0 >= jb.rows) return;
    float mx = 0.f;
    if (c < jb.cols)
        for (int r = r0 + warp; r < r1; r += 8) mx = fmaxf(mx, fabsf(__ldg(jb.src + (size_t)r * jb.ld + c)));
    cmax[warp][lane] = mx;
    __syncthreads();
    if (tid < 32 && c < jb.cols) {
        float m = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) m = fmaxf(m, cmax[w][lane]);
        if (m > 0.f && m < INFINITY) atomicMax(jb.cmax + c, __float_as_uint(m));
    }
}

// pass 2: one CTA per (32 source columns, 64 source rows): scale, transpose through shared memory, write both planes.
// grid (column blocks of all jobs, largest kp / 64), 256 threads.
__global__ void __launch_bounds__(256)
split_t_tiles_kernel(const SplitTJobs jobs) {
    __shared__ float tile[32][65];
    const SplitTJob& jb = split_t_pick(jobs, (int)blockIdx.x);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = ((int)blockIdx.x - jb.cta0) * 32;
    const int c = c0 + lane;
    const int k0 = (int)blockIdx.y * 64;
    if (k0 >= jb.kp) return;
    const bool cin = c < jb.cols;
    const int e = tc::scale_exp(cin ? __uint_as_float(jb.cmax[c]) : 0.f);
    const float sc = ldexpf(1.f, e);
    if (blockIdx.y == 0 && warp == 0 && cin) jb.inv[c] = ldexpf(1.f, -e);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int r = k0 + warp * 8 + rr;
        float v = 0.f;
        if (cin && r < jb.rows) v = __ldg(jb.src + (size_t)r * jb.ld + c) * sc;
        tile[lane][warp * 8 + rr] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int cc = warp * 4 + q;
        if (c0 + cc < jb.cols) {
            const float v0 = tile[cc][2 * lane], v1 = tile[cc][2 * lane + 1];
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const size_t o = (size_t)(c0 + cc) * jb.kp + k0 + 2 * lane;
            *reinterpret_cast<__half2*>(jb.hi + o) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2*>(jb.lo + o) =
                __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
        }
    }
}

// q | k | v operand planes (gemm_tc.cuh EPI_QKV_PLANES) -> fp32 [R][1536]; q arrives pre-scaled by 1/8 (its slot scale
// carries it), which is how every backward kernel below wants it.
__global__ void __launch_bounds__(256)
qkv_planes_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, const float* __restrict__ inv,
                         float* __restrict__ out, int rows) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;      // one thread per 4 columns
    if (idx >= (size_t)rows * (kQkvCols / 4)) return;
    const size_t row = idx / (kQkvCols / 4);
    const int c4 = (int)(idx - row * (kQkvCols / 4)) * 4;
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + row * kQkvCols + c4));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(lo + row * kQkvCols + c4));
    const float s = __ldg(inv + row * 24 + (c4 >> 6));
    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
    st4(out + row * kQkvCols + c4, make_float4((h0.x + l0.x) * s, (h0.y + l0.y) * s, (h1.x + l1.x) * s, (h1.y + l1.y) * s));
}

// ---------------------------------------------------------------------------------------------------------
// Losses of one video (anchor_based/losses.py:5-57 as anchor_based/train.py:119-123 combines them) and their gradient
// with respect to the LOGITS and the offsets.  label: 1 positive, -1 negative, 0 ignored.
//   cls = 0.5 (mean_pos -log p + mean_neg -log(1 - p))     d/dlogit: -0.5 (1 - p) / n_pos | 0.5 p / n_neg
//   loc = smooth-L1 mean over the 2 n_pos offsets of the positives      d/dloc: clamp(d, -1, 1) / (2 n_pos)
// `scale` multiplies every gradient (1 / videos of the step).  A class without members contributes nothing (the
// reference's mean over an empty selection is NaN; its training loop never gets there: train.py:83-84 skips empty
// targets).  grid = videos, 256 threads.  loss_out [V][3] = {loss, cls, loc} (unscaled).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(256)
loss_grad_kernel(const float* __restrict__ pred_cls, const float* __restrict__ pred_loc, const int* __restrict__ cls_label,
                 const float* __restrict__ loc_label, const int* __restrict__ cu_rows, int S, float lambda_reg, float scale,
                 float* __restrict__ d_logit, float* __restrict__ d_loc, float* __restrict__ loss_out) {
    __shared__ float red[8];
    const int v = blockIdx.x, tid = threadIdx.x;
    const size_t a0 = (size_t)cu_rows[v] * S;
    const int n = (cu_rows[v + 1] - cu_rows[v]) * S;
    float npos = 0.f, nneg = 0.f, lpos = 0.f, lneg = 0.f, lloc = 0.f;
    for (int i = tid; i < n; i += 256) {
        const int lab = cls_label[a0 + i];
        const float p = pred_cls[a0 + i];
        if (lab == 1) {
            npos += 1.f;
            lpos -= logf(p);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float d = fabsf(pred_loc[(a0 + i) * 2 + q] - loc_label[(a0 + i) * 2 + q]);
                lloc += d < 1.f ? 0.5f * d * d : d - 0.5f;
            }
        } else if (lab == -1) {
            nneg += 1.f;
            lneg -= logf(1.f - p);
        }
    }
    npos = block_sum256(npos, red);
    nneg = block_sum256(nneg, red);
    lpos = block_sum256(lpos, red);
    lneg = block_sum256(lneg, red);
    lloc = block_sum256(lloc, red);
    const float ipos = npos > 0.f ? 1.f / npos : 0.f, ineg = nneg > 0.f ? 1.f / nneg : 0.f;
    if (tid == 0) {
        const float cls = 0.5f * (lpos * ipos + lneg * ineg), loc = lloc * 0.5f * ipos;
        loss_out[v * 3 + 0] = cls + lambda_reg * loc;
        loss_out[v * 3 + 1] = cls;
        loss_out[v * 3 + 2] = loc;
    }
    for (int i = tid; i < n; i += 256) {
        const int lab = cls_label[a0 + i];
        const float p = pred_cls[a0 + i];
        float g = 0.f, g0 = 0.f, g1 = 0.f;
        if (lab == 1) {
            g = -0.5f * (1.f - p) * ipos * scale;
            const float d0 = pred_loc[(a0 + i) * 2] - loc_label[(a0 + i) * 2];
            const float d1 = pred_loc[(a0 + i) * 2 + 1] - loc_label[(a0 + i) * 2 + 1];
            g0 = lambda_reg * fminf(fmaxf(d0, -1.f), 1.f) * 0.5f * ipos * scale;
            g1 = lambda_reg * fminf(fmaxf(d1, -1.f), 1.f) * 0.5f * ipos * scale;
        } else if (lab == -1) {
            g = 0.5f * p * ineg * scale;
        }
        d_logit[a0 + i] = g;
        d_loc[(a0 + i) * 2] = g0;
        d_loc[(a0 + i) * 2 + 1] = g1;
    }
}

// d pred_cls (after the sigmoid) -> d logit, for gradients that arrive from a loss written in torch ops
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ pred_cls, const float* __restrict__ d_cls, float* __restrict__ d_logit, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) { const float p = pred_cls[i]; d_logit[i] = d_cls[i] * p * (1.f - p); }
}

// ---------------------------------------------------------------------------------------------------------
// ROI pooling + heads, backward (anchor_based/dsnet.py:110-115).  Forward: pre[t][s][c] = (1/s) sum_{j=t-s/2}^{t+s/2-1}
// d[j][c] + b_c with d[j] = u[j] . {w_cls, w_loc0, w_loc1}.  Hence
//   g[j][c] = d loss / d d[j][c] = sum_s (1/s) sum_{t = j - s/2 + 1}^{j + s/2} dpre[t][s][c]     (t inside the video)
//   d b_c   = sum_{t, s} dpre[t][s][c]
// One CTA per 128-row tile of one video (+ halo); writes g [R][4]; the bias gradients are added atomically.
// ---------------------------------------------------------------------------------------------------------
constexpr int kRoiBwdRows = 128 + 2 * kRoiMaxHalo;

__global__ void __launch_bounds__(256)
roi_heads_bwd_kernel(const float* __restrict__ d_logit, const float* __restrict__ d_loc, const int* __restrict__ cu_rows,
                     const int2* __restrict__ tiles, ScaleList scales, int halo, float* __restrict__ g_out,
                     float* __restrict__ d_cls_b, float* __restrict__ d_loc_b) {
    extern __shared__ float sd[];                                   // [S][3][kRoiBwdRows + 1]
    __shared__ float red[8];
    const int tid = threadIdx.x;
    const int2 tile = tiles[blockIdx.x];
    const VidInfo vi = vid_info(cu_rows, tile.x);
    const int t0 = tile.y, S = scales.n;
    const int nrows = 128 + 2 * halo;
    constexpr int ldr = kRoiBwdRows + 1;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    for (int idx = tid; idx < nrows * S; idx += 256) {
        const int r = idx / S, si = idx - r * S;
        const int t = t0 - halo + r;
        float a = 0.f, l0 = 0.f, l1 = 0.f;
        if (t >= 0 && t < vi.T) {
            const size_t o = (size_t)(vi.row0 + t) * S + si;
            a = d_logit[o];
            l0 = d_loc[o * 2];
            l1 = d_loc[o * 2 + 1];
            if (r >= halo && r < halo + 128) { b0 += a; b1 += l0; b2 += l1; }
        }
        sd[(si * 3 + 0) * ldr + r] = a;
        sd[(si * 3 + 1) * ldr + r] = l0;
        sd[(si * 3 + 2) * ldr + r] = l1;
    }
    b0 = block_sum256(b0, red);
    b1 = block_sum256(b1, red);
    b2 = block_sum256(b2, red);
    if (tid == 0) { atomicAdd(d_cls_b, b0); atomicAdd(d_loc_b, b1); atomicAdd(d_loc_b + 1, b2); }
    __syncthreads();
    for (int idx = tid; idx < 128 * 4; idx += 256) {
        const int i = idx >> 2, c = idx & 3;
        const int t = t0 + i;
        if (t >= vi.T) continue;
        float acc = 0.f;
        if (c < 3) {
            for (int si = 0; si < S; ++si) {
                const int sc = scales.s[si];
                const float* p = sd + (si * 3 + c) * ldr + (i + halo - sc / 2 + 1);
                float a = 0.f;
                for (int j = 0; j < sc; ++j) a += p[j];
                acc += a / (float)sc;
            }
        }
        g_out[(size_t)(vi.row0 + t) * 4 + c] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Shared fc block x D, backward (anchor_based/dsnet.py:91-96,107-108), plus the head weights.
// Saved by the forward: h_l = Dropout(ReLU(Linear(u_l)))  [D][R][128] (before the LayerNorm; h > 0 <=> kept AND a > 0),
// u_0 = fc1 output in uin[0].  Per 64-row tile, layers D-1 .. 0:
//   hh = (h - mean) rstd;  u_{l+1} = hh gamma + beta (written to uin[l+1] for the dW GEMM; l = D-1: the final hidden rows)
//   d gamma += do hh;  d beta += do;  dhh = do gamma;  dh = rstd (dhh - mean(dhh) - hh mean(dhh hh))
//   da = dh keep_scale [h > 0];  d bias += da;  da -> das[l];  do <- da W  (64 x 128 x 128 on CUDA cores, W in smem)
// Prologue (first iteration): do = g . W_heads, d W_heads += g^T u_D.  Epilogue: du0 = do, d fc1.bias += column sums.
// Small gradients are reduced per CTA and added atomically.  grid = ceil(R / 64), 256 threads.
// ---------------------------------------------------------------------------------------------------------
constexpr int kFcBwdSmem = (128 * kLd128 + 64 * kLd128 + 16 * 128 + 6 * 128) * (int)sizeof(float);

struct FcBwdGrads {
    float* fcb_b;       // [128]
    float* fcb_ln_w;    // [128]
    float* fcb_ln_b;    // [128]
    float* fc1_b;       // [128]
    float* cls_w;       // [128]
    float* loc_w;       // [2][128]
};

__device__ __forceinline__ void fc_colsum_flush(float (&part)[8], float* __restrict__ red, float* __restrict__ dst,
                                                int ty, int tx, int tid) {
    // part[j]: this thread's partial column sums (its 4 rows) of columns tx*4+j (j < 4) and 64+tx*4+(j-4)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty * 128 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4))] = part[j];
    __syncthreads();
    if (tid < 128) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) s += red[q * 128 + tid];
        atomicAdd(dst + tid, s);
    }
}

__global__ void __launch_bounds__(256)
fc_stack_bwd_kernel(const float* __restrict__ g, const float* __restrict__ hs, float* __restrict__ uin,
                    float* __restrict__ das, float* __restrict__ du0, const float* __restrict__ w,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ w_cls,
                    const float* __restrict__ w_loc, int rows, int depth, float keep_scale, FcBwdGrads gr) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                       // [n][k] = w[n][k] (row-major as nn.Linear stores it)
    float* Da = Ws + 128 * kLd128;          // [row][n]
    float* red = Da + 64 * kLd128;          // [16][128]
    float* gs = red + 16 * 128;             // gamma
    float* es = gs + 128;                   // beta
    float* wh = es + 128;                   // w_cls | w_loc0 | w_loc1  (3 x 128), one spare row
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int r0 = blockIdx.x * 64;
    for (int idx = tid; idx < 128 * 32; idx += 256) {
        const int n = idx >> 5, k4 = (idx & 31) * 4;
        st4(Ws + n * kLd128 + k4, ldg4(w + n * 128 + k4));
    }
    if (tid < 128) {
        gs[tid] = __ldg(gamma + tid);
        es[tid] = __ldg(beta + tid);
        wh[tid] = __ldg(w_cls + tid);
        wh[128 + tid] = __ldg(w_loc + tid);
        wh[256 + tid] = __ldg(w_loc + 128 + tid);
    }
    __syncthreads();
    int col[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) col[j] = j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4);
    // do = g . W_heads
    float dout[4][8];
    float gv[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + ty * 4 + i;
        float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) gg = ldg4(g + (size_t)row * 4);
        gv[i][0] = gg.x; gv[i][1] = gg.y; gv[i][2] = gg.z;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            dout[i][j] = fmaf(gg.x, wh[col[j]], fmaf(gg.y, wh[128 + col[j]], gg.z * wh[256 + col[j]]));
    }
    float p_gam[8], p_bet[8], p_b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { p_gam[j] = 0.f; p_bet[j] = 0.f; p_b[j] = 0.f; }

    for (int l = depth - 1; l >= 0; --l) {
        float hv[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + ty * 4 + i;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (row < rows) {
                const float* src = hs + ((size_t)l * rows + row) * kHidden;
                a = ldg4(src + tx * 4);
                b = ldg4(src + 64 + tx * 4);
            }
            hv[i][0] = a.x; hv[i][1] = a.y; hv[i][2] = a.z; hv[i][3] = a.w;
            hv[i][4] = b.x; hv[i][5] = b.y; hv[i][6] = b.z; hv[i][7] = b.w;
        }
        float da[4][8];
        float p_wh[3][8];
        if (l == depth - 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) p_wh[c][j] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + ty * 4 + i;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += hv[i][j];
            const float mean = half_warp_sum(s) * (1.f / 128.f);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = hv[i][j] - mean; q = fmaf(d, d, q); }
            const float rstd = 1.f / sqrtf(half_warp_sum(q) * (1.f / 128.f) + 1e-5f);
            float hh[8], dhh[8];
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hh[j] = (hv[i][j] - mean) * rstd;
                dhh[j] = dout[i][j] * gs[col[j]];
                m1 += dhh[j];
                m2 = fmaf(dhh[j], hh[j], m2);
                p_gam[j] = fmaf(dout[i][j], hh[j], p_gam[j]);
                p_bet[j] += dout[i][j];
            }
            // the block's output rows: input of the next block (dW GEMM operand) / final hidden rows (head weights)
            if (row < rows) {
                float un[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) un[j] = fmaf(hh[j], gs[col[j]], es[col[j]]);
                if (l + 1 < depth) {
                    float* dst = uin + ((size_t)(l + 1) * rows + row) * kHidden;
                    st4(dst + tx * 4, make_float4(un[0], un[1], un[2], un[3]));
                    st4(dst + 64 + tx * 4, make_float4(un[4], un[5], un[6], un[7]));
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int j = 0; j < 8; ++j) p_wh[c][j] = fmaf(gv[i][c], un[j], p_wh[c][j]);
                }
            }
            m1 = half_warp_sum(m1) * (1.f / 128.f);
            m2 = half_warp_sum(m2) * (1.f / 128.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dh = rstd * (dhh[j] - m1 - hh[j] * m2);
                da[i][j] = hv[i][j] > 0.f ? dh * keep_scale : 0.f;
                p_b[j] += da[i][j];
            }
            st4(Da + (ty * 4 + i) * kLd128 + tx * 4, make_float4(da[i][0], da[i][1], da[i][2], da[i][3]));
            st4(Da + (ty * 4 + i) * kLd128 + 64 + tx * 4, make_float4(da[i][4], da[i][5], da[i][6], da[i][7]));
            if (row < rows) {
                float* dst = das + ((size_t)l * rows + row) * kHidden;
                st4(dst + tx * 4, make_float4(da[i][0], da[i][1], da[i][2], da[i][3]));
                st4(dst + 64 + tx * 4, make_float4(da[i][4], da[i][5], da[i][6], da[i][7]));
            }
        }
        if (l == depth - 1) {
            fc_colsum_flush(p_wh[0], red, gr.cls_w, ty, tx, tid);
            fc_colsum_flush(p_wh[1], red, gr.loc_w, ty, tx, tid);
            fc_colsum_flush(p_wh[2], red, gr.loc_w + 128, ty, tx, tid);
        }
        __syncthreads();                    // Da complete
        // do <- da W : dout[r][k] = sum_n Da[r][n] Ws[n][k]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dout[i][j] = 0.f;
#pragma unroll 2
        for (int n4 = 0; n4 < 128; n4 += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = lds4(Da + (ty * 4 + i) * kLd128 + n4);
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                const float4 b0 = lds4(Ws + (n4 + nn) * kLd128 + tx * 4);
                const float4 b1 = lds4(Ws + (n4 + nn) * kLd128 + 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float av = nn == 0 ? a[i].x : nn == 1 ? a[i].y : nn == 2 ? a[i].z : a[i].w;
                    dout[i][0] = fmaf(av, b0.x, dout[i][0]); dout[i][1] = fmaf(av, b0.y, dout[i][1]);
                    dout[i][2] = fmaf(av, b0.z, dout[i][2]); dout[i][3] = fmaf(av, b0.w, dout[i][3]);
                    dout[i][4] = fmaf(av, b1.x, dout[i][4]); dout[i][5] = fmaf(av, b1.y, dout[i][5]);
                    dout[i][6] = fmaf(av, b1.z, dout[i][6]); dout[i][7] = fmaf(av, b1.w, dout[i][7]);
                }
            }
        }
        __syncthreads();                    // every read of Da done before the next layer overwrites it
    }
    float p_f1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p_f1[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + ty * 4 + i;
        if (row < rows) {
            float* dst = du0 + (size_t)row * kHidden;
            st4(dst + tx * 4, make_float4(dout[i][0], dout[i][1], dout[i][2], dout[i][3]));
            st4(dst + 64 + tx * 4, make_float4(dout[i][4], dout[i][5], dout[i][6], dout[i][7]));
#pragma unroll
            for (int j = 0; j < 8; ++j) p_f1[j] += dout[i][j];
        }
    }
    fc_colsum_flush(p_gam, red, gr.fcb_ln_w, ty, tx, tid);
    fc_colsum_flush(p_bet, red, gr.fcb_ln_b, ty, tx, tid);
    fc_colsum_flush(p_b, red, gr.fcb_b, ty, tx, tid);
    fc_colsum_flush(p_f1, red, gr.fc1_b, ty, tx, tid);
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm(1024) backward (anchor_based/dsnet.py:106): y = to_out + x (saved), dyn = d LN output.
//   yh = (y - mean) rstd;  d gamma += dyn yh;  d beta += dyn;  dyh = dyn gamma;
//   dy = rstd (dyh - mean(dyh) - yh mean(dyh yh));  d to_out.bias += dy (column sums)
// One warp per row, 8 rows per warp; the three column-sum vectors live in registers (32 columns per lane) and are
// reduced over the CTA's 8 warps in shared memory, then added atomically.  grid = ceil(R / 64), 256 threads.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln1024_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dyn, const float* __restrict__ gamma,
                  float* __restrict__ dy, int rows, float* __restrict__ d_gamma, float* __restrict__ d_beta,
                  float* __restrict__ d_bias) {
    __shared__ float red[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 a_g[8], a_b[8], a_o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a_g[i] = a_b[i] = a_o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int rr = 0; rr < 8; ++rr) {
        const int row = blockIdx.x * 64 + warp * 8 + rr;
        if (row >= rows) break;
        const float* ys = y + (size_t)row * kFeat;
        const float* ds = dyn + (size_t)row * kFeat;
        float4 x[8], d[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = ldg4(ys + (i * 32 + lane) * 4);
            d[i] = ldg4(ds + (i * 32 + lane) * 4);
            s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
        }
        const float mean = warp_sum(s) * (1.f / (float)kFeat);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
            q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
        }
        const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / (float)kFeat) + 1e-5f);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i].x *= rstd; x[i].y *= rstd; x[i].z *= rstd; x[i].w *= rstd;            // yh
            a_g[i].x = fmaf(d[i].x, x[i].x, a_g[i].x); a_g[i].y = fmaf(d[i].y, x[i].y, a_g[i].y);
            a_g[i].z = fmaf(d[i].z, x[i].z, a_g[i].z); a_g[i].w = fmaf(d[i].w, x[i].w, a_g[i].w);
            a_b[i].x += d[i].x; a_b[i].y += d[i].y; a_b[i].z += d[i].z; a_b[i].w += d[i].w;
            const float4 gm = ldg4(gamma + (i * 32 + lane) * 4);
            d[i].x *= gm.x; d[i].y *= gm.y; d[i].z *= gm.z; d[i].w *= gm.w;              // dyh
            m1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
            m2 += (d[i].x * x[i].x + d[i].y * x[i].y) + (d[i].z * x[i].z + d[i].w * x[i].w);
        }
        m1 = warp_sum(m1) * (1.f / (float)kFeat);
        m2 = warp_sum(m2) * (1.f / (float)kFeat);
        float* dst = dy + (size_t)row * kFeat;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 o = make_float4(rstd * (d[i].x - m1 - x[i].x * m2), rstd * (d[i].y - m1 - x[i].y * m2),
                                         rstd * (d[i].z - m1 - x[i].z * m2), rstd * (d[i].w - m1 - x[i].w * m2));
            a_o[i].x += o.x; a_o[i].y += o.y; a_o[i].z += o.z; a_o[i].w += o.w;
            st4(dst + (i * 32 + lane) * 4, o);
        }
    }
    // reduce the three [1024] vectors over the 8 warps: 128 columns (one i) at a time
    float* outs[3] = {d_gamma, d_beta, d_bias};
#pragma unroll
    for (int which = 0; which < 3; ++which) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = which == 0 ? a_g[i] : which == 1 ? a_b[i] : a_o[i];
            __syncthreads();
            st4(&red[warp][lane * 4], v);
            __syncthreads();
            if (threadIdx.x < 128) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
                atomicAdd(outs[which] + i * 128 + threadIdx.x, s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// torch.optim.Adam(lr, betas, eps, weight_decay) on flat buffers (anchor_based/train.py:53-55):
//   g += wd p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
// `grad_scale` multiplies the incoming gradient first (1 / world after a summing all-reduce).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int n,
            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * grad_scale);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

// =========================================================================================================
// Nystrom attention block, backward (transformer/nystroformer.py:95-142 in reverse).  Per (video, head), with
//   A1 = softmax(q kl^T) (n x 64), A2 = softmax(ql kl^T), A3 = softmax_keys(ql k^T) (64 x n, zero pad keys included in the
//   normalisation), B = A3 v, Z = pinv(A2), W = Z B, out = A1 W + conv(v):
//   rows kernel   dA1 = dout W^T, dS1 = A1 (dA1 - rowsum(dA1 A1)), dq = dS1 kl, dW += A1^T dout, dkl += dS1^T q,
//                 conv: dv[r] = sum_t taps[t] dout[r - t + 16], dtaps[t] += sum_r dout[r] . v[r + t - 16]
//   pinv kernel   dZ = dW B^T, dB = Z^T dW, six Newton-Schulz steps in reverse -> dA2 (without the start-scale term), dc
//   attn2 kernel  start-scale term, dS2 = A2 (dA2 - rowsum(dA2 A2)), dql += dS2 kl, dkl += dS2^T ql
//   keys kernel   dA3 = dB v^T, delta_j = <dB_j, B_j>, dS3 = A3 (dA3 - delta), dk = dS3^T ql, dv += A3^T dB, dql += dS3 k
//   finish        landmark means back to their rows (1 / seg), q's 1/8
// Everything reads the fp32 copy qkv [R][1536] (q pre-scaled by 1/8) and writes dqkv [R][1536] in the same layout.
// =========================================================================================================

// Row maximum and normaliser of S3 = ql k^T over ALL keys of the padded sequence (zero pad keys: logit 0), per
// (video, head): m3, l3 [V][8][64].  grid (8, V), 256 threads.
__global__ void __launch_bounds__(256)
a3_stats_kernel(const float* __restrict__ qkv, const int* __restrict__ cu_rows, const float* __restrict__ q_land,
                float* __restrict__ m3, float* __restrict__ l3) {
    __shared__ __align__(16) float Ql[64 * kLd64];
    __shared__ __align__(16) float Ks[64 * kLd64];
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const VidInfo vi = vid_info(cu_rows, v);
    load64_rowmajor(Ql, q_land + ((size_t)v * kHeads + h) * 4096, 64, tid);
    float run_max[4], run_sum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { run_max[i] = vi.pad > 0 ? 0.f : -INFINITY; run_sum[i] = (float)vi.pad; }
    const float* kbase = qkv + (size_t)vi.row0 * kQkvCols + kInner + h * kDimHead;
    for (int r0 = 0; r0 < vi.T; r0 += 64) {
        __syncthreads();
        load64_rows(Ks, kbase + (size_t)r0 * kQkvCols, kQkvCols, vi.T - r0, tid);
        __syncthreads();
        float s[4][4];
        zero44(s);
        mm64<MM_NT>(s, Ql, Ks, ty, tx);                         // [landmark ty*4+i][key tx+16j]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (r0 + tx + 16 * j >= vi.T) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
            mx = half_warp_max(mx);
            const float nm = fmaxf(run_max[i], mx);
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) ps += expf(s[i][j] - nm);
            ps = half_warp_sum(ps);
            run_sum[i] = run_sum[i] * expf(run_max[i] - nm) + ps;
            run_max[i] = nm;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m3[((size_t)v * kHeads + h) * 64 + ty * 4 + i] = run_max[i];
            l3[((size_t)v * kHeads + h) * 64 + ty * 4 + i] = run_sum[i];
        }
    }
}

// rows kernel: one CTA per (64-row tile of one video, head); grid (n_tiles64, 8).
// dmerged [R][512] = d(out), head-merged.  Writes the q part (without the landmark term and the 1/8) and the conv part
// of v into dqkv; adds dW [V][8][64][64], dkl [V][8][64][64] and d res_conv.weight [8][33] atomically.
constexpr int kAbrWin = 64 + 2 * (kTaps / 2);                                // 96 window rows
constexpr int kAttnBwdRowsSmem = (5 * 64 * kLd64 + 2 * kAbrWin * kLd64 + 64 + 8 * 40) * (int)sizeof(float);

__global__ void __launch_bounds__(256)
attn_bwd_rows_kernel(const float* __restrict__ qkv, const float* __restrict__ dmerged, const int* __restrict__ cu_rows,
                     const int2* __restrict__ tiles, const float* __restrict__ k_land, const float* __restrict__ w_mat,
                     const float* __restrict__ conv_w, float* __restrict__ dqkv, float* __restrict__ dW,
                     float* __restrict__ dkl, float* __restrict__ d_conv_w) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                          // q tile [row][d]
    float* KL = Qs + 64 * kLd64;               // k_land [j][d]
    float* Ws = KL + 64 * kLd64;               // W [j][d]
    float* Ps = Ws + 64 * kLd64;               // A1 [row][j]
    float* Ds = Ps + 64 * kLd64;               // dS1 [row][j], later scratch
    float* Dw = Ds + 64 * kLd64;               // dout window rows r0-16 .. r0+79, [96][d]; the tile itself is Dw + 16 rows
    float* Vw = Dw + kAbrWin * kLd64;          // v window, same rows
    float* taps = Vw + kAbrWin * kLd64;        // [33]
    float* tred = taps + 64;                   // [8 warps][33 (+pad)]
    const int h = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int2 tile = tiles[blockIdx.x];
    const int v = tile.x, r0 = tile.y;
    const VidInfo vi = vid_info(cu_rows, v);
    const size_t hoff = ((size_t)v * kHeads + h) * 4096;
    const float* qbase = qkv + (size_t)vi.row0 * kQkvCols + h * kDimHead;
    const float* vbase = qbase + 2 * kInner;
    const float* dobase = dmerged + (size_t)vi.row0 * kInner + h * kDimHead;
    float* Do = Dw + (kTaps / 2) * kLd64;

    load64_rows(Qs, qbase + (size_t)r0 * kQkvCols, kQkvCols, vi.T - r0, tid);
    load64_rowmajor(KL, k_land + hoff, 64, tid);
    load64_rowmajor(Ws, w_mat + hoff, 64, tid);
    for (int idx = tid; idx < kAbrWin * 16; idx += 256) {
        const int r = idx >> 4, c4 = (idx & 15) * 4;
        const int rr = r0 - kTaps / 2 + r;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (rr >= 0 && rr < vi.T) {
            a = ldg4(dobase + (size_t)rr * kInner + c4);
            b = ldg4(vbase + (size_t)rr * kQkvCols + c4);
        }
        st4(Dw + r * kLd64 + c4, a);
        st4(Vw + r * kLd64 + c4, b);
    }
    if (tid < kTaps) taps[tid] = __ldg(conv_w + h * kTaps + tid);
    __syncthreads();

    // S1 = q kl^T -> A1; dA1 = dout W^T; dS1 = A1 (dA1 - rowsum(dA1 A1))      (NT ownership: rows ty*4+i, cols tx+16j)
    {
        float s[4][4], dp[4][4];
        zero44(s);
        mm64<MM_NT>(s, Qs, KL, ty, tx);
        zero44(dp);
        mm64<MM_NT>(dp, Do, Ws, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
            mx = half_warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - mx); sum += s[i][j]; }
            sum = half_warp_sum(sum);
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = s[i][j] / sum; dot = fmaf(dp[i][j], s[i][j], dot); }
            dot = half_warp_sum(dot);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                Ps[(ty * 4 + i) * kLd64 + tx + 16 * j] = s[i][j];
                Ds[(ty * 4 + i) * kLd64 + tx + 16 * j] = s[i][j] * (dp[i][j] - dot);
            }
        }
    }
    __syncthreads();
    // dq = dS1 kl  (q part of dqkv, landmark term and 1/8 added by dqkv_finish_kernel)
    {
        float acc[4][4];
        zero44(acc);
        mm64<MM_NN>(acc, Ds, KL, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty * 4 + i;
            if (r < vi.T)
                st4(dqkv + (size_t)(vi.row0 + r) * kQkvCols + h * kDimHead + tx * 4,
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        }
    }
    // conv: dv[r] = sum_t taps[t] dout[r - t + 16]  (window row of output row i and tap t: i + 32 - t)
    {
        float cv[4][4];
        zero44(cv);
#pragma unroll
        for (int rr = 0; rr < 4 + kTaps - 1; ++rr) {
            const float4 x = lds4(Dw + (ty * 4 + rr) * kLd64 + tx * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = i + 32 - rr;
                if (t >= 0 && t < kTaps) {
                    const float wv = taps[t];
                    cv[i][0] = fmaf(wv, x.x, cv[i][0]); cv[i][1] = fmaf(wv, x.y, cv[i][1]);
                    cv[i][2] = fmaf(wv, x.z, cv[i][2]); cv[i][3] = fmaf(wv, x.w, cv[i][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty * 4 + i;
            if (r < vi.T)
                st4(dqkv + (size_t)(vi.row0 + r) * kQkvCols + 2 * kInner + h * kDimHead + tx * 4,
                    make_float4(cv[i][0], cv[i][1], cv[i][2], cv[i][3]));
        }
    }
    // d taps[t] = sum_{r in tile} dout[r] . v[r + t - 16]   (window rows: dout i+16, v i+t)
    {
        float4 dd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dd[i] = lds4(Do + (ty * 4 + i) * kLd64 + tx * 4);
        const int warp = tid >> 5, lane = tid & 31;
        for (int t = 0; t < kTaps; ++t) {
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 x = lds4(Vw + (ty * 4 + i + t) * kLd64 + tx * 4);
                p = fmaf(dd[i].x, x.x, fmaf(dd[i].y, x.y, fmaf(dd[i].z, x.z, fmaf(dd[i].w, x.w, p))));
            }
            p = warp_sum(p);
            if (lane == 0) tred[warp * 40 + t] = p;
        }
    }
    __syncthreads();
    if (tid < kTaps) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += tred[w * 40 + tid];
        atomicAdd(d_conv_w + h * kTaps + tid, s);
    }
    // dW += A1^T dout ; dkl += dS1^T q      (TN products into scratch tiles, then atomics)
    {
        float acc[4][4];
        zero44(acc);
        mm64<MM_TN>(acc, Ps, Do, ty, tx);
        float acc2[4][4];
        zero44(acc2);
        mm64<MM_TN>(acc2, Ds, Qs, ty, tx);
        __syncthreads();                                  // all reads of Ps / Ds done: reuse them as staging tiles
        mm64_store<MM_TN>(Ps, acc, 1.f, 0.f, 0.f, ty, tx);
        mm64_store<MM_TN>(Ds, acc2, 1.f, 0.f, 0.f, ty, tx);
        __syncthreads();
        atomic_add64(dW + hoff, Ps, tid);
        atomic_add64(dkl + hoff, Ds, tid);
    }
}

// pinv kernels: one CTA per (video, head); grid (8, V), 512 threads (two rows per thread: a dependent chain of 64^3 products
// on ONE SM is bound by how well the FFMA latency is hidden).  attn2 = A, a3v = B, stats as attn2_kernel wrote them.
//   pinv_hist_kernel  recomputes the forward chain in fp32 and keeps Z_k (input of iteration k), P_k = A Z_k, T2_k, T3_k
//                     and the final Z: hist [V][8][iters * 4 + 1][64][64].  It depends on the forward only, so the
//                     backward runs it on its side stream under the kernels in front of the attention block.
//   pinv_bwd_kernel   dZ = dW B^T, dB = Z^T dW, then the iterations in reverse, reading the kept tiles back (L2) instead of
//                     recomputing three products per iteration.
// Outputs: dB [V][8][64][64], dA2 (without the start-scale term) [V][8][64][64], dc_part [V][8].
constexpr int kPinvBwdThreads = 512;
constexpr int kPinvHistSmem = 7 * 64 * kLd64 * (int)sizeof(float);
constexpr int kPinvBwdSmem = (11 * 64 * kLd64 + 16) * (int)sizeof(float);

__device__ __forceinline__ float pinv_start_scale(const float* __restrict__ stats, int v) {
    float mrow = 0.f, mcol = 0.f;
#pragma unroll
    for (int hh = 0; hh < kHeads; ++hh) {
        mrow = fmaxf(mrow, __ldg(stats + ((size_t)v * kHeads + hh) * 2 + 0));
        mcol = fmaxf(mcol, __ldg(stats + ((size_t)v * kHeads + hh) * 2 + 1));
    }
    return mrow * mcol;
}

__global__ void __launch_bounds__(kPinvBwdThreads)
pinv_hist_kernel(const float* __restrict__ attn2, const float* __restrict__ stats, float* __restrict__ hist, int iters) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Zs = As + 64 * kLd64;
    float* Pz = Zs + 64 * kLd64;
    float* T1 = Pz + 64 * kLd64;
    float* T2 = T1 + 64 * kLd64;
    float* T3 = T2 + 64 * kLd64;
    float* Zn = T3 + 64 * kLd64;
    constexpr int RT = 2, NT = kPinvBwdThreads;
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const size_t off = ((size_t)v * kHeads + h) * 4096;
    float* hs = hist + off * (iters * 4 + 1);
    const float denom = pinv_start_scale(stats, v);
    for (int idx = tid; idx < 1024; idx += NT) {
        const int r = idx >> 4, c4 = (idx & 15) * 4;
        const float4 a = ldg4(attn2 + off + r * 64 + c4);
        st4(As + r * kLd64 + c4, a);
        Zs[(c4 + 0) * kLd64 + r] = a.x / denom; Zs[(c4 + 1) * kLd64 + r] = a.y / denom;
        Zs[(c4 + 2) * kLd64 + r] = a.z / denom; Zs[(c4 + 3) * kLd64 + r] = a.w / denom;
    }
    __syncthreads();
    for (int k = 0; k < iters; ++k) {
        float* hk = hs + (size_t)k * 4 * 4096;
        tile_store(hk, Zs, tid, NT);
        mm64_to<MM_NN, RT>(Pz, As, Zs, 1.f, 0.f, 0.f, ty, tx);
        __syncthreads();
        tile_store(hk + 4096, Pz, tid, NT);
        for (int idx = tid; idx < 4096; idx += NT) {
            const int r = idx >> 6, c = idx & 63;
            T1[r * kLd64 + c] = (r == c ? 7.f : 0.f) - Pz[r * kLd64 + c];
        }
        __syncthreads();
        mm64_to<MM_NN, RT>(T2, Pz, T1, -1.f, 0.f, 15.f, ty, tx);
        __syncthreads();
        tile_store(hk + 2 * 4096, T2, tid, NT);
        mm64_to<MM_NN, RT>(T3, Pz, T2, -1.f, 0.f, 13.f, ty, tx);
        __syncthreads();
        tile_store(hk + 3 * 4096, T3, tid, NT);
        mm64_to<MM_NN, RT>(Zn, Zs, T3, 0.25f, 0.f, 0.f, ty, tx);
        __syncthreads();
        float* t = Zs; Zs = Zn; Zn = t;
    }
    tile_store(hs + (size_t)iters * 4 * 4096, Zs, tid, NT);
}

__global__ void __launch_bounds__(kPinvBwdThreads)
pinv_bwd_kernel(const float* __restrict__ attn2, const float* __restrict__ stats, const float* __restrict__ a3v,
                const float* __restrict__ dW, const float* __restrict__ hist, float* __restrict__ dB,
                float* __restrict__ dA2, float* __restrict__ dc_part, int iters) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Zs = As + 64 * kLd64;
    float* Pz = Zs + 64 * kLd64;
    float* T1 = Pz + 64 * kLd64;
    float* T2 = T1 + 64 * kLd64;
    float* T3 = T2 + 64 * kLd64;
    float* dZ = T3 + 64 * kLd64;
    float* Zn = dZ + 64 * kLd64;
    float* Us = Zn + 64 * kLd64;
    float* Vs = Us + 64 * kLd64;
    float* DP = Vs + 64 * kLd64;
    float* red = DP + 64 * kLd64;
    constexpr int RT = 2, NT = kPinvBwdThreads;
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const size_t off = ((size_t)v * kHeads + h) * 4096;
    const float* hs = hist + off * (iters * 4 + 1);
    const float denom = pinv_start_scale(stats, v);
    // ---- W = Z B:  dZ = dW B^T,  dB = Z^T dW ----
    tile_load(As, attn2 + off, tid, NT);
    tile_load(Zs, hs + (size_t)iters * 4 * 4096, tid, NT);
    tile_load(Us, a3v + off, tid, NT);
    tile_load(Vs, dW + off, tid, NT);
    __syncthreads();
    mm64_to<MM_NT, RT>(dZ, Vs, Us, 1.f, 0.f, 0.f, ty, tx);
    mm64_to<MM_TN, RT>(DP, Zs, Vs, 1.f, 0.f, 0.f, ty, tx);
    __syncthreads();
    tile_store(dB + off, DP, tid, NT);
    float dA[RT][4];                                  // NT ownership: rows ty*2+i, cols tx+16j
    zero_acc<RT>(dA);
    // ---- the iterations in reverse ----
    for (int k = iters - 1; k >= 0; --k) {
        const float* hk = hs + (size_t)k * 4 * 4096;
        __syncthreads();                              // everything of the previous round (and the dB store) has read its tiles
        tile_load(Zs, hk, tid, NT);
        tile_load(Pz, hk + 4096, tid, NT);
        tile_load(T2, hk + 2 * 4096, tid, NT);
        tile_load(T3, hk + 3 * 4096, tid, NT);
        __syncthreads();
        for (int idx = tid; idx < 4096; idx += NT) {
            const int r = idx >> 6, c = idx & 63;
            T1[r * kLd64 + c] = (r == c ? 7.f : 0.f) - Pz[r * kLd64 + c];
        }
        mm64_to<MM_TN, RT>(Us, Zs, dZ, 0.25f, 0.f, 0.f, ty, tx);          // dT3 = 0.25 Z^T dZ
        __syncthreads();
        mm64_to<MM_NT, RT>(Zn, dZ, T3, 0.25f, 0.f, 0.f, ty, tx);          // dZ' = 0.25 dZ T3^T
        mm64_to<MM_NT, RT>(DP, Us, T2, -1.f, 0.f, 0.f, ty, tx);           // dP  = -dT3 T2^T
        mm64_to<MM_TN, RT>(Vs, Pz, Us, -1.f, 0.f, 0.f, ty, tx);           // dT2 = -P^T dT3
        __syncthreads();
        mm64_to<MM_NT, RT>(DP, Vs, T1, -1.f, 1.f, 0.f, ty, tx);           // dP -= dT2 T1^T      (same ownership as above)
        __syncthreads();                                                  // Us (dT3) no longer read
        mm64_to<MM_TN, RT>(Us, Pz, Vs, -1.f, 0.f, 0.f, ty, tx);           // dT1 = -P^T dT2
        __syncthreads();
        for (int idx = tid; idx < 4096; idx += NT) {
            const int r = idx >> 6, c = idx & 63;
            DP[r * kLd64 + c] -= Us[r * kLd64 + c];                       // dP -= dT1
        }
        __syncthreads();
        mm64<MM_NT, RT>(dA, DP, Zs, ty, tx);                              // dA += dP Z^T
        mm64_to<MM_TN, RT>(Zn, As, DP, 1.f, 1.f, 0.f, ty, tx);            // dZ' += A^T dP
        __syncthreads();
        float* t = dZ; dZ = Zn; Zn = t;
    }
    // Z_0 = A^T / c:  dA += dZ^T / c,  dc = - sum dZ o A^T / c^2
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ty * RT + i, c = tx + 16 * j;
            const float dzt = dZ[c * kLd64 + r];
            dA[i][j] += dzt / denom;
            part = fmaf(dzt, As[r * kLd64 + c], part);
            dA2[off + r * 64 + c] = dA[i][j];
        }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) sum += red[w];
        dc_part[v * kHeads + h] = -sum / (denom * denom);
    }
}

// attn2 kernel: start-scale term + softmax back-substitution; grid (8, V).
//   c = max_h max_i rowsum x max_h max_j colsum (nystroformer.py:16-19).  The row sums of a softmax are all 1: whichever
//   row holds the maximum, its gradient is a constant along that row and is annihilated by the softmax back-substitution,
//   so only the column part is applied: + dc * rowmax on column j* of head h* (first maximum, as attn2_kernel's sums).
__global__ void __launch_bounds__(256)
attn2_bwd_kernel(const float* __restrict__ attn2, const float* __restrict__ stats, const float* __restrict__ dA2,
                 const float* __restrict__ dc_part, const float* __restrict__ q_land, const float* __restrict__ k_land,
                 float* __restrict__ dql, float* __restrict__ dkl) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Ds = As + 64 * kLd64;
    float* QL = Ds + 64 * kLd64;
    float* KL = QL + 64 * kLd64;
    __shared__ float colsum[64];
    __shared__ int jstar;
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const size_t off = ((size_t)v * kHeads + h) * 4096;
    load64_rowmajor(As, attn2 + off, 64, tid);
    load64_rowmajor(Ds, dA2 + off, 64, tid);
    load64_rowmajor(QL, q_land + off, 64, tid);
    load64_rowmajor(KL, k_land + off, 64, tid);
    float mrow = 0.f, mcol = 0.f, dc = 0.f;
    int hstar = 0;
#pragma unroll
    for (int hh = 0; hh < kHeads; ++hh) {
        mrow = fmaxf(mrow, __ldg(stats + ((size_t)v * kHeads + hh) * 2 + 0));
        const float cm = __ldg(stats + ((size_t)v * kHeads + hh) * 2 + 1);
        if (cm > mcol) { mcol = cm; hstar = hh; }
        dc += __ldg(dc_part + v * kHeads + hh);
    }
    __syncthreads();
    if (h == hstar) {
        // the same sequential column sums attn2_kernel took its maximum from: bit-identical, so == finds the column
        if (tid < 64) {
            float s = 0.f;
            for (int i = 0; i < 64; ++i) s += As[i * kLd64 + tid];
            colsum[tid] = s;
        }
        __syncthreads();
        if (tid == 0) {
            int js = 0;
            float best = colsum[0];
            for (int j = 1; j < 64; ++j)
                if (colsum[j] > best) { best = colsum[j]; js = j; }
            jstar = js;
        }
        __syncthreads();
        if (tid < 64) Ds[tid * kLd64 + jstar] += dc * mrow;
        __syncthreads();
    }
    // dS2 = A2 (dA2 - rowsum(dA2 A2)), in place in Ds
    {
        float a[4][4], d[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 av = lds4(As + (ty * 4 + i) * kLd64 + tx * 4), dv = lds4(Ds + (ty * 4 + i) * kLd64 + tx * 4);
            a[i][0] = av.x; a[i][1] = av.y; a[i][2] = av.z; a[i][3] = av.w;
            d[i][0] = dv.x; d[i][1] = dv.y; d[i][2] = dv.z; d[i][3] = dv.w;
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) dot = fmaf(a[i][j], d[i][j], dot);
            dot = half_warp_sum(dot);
            st4(Ds + (ty * 4 + i) * kLd64 + tx * 4, make_float4(a[i][0] * (d[i][0] - dot), a[i][1] * (d[i][1] - dot),
                                                                a[i][2] * (d[i][2] - dot), a[i][3] * (d[i][3] - dot)));
        }
    }
    __syncthreads();
    float acc[4][4], acc2[4][4];
    zero44(acc);
    mm64<MM_NN>(acc, Ds, KL, ty, tx);                     // dql += dS2 kl
    zero44(acc2);
    mm64<MM_TN>(acc2, Ds, QL, ty, tx);                    // dkl += dS2^T ql
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(dql + off + (ty * 4 + i) * 64 + tx * 4 + j, acc[i][j]);
            atomicAdd(dkl + off + (ty * 4 + i) * 64 + tx * 4 + j, acc2[i][j]);
        }
}

// keys kernel: one CTA per (64-key tile of one video, head); grid (n_tiles64, 8).  Writes the k part of dqkv, ADDS the
// aggregation part to the v part (the rows kernel wrote the convolution part before), adds dql atomically.
constexpr int kAttnBwdKeysSmem = (6 * 64 * kLd64 + 64) * (int)sizeof(float);

__global__ void __launch_bounds__(256)
attn_bwd_keys_kernel(const float* __restrict__ qkv, const int* __restrict__ cu_rows, const int2* __restrict__ tiles,
                     const float* __restrict__ q_land, const float* __restrict__ a3v, const float* __restrict__ dB,
                     const float* __restrict__ m3, const float* __restrict__ l3, float* __restrict__ dqkv,
                     float* __restrict__ dql) {
    extern __shared__ __align__(16) float smem[];
    float* QL = smem;                          // ql [j][d]
    float* Ks = QL + 64 * kLd64;               // k tile [key][d]
    float* Vt = Ks + 64 * kLd64;               // v tile [key][d]
    float* Bs = Vt + 64 * kLd64;               // dB [j][d]
    float* Ps = Bs + 64 * kLd64;               // A3 [j][key]
    float* Ds = Ps + 64 * kLd64;               // dS3 [j][key]
    float* delta = Ds + 64 * kLd64;            // [64]
    const int h = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int2 tile = tiles[blockIdx.x];
    const int v = tile.x, r0 = tile.y;
    const VidInfo vi = vid_info(cu_rows, v);
    const size_t hoff = ((size_t)v * kHeads + h) * 4096;
    const float* kbase = qkv + (size_t)vi.row0 * kQkvCols + kInner + h * kDimHead;
    load64_rowmajor(QL, q_land + hoff, 64, tid);
    load64_rows(Ks, kbase + (size_t)r0 * kQkvCols, kQkvCols, vi.T - r0, tid);
    load64_rows(Vt, kbase + kInner + (size_t)r0 * kQkvCols, kQkvCols, vi.T - r0, tid);
    load64_rowmajor(Bs, dB + hoff, 64, tid);
    // delta_j = <dB_j, B_j>
    {
        float part[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 b = ldg4(a3v + hoff + (ty * 4 + i) * 64 + tx * 4), d = ldg4(dB + hoff + (ty * 4 + i) * 64 + tx * 4);
            part[i] = half_warp_sum(fmaf(b.x, d.x, fmaf(b.y, d.y, fmaf(b.z, d.z, b.w * d.w))));
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) delta[ty * 4 + i] = part[i];
        }
    }
    __syncthreads();
    {
        float s[4][4], da[4][4];
        zero44(s);
        mm64<MM_NT>(s, QL, Ks, ty, tx);                   // S3 [j][key]
        zero44(da);
        mm64<MM_NT>(da, Bs, Vt, ty, tx);                  // dA3 = dB v^T
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = ty * 4 + i;
            const float mx = __ldg(m3 + ((size_t)v * kHeads + h) * 64 + j), ls = __ldg(l3 + ((size_t)v * kHeads + h) * 64 + j);
            const float dl = delta[j];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int key = tx + 16 * q;
                const float p = (r0 + key < vi.T) ? expf(s[i][q] - mx) / ls : 0.f;
                Ps[j * kLd64 + key] = p;
                Ds[j * kLd64 + key] = p * (da[i][q] - dl);
            }
        }
    }
    __syncthreads();
    {
        float dk[4][4], dv[4][4], dq[4][4];
        zero44(dk);
        mm64<MM_TN>(dk, Ds, QL, ty, tx);                  // dk [key][d] = dS3^T ql
        zero44(dv);
        mm64<MM_TN>(dv, Ps, Bs, ty, tx);                  // dv [key][d] = A3^T dB
        zero44(dq);
        mm64<MM_NN>(dq, Ds, Ks, ty, tx);                  // dql [j][d] += dS3 k
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty * 4 + i;
            if (r < vi.T) {
                float* base = dqkv + (size_t)(vi.row0 + r) * kQkvCols + h * kDimHead + tx * 4;
                st4(base + kInner, make_float4(dk[i][0], dk[i][1], dk[i][2], dk[i][3]));
                const float4 old = lds4(base + 2 * kInner);
                st4(base + 2 * kInner, make_float4(old.x + dv[i][0], old.y + dv[i][1], old.z + dv[i][2], old.w + dv[i][3]));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dql + hoff + (ty * 4 + i) * 64 + tx * 4 + j, dq[i][j]);
        }
    }
}

// finish: dq = (dq_part + dql[landmark of the row] / seg) / 8,  dk = dk_part + dkl[landmark] / seg; in place.
// One CTA per 64-row tile; thread <-> 4 of the 1024 q|k columns.
constexpr int kFinishSplit = 8;                       // grid.y: a 64-row tile is finished by 8 CTAs of 8 rows each
__global__ void __launch_bounds__(256)
dqkv_finish_kernel(const int* __restrict__ cu_rows, const int2* __restrict__ tiles, const float* __restrict__ dql,
                   const float* __restrict__ dkl, float* __restrict__ dqkv) {
    const int2 tile = tiles[blockIdx.x];
    const int v = tile.x, r0 = tile.y + (int)blockIdx.y * (64 / kFinishSplit);
    const VidInfo vi = vid_info(cu_rows, v);
    const int c4 = threadIdx.x * 4;
    const bool is_k = c4 >= kInner;
    const int cc = c4 & (kInner - 1);
    const int hd = cc >> 6, d = cc & 63;
    const float* land = (is_k ? dkl : dql) + ((size_t)v * kHeads + hd) * 4096 + d;
    const float inv_seg = 1.f / (float)vi.seg, mul = is_k ? 1.f : 0.125f;
    const int rend = min(r0 + 64 / kFinishSplit, vi.T);
    for (int r = r0; r < rend; ++r) {
        const int j = (r + vi.pad) / vi.seg;
        const float4 l = ldg4(land + j * 64);
        float* p = dqkv + (size_t)(vi.row0 + r) * kQkvCols + c4;
        const float4 x = lds4(p);
        st4(p, make_float4((x.x + l.x * inv_seg) * mul, (x.y + l.y * inv_seg) * mul, (x.z + l.z * inv_seg) * mul,
                           (x.w + l.w * inv_seg) * mul));
    }
}
