// Evaluation metrics of the keyshot summaries on the device: evaluate.py:31-37 ->
// helpers/vsumm_helper.py:142-172 (get_summ_f1score), :8-23 (f1_score), :48-50 (downsample_summ), :119-139
// (get_summ_diversity).  One CTA per video:
//   F-score: the predicted summary is cut / zero-padded to the users' frame count N; per user integer counts
//            overlap = |pred & user|, |pred|, |user| (block reductions), then float64 precision / recall / F1 exactly as
//            NumPy evaluates them; 'avg' = NumPy's pairwise float64 mean over the users, 'max' = their maximum.
//   diversity: rows t with summary[15 t] set; mean over ordered pairs i != j of f_i . f_j
//            = (|sum_i f_i|^2 - sum_i |f_i|^2) / (P (P - 1)), accumulated in float64 (the reference sums float32
//            products pair-wise: agreement to ~1e-6 relative, not bit-wise).
#pragma once
#include "common.cuh"
#include "summary.cuh"

struct EvalTruth {
    const int* cu_users;              // [V+1] first user row of every video
    const long long* user_off;        // [total_users] byte offset of the user's 0/1 row in user_summ (4-byte aligned)
    const int* user_frames;           // [V] frames per user row of the video
    const unsigned char* user_summ;   // 0/1 bytes
    const int* metric;                // [V] 0 = avg (TVSum), 1 = max (SumMe)
};

__device__ __forceinline__ int block_sum_int(int v, int* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    return t;
}
__device__ __forceinline__ double block_sum_f64(double v, double* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    return t;
}

// float64 sum of a[0..n) in NumPy's pairwise order (summary.cuh)
__device__ __forceinline__ double numpy_pairwise_sum_f64(const double* __restrict__ a, int n) {
    return numpy_pairwise_sum_t<double>(a, n);
}

__global__ void __launch_bounds__(256)
eval_metrics_kernel(const int* __restrict__ cu_rows, const long long* __restrict__ cu_frames,
                    const unsigned char* __restrict__ summary, EvalTruth tr, const float* __restrict__ x,
                    double* __restrict__ fscore, double* __restrict__ diversity, double* __restrict__ user_f1,
                    int* __restrict__ counts) {
    __shared__ int s_red[8];
    __shared__ double s_redd[8];
    const int v = blockIdx.x, tid = threadIdx.x;
    const VidInfo vi = vid_info(cu_rows, v);
    const long long f0 = cu_frames[v];
    const int n_frames = (int)(cu_frames[v + 1] - f0);
    const unsigned char* summ = summary + f0;
    const int u0 = tr.cu_users[v], n_users = tr.cu_users[v + 1] - u0;
    const int N = tr.user_frames[v];
    const int n_cmp = min(N, n_frames);                    // frames where the cut / padded prediction can be set

    // ---- F-score ----
    int ps = 0;
    for (int f = tid; f < n_cmp; f += 256) ps += summ[f] != 0;
    const int pred_sum = block_sum_int(ps, s_red);
    for (int u = 0; u < n_users; ++u) {
        const unsigned char* row = tr.user_summ + tr.user_off[u0 + u];
        const uchar4* row4 = reinterpret_cast<const uchar4*>(row);
        int ov = 0, ts = 0;
        for (int f4 = tid; f4 * 4 < N; f4 += 256) {
            const int f = f4 * 4;
            if (f + 3 < N) {
                const uchar4 w = row4[f4];
                const int b0 = w.x != 0, b1 = w.y != 0, b2 = w.z != 0, b3 = w.w != 0;
                ts += b0 + b1 + b2 + b3;
                if (b0 && f < n_cmp) ov += summ[f] != 0;
                if (b1 && f + 1 < n_cmp) ov += summ[f + 1] != 0;
                if (b2 && f + 2 < n_cmp) ov += summ[f + 2] != 0;
                if (b3 && f + 3 < n_cmp) ov += summ[f + 3] != 0;
            } else {
                for (int q = f; q < N; ++q) {
                    const int b = row[q] != 0;
                    ts += b;
                    if (b && q < n_cmp) ov += summ[q] != 0;
                }
            }
        }
        const int overlap = block_sum_int(ov, s_red);
        const int test_sum = block_sum_int(ts, s_red);
        if (tid == 0) {
            double f1 = 0.0;
            if (overlap != 0) {
                // f1_score(pred=user, test=prediction): precision = overlap / |user|, recall = overlap / |prediction|
                const double precision = __ddiv_rn((double)overlap, (double)test_sum);
                const double recall = __ddiv_rn((double)overlap, (double)pred_sum);
                f1 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, precision), recall), __dadd_rn(precision, recall));
            }
            user_f1[u0 + u] = f1;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double r = 0.0;
        if (n_users > 0) {
            if (tr.metric[v] == 0) {
                r = __ddiv_rn(numpy_pairwise_sum_f64(user_f1 + u0, n_users), (double)n_users);
            } else {
                r = user_f1[u0];
                for (int u = 1; u < n_users; ++u) r = fmax(r, user_f1[u0 + u]);
            }
        }
        fscore[v] = r;
        counts[v * 2 + 0] = pred_sum;
    }

    // ---- diversity over the down-sampled summary (summ[::15] <-> feature rows) ----
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, sq = 0.0;
    int P = 0;
    for (int t = 0; t < vi.T; ++t) {
        const long long f = (long long)t * 15;
        if (f >= n_frames || summ[f] == 0) continue;       // uniform across the block
        ++P;
        const float4 a = ldg4(x + (size_t)(vi.row0 + t) * kFeat + tid * 4);
        s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
        sq += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
    }
    const double ss = block_sum_f64(s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3, s_redd);
    const double sqs = block_sum_f64(sq, s_redd);
    if (tid == 0) {
        diversity[v] = P < 2 ? 0.0 : (ss - sqs) / ((double)P * (double)(P - 1));
        counts[v * 2 + 1] = P;
    }
}
