// Kept proposals -> keyshot summary on the device: helpers/vsumm_helper.py:101-116 (bbox2summary) and
// :53-98 (get_keyshot_summ) with the knapsack of :26-45 solved exactly by dynamic programming over the capacity.
// One CTA per video:
//   1. pos_score[t]   = max over kept boxes [lo, hi) covering position t (0 if none)
//   2. frame_score[f] = pos_score[i] for picks[i] <= f < picks[i+1] (last pick extends to n_frames)
//   3. seg_score[j]   = int(1000 * float32 mean of frame_score[first..last]) with NumPy's arithmetic: pair-wise
//                       float32 summation (8 accumulators, blocks of 128), sum / count in float64 rounded to float32,
//                       float32 product with 1000, truncation
//   4. 0/1 knapsack, value seg_score, weight nfps, capacity int(0.15 n_frames): rows of the DP table over the capacity
//      in parallel, one decision bit per (item, capacity), back-tracked from the last item.  ortools (the reference's
//      solver) is not available; among equal-value optima this picks the set the oracle's DP picks (an item is taken
//      only if leaving it out loses value, scanning items last to first) -- the reference's own choice is unpinned.
//   5. summary[first..last] = 1 for the chosen shots
#pragma once
#include "common.cuh"

// float32 sum of a[0..n) exactly as NumPy's pairwise_sum (numpy/_core/src/umath/loops_utils.h.src) computes it:
// blocks of at most 128 values are summed with 8 interleaved accumulators, longer ranges are split in two (the left half
// a multiple of 8 long) and the halves' sums added.  The recursion is unrolled onto an explicit stack (depth
// <= log2(n / 64)), so the kernel needs no device call stack.
template <typename T>
__device__ __forceinline__ T numpy_pairwise_leaf(const T* __restrict__ a, int n) {
    if (n < 8) {
        T res = 0;
        for (int i = 0; i < n; ++i) res = res + a[i];
        return res;
    }
    T r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = r[k] + a[i + k];
    }
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + a[i];
    return res;
}
template <typename T>
__device__ T numpy_pairwise_sum_t(const T* __restrict__ a, int n) {
    if (n <= 128) return numpy_pairwise_leaf(a, n);
    int s_off[32], s_n[32];
    T s_left[32];
    unsigned char s_stage[32];
    int sp = 0;
    s_off[0] = 0; s_n[0] = n; s_stage[0] = 0;
    T ret = 0;
    while (sp >= 0) {
        const int off = s_off[sp], len = s_n[sp];
        if (len <= 128) { ret = numpy_pairwise_leaf(a + off, len); --sp; continue; }
        int n2 = len / 2;
        n2 -= n2 % 8;
        if (s_stage[sp] == 0) {
            s_stage[sp] = 1;
            ++sp; s_off[sp] = off; s_n[sp] = n2; s_stage[sp] = 0;
        } else if (s_stage[sp] == 1) {
            s_left[sp] = ret;
            s_stage[sp] = 2;
            ++sp; s_off[sp] = off + n2; s_n[sp] = len - n2; s_stage[sp] = 0;
        } else {
            ret = s_left[sp] + ret;
            --sp;
        }
    }
    return ret;
}
// (plain + on float / double compiles to one IEEE add each: no contraction is possible without a multiply)
__device__ __forceinline__ float numpy_pairwise_sum(const float* __restrict__ a, int n) { return numpy_pairwise_sum_t<float>(a, n); }

struct ShotTables {
    const int* cu_seg;            // [V+1]
    const int* cps;               // [total_seg][2] first / last frame (inclusive)
    const int* nfps;              // [total_seg]
    const int* picks;             // [total_rows], aligned with cu_rows
    const long long* cu_frames;   // [V+1]
    const int* capacity;          // [V] int(n_frames * proportion) / gcd
    const int* gcd;               // [V] gcd(capacity, weights...) computed by the host (1 is always valid)
    const long long* dp_off;      // [V] byte offsets into dp_scratch: 2 rows of (cap+1) int32, then n_seg x ceil((cap+1)/32) words
};

__global__ void __launch_bounds__(256)
keyshot_summary_kernel(const int* __restrict__ cu_rows, int S, ShotTables sh, const int* __restrict__ keep_count,
                       const float* __restrict__ keep_scores, const int* __restrict__ keep_boxes,
                       float* __restrict__ pos_score, float* __restrict__ frame_score, int* __restrict__ seg_score,
                       unsigned char* __restrict__ picked, unsigned char* __restrict__ summary,
                       unsigned char* __restrict__ dp_scratch) {
    __shared__ float s_ks[256];
    __shared__ int2 s_kb[256];
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const VidInfo vi = vid_info(cu_rows, v);
    const int T = vi.T;
    const size_t g0 = (size_t)vi.row0 * S;
    const int K = keep_count != nullptr ? keep_count[v] : 0;
    const int seg0 = sh.cu_seg[v], n_seg = sh.cu_seg[v + 1] - seg0;
    const long long f0 = sh.cu_frames[v];
    const int n_frames = (int)(sh.cu_frames[v + 1] - f0);
    const int* picks = sh.picks + vi.row0;
    float* pscore = pos_score + vi.row0;
    float* fscore = frame_score + f0;
    unsigned char* summ = summary + f0;

    // 1. per-position score: running max over the kept boxes (keep_count == nullptr: pos_score is an INPUT, e.g. the
    //    ground-truth importance scores of anchor_based/train.py:79, and this step is skipped)
    for (int t0 = 0; keep_count != nullptr && t0 < T; t0 += 256) {
        const int t = t0 + tid;
        float best = 0.f;
        for (int k0 = 0; k0 < K; k0 += 256) {
            __syncthreads();
            if (k0 + tid < K) {
                s_ks[tid] = keep_scores[g0 + k0 + tid];
                s_kb[tid] = reinterpret_cast<const int2*>(keep_boxes)[g0 + k0 + tid];
            }
            __syncthreads();
            const int n = min(256, K - k0);
            if (t < T)
                for (int q = 0; q < n; ++q)
                    if (t >= s_kb[q].x && t < s_kb[q].y) best = fmaxf(best, s_ks[q]);
        }
        if (t < T) pscore[t] = best;
    }
    __syncthreads();
    // 2. frame scores: frame f belongs to the last pick <= f
    for (int f = tid; f < n_frames; f += 256) {
        int lo = 0, hi = T;                                   // first index with picks[idx] > f
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (picks[mid] <= f) lo = mid + 1; else hi = mid; }
        fscore[f] = lo > 0 ? pscore[lo - 1] : 0.f;
        summ[f] = 0;
    }
    __syncthreads();
    // 3. shot scores
    for (int j = tid; j < n_seg; j += 256) {
        const int first = sh.cps[(seg0 + j) * 2], last = sh.cps[(seg0 + j) * 2 + 1];
        const int n = last - first + 1;
        int val = 0;
        if (n > 0 && first >= 0 && last < n_frames) {
            const float sum = numpy_pairwise_sum(fscore + first, n);
            const float mean = (float)((double)sum / (double)n);
            val = (int)__fmul_rn(1000.f, mean);
        }
        seg_score[seg0 + j] = val;
        picked[seg0 + j] = 0;
    }
    __syncthreads();
    // 4. knapsack by DP over the (gcd-reduced) capacity
    const int cap = sh.capacity[v], g = max(sh.gcd[v], 1);
    if (cap >= 0 && n_seg > 0) {
        int* row0 = reinterpret_cast<int*>(dp_scratch + sh.dp_off[v]);
        int* row1 = row0 + (cap + 1);
        const int words = (cap + 32) / 32;
        unsigned* bits = reinterpret_cast<unsigned*>(row1 + (cap + 1));
        for (int c = tid; c <= cap; c += 256) row0[c] = 0;
        __syncthreads();
        int* prev = row0;
        int* cur = row1;
        for (int i = 0; i < n_seg; ++i) {
            const int w = sh.nfps[seg0 + i] / g, val = seg_score[seg0 + i];
            for (int cb = 0; cb <= cap; cb += 256) {
                const int c = cb + tid;
                bool take = false;
                if (c <= cap) {
                    int b = prev[c];
                    if (c >= w && w >= 0) {
                        const int cand = prev[c - w] + val;
                        if (cand > b) { b = cand; take = true; }
                    }
                    cur[c] = b;
                }
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (lane == 0 && (c >> 5) < words) bits[(size_t)i * words + (c >> 5)] = m;
            }
            __syncthreads();
            int* t = prev; prev = cur; cur = t;
        }
        if (tid == 0) {
            int c = cap;
            for (int i = n_seg - 1; i >= 0; --i) {
                if ((bits[(size_t)i * words + (c >> 5)] >> (c & 31)) & 1u) {
                    picked[seg0 + i] = 1;
                    c -= sh.nfps[seg0 + i] / g;
                }
            }
        }
        __syncthreads();
    }
    // 5. mark the chosen shots
    for (int j = 0; j < n_seg; ++j) {
        if (!picked[seg0 + j]) continue;
        const int first = max(sh.cps[(seg0 + j) * 2], 0), last = min(sh.cps[(seg0 + j) * 2 + 1], n_frames - 1);
        for (int f = first + tid; f <= last; f += 256) summ[f] = 1;
    }
}
