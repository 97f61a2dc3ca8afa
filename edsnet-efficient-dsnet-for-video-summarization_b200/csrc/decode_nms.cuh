// Proposal decoding + temporal NMS (anchor_based/dsnet.py:140-153, anchor_helper.py:8-19,74-93,
// helpers/bbox_helper.py:21-31,49-70,80-118, evaluate.py:26-28).  Integer / index work: bit-exact parity is required.
#pragma once
#include "common.cuh"
#include "tail.cuh"   // ScaleList

// ---------------------------------------------------------------------------------------------------------
// decode: anchor (t, scale) + offsets (oc, ow) -> float32 [left, right] (what DSNet.predict returns) and the
// clipped / half-even-rounded int32 box evaluate.py:26 feeds to NMS.
// NumPy semantics reproduced: exp in float32; float32 x int32 products and the + anchor centre in float64;
// cast to float32; left/right = c -/+ w/2 in float32 (no fused multiply-add anywhere).
// The float32 exp is computed correctly rounded (double exp, one rounding); NumPy's SIMD expf may differ from
// that by 1 ulp for a few arguments -- see DESIGN.md "decode exp".
// grid (ceil(maxN/256), V)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decode_boxes_kernel(const float* __restrict__ pred_loc, const int* __restrict__ cu_rows, ScaleList scales,
                    float* __restrict__ boxes_f32, int* __restrict__ boxes_i32) {
    const int v = blockIdx.y;
    const VidInfo vi = vid_info(cu_rows, v);
    const int S = scales.n;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= vi.T * S) return;
    const int t = i / S, si = i - t * S;
    const size_t g = (size_t)vi.row0 * S + i;
    const float oc = __ldg(pred_loc + g * 2), ow = __ldg(pred_loc + g * 2 + 1);
    const double aw = (double)scales.s[si];
    const double c = (double)oc * aw + (double)t;
    const float e = (float)exp((double)ow);
    const double w = (double)e * aw;
    const float c32 = (float)c, w32 = (float)w;
    const float half = __fdiv_rn(w32, 2.0f);
    const float lo = __fsub_rn(c32, half), hi = __fadd_rn(c32, half);
    if (boxes_f32 != nullptr) { boxes_f32[g * 2] = lo; boxes_f32[g * 2 + 1] = hi; }
    const float Tf = (float)vi.T;
    if (boxes_i32 != nullptr) {
        boxes_i32[g * 2]     = (int)rintf(fminf(fmaxf(lo, 0.f), Tf));
        boxes_i32[g * 2 + 1] = (int)rintf(fminf(fmaxf(hi, 0.f), Tf));
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-video greedy NMS.  One CTA (256 threads) per video, everything for the video in shared memory:
//   1. key = (order-preserving bits of score) << 32 | anchor index, 0 for boxes with left >= right (dropped,
//      bbox_helper.py:91-93); bitonic sort, descending => visiting order of bbox_helper.py:95 with the tie rule
//      "higher index first" (stable ascending argsort reversed; the reference's own tie order is
//      implementation-defined).
//   2. chunks of 32 candidates in visiting order.  Phase A (all warps): every candidate (lane) is tested against
//      the kept list (warps stride over kept boxes) and the 32x32 intra-chunk suppression matrix is built with
//      ballots (warp w owns rows 4w..4w+3).  Phase B (warp 0): the chunk is resolved serially on register masks
//      (no arithmetic in the serial loop) and the survivors are appended.
//   overlap = max(0, min(r) - max(l)) / (max(r) - min(l)); the reference divides two int32 in float64 and keeps a
//   box iff quotient < thresh.  The division is replaced by one exact sign test, see nms_suppresses().
// Shared memory is sized by the host to the largest video of the launch (24 bytes per anchor rounded up to a power
// of two, at most 4096 anchors); larger videos use the global scratch the caller provides.
// ---------------------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 256;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kNmsSmemCap = 4096;

// true iff NOT (RN(inter / hull) < thresh), bit-for-bit what `iou < thresh` decides in float64.
// d = fma(-thresh, hull, inter) has the exact sign of inter - thresh * hull (one rounding cannot flip a sign), so
// d >= 0  <=>  inter/hull >= thresh  =>  RN(inter/hull) >= thresh (rounding is monotone, thresh is a double).
// For d < 0 the exact ratio is below thresh; the rounded quotient could still EQUAL thresh if the ratio is within
// half an ulp of it -- only then (never seen in practice) the real division is evaluated.
__device__ __forceinline__ bool nms_suppresses(int2 a, int2 b, double thresh, double slack) {
    const int inter = min(a.y, b.y) - max(a.x, b.x);
    if (inter <= 0) return !(0.0 < thresh);          // disjoint boxes: quotient exactly 0, no float64 work (the common case)
    const int hull = max(a.y, b.y) - min(a.x, b.x);
    const double di = (double)inter, dh = (double)hull;
    const double d = fma(-thresh, dh, di);
    if (d >= 0.0) return true;
    if (-d > slack * dh) return false;
    return !(di / dh < thresh);
}

__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(const float* __restrict__ scores, const int* __restrict__ boxes_i32, const int* __restrict__ cu_rows,
           int S, double thresh, int smem_cap, int skip_large, const long long* __restrict__ scratch_off,
           unsigned char* scratch,
           int* keep_count, int* keep_idx, float* keep_scores, int* keep_boxes) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ int s_nvalid, s_kept;
    __shared__ unsigned s_sup;
    __shared__ unsigned s_rowmask[32];
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const VidInfo vi = vid_info(cu_rows, v);
    const int N = vi.T * S;
    int P = 32;
    while (P < N) P <<= 1;
    unsigned long long* keys;
    int2* sbox;
    int2* kept;
    if (P > smem_cap && skip_large) return;       // handled by the multi-CTA path (nms_large.cuh)
    if (P <= smem_cap) {
        keys = reinterpret_cast<unsigned long long*>(nms_smem);
        sbox = reinterpret_cast<int2*>(nms_smem + (size_t)smem_cap * 8);
        kept = reinterpret_cast<int2*>(nms_smem + (size_t)smem_cap * 16);
    } else {
        unsigned char* base = scratch + scratch_off[v];
        keys = reinterpret_cast<unsigned long long*>(base);
        sbox = reinterpret_cast<int2*>(base + (size_t)P * 8);
        kept = reinterpret_cast<int2*>(base + (size_t)P * 16);
    }
    const size_t g0 = (size_t)vi.row0 * S;
    const int2* gbox = reinterpret_cast<const int2*>(boxes_i32) + g0;
    const double slack = fabs(thresh) * 4.440892098500626e-16;     // 2^-51 * |thresh| >= ulp(thresh)
    if (tid == 0) { s_nvalid = 0; s_kept = 0; s_sup = 0u; }
    __syncthreads();
    int local_valid = 0;
    for (int i = tid; i < P; i += kNmsThreads) {
        unsigned long long key = 0ull;
        if (i < N) {
            const int2 b = gbox[i];
            if (b.x < b.y) {
                unsigned u = __float_as_uint(scores[g0 + i]);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                key = ((unsigned long long)u << 32) | (unsigned)i;
                ++local_valid;
            }
        }
        keys[i] = key;
    }
    local_valid = (int)warp_sum((float)local_valid);       // <= 16 per thread * 32: exact in fp32
    if (lane == 0 && local_valid) atomicAdd(&s_nvalid, local_valid);
    __syncthreads();
    // bitonic sort, descending; pair p of a (k, j) step: i = p with a zero inserted at bit log2(j), partner i | j
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = tid; p < (P >> 1); p += kNmsThreads) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = keys[i], b = keys[ixj];
                const bool desc = (i & k) == 0;
                if ((a < b) == desc && a != b) { keys[i] = b; keys[ixj] = a; }
            }
            __syncthreads();
        }
    }
    const int nvalid = s_nvalid;
    for (int i = tid; i < nvalid; i += kNmsThreads) sbox[i] = gbox[(int)(keys[i] & 0xffffffffull)];
    __syncthreads();

    for (int base = 0; base < nvalid; base += 32) {
        const int cand = base + lane;
        const bool in_range = cand < nvalid;
        const int2 cb = in_range ? sbox[cand] : make_int2(0, 1);
        const int K = s_kept;
        // ---- phase A ----
        bool sup = false;
        for (int k = warp; k < K; k += kNmsWarps) sup = sup || nms_suppresses(kept[k], cb, thresh, slack);
        const unsigned m = __ballot_sync(0xffffffffu, sup && in_range);
        if (lane == 0 && m) atomicOr(&s_sup, m);
#pragma unroll
        for (int rr = 0; rr < 32 / kNmsWarps; ++rr) {
            const int r = warp * (32 / kNmsWarps) + rr;
            const int rx = __shfl_sync(0xffffffffu, cb.x, r), ry = __shfl_sync(0xffffffffu, cb.y, r);
            const bool bit = in_range && r < lane && nms_suppresses(make_int2(rx, ry), cb, thresh, slack);
            const unsigned rm = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) s_rowmask[r] = rm;                 // later candidates of the chunk that r suppresses
        }
        __syncthreads();
        // ---- phase B ----
        if (warp == 0) {
            const unsigned alive = __ballot_sync(0xffffffffu, in_range) & ~s_sup;
            const unsigned myrow = s_rowmask[lane];
            unsigned removed = 0u, keptmask = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const unsigned row_i = __shfl_sync(0xffffffffu, myrow, i);
                if (((alive & ~removed) >> i) & 1u) { keptmask |= 1u << i; removed |= row_i; }
            }
            if ((keptmask >> lane) & 1u) {
                const int pos = K + __popc(keptmask & ((1u << lane) - 1u));
                kept[pos] = cb;
                const int idx = (int)(keys[cand] & 0xffffffffull);
                keep_idx[g0 + pos] = idx;
                keep_scores[g0 + pos] = scores[g0 + idx];
                reinterpret_cast<int2*>(keep_boxes)[g0 + pos] = cb;
            }
            if (lane == 0) { s_kept = K + __popc(keptmask); s_sup = 0u; }
        }
        __syncthreads();
    }
    if (tid == 0) keep_count[v] = s_kept;
}
