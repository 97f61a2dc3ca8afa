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
// per-video greedy NMS.  One CTA (512 threads) per video:
//   1. key = (order-preserving bits of score) << 32 | anchor index, 0 for boxes with left >= right (dropped,
//      bbox_helper.py:91-93); bitonic sort, descending => visiting order of bbox_helper.py:95 with the tie rule
//      "higher index first" (stable ascending argsort reversed; the reference's own tie order is
//      implementation-defined).
//   2. chunks of 32 candidates in visiting order: (a) all 16 warps test the chunk against the kept list,
//      lane = candidate, warps stride over kept boxes; (b) warp 0 resolves the chunk with a 32x32 suppression
//      bitmask (shuffles) and appends the survivors.
//   overlap = max(0, min(r) - max(l)) / (max(r) - min(l)) in float64 (int32 / int32 true division in NumPy),
//   a candidate survives a kept box iff overlap < thresh.
// N <= 4096 boxes: everything in shared memory; larger videos use the global scratch the caller provides.
// ---------------------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 512;
constexpr int kNmsSmemCap = 4096;
constexpr int kNmsSmemBytes = kNmsSmemCap * (8 + 8 + 8);

__device__ __forceinline__ bool nms_suppresses(int2 a, int2 b, double thresh) {
    const int inter = min(a.y, b.y) - max(a.x, b.x);
    const int hull = max(a.y, b.y) - min(a.x, b.x);
    const double iou = (double)(inter > 0 ? inter : 0) / (double)hull;
    return !(iou < thresh);
}

__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(const float* __restrict__ scores, const int* __restrict__ boxes_i32, const int* __restrict__ cu_rows,
           int S, double thresh, const long long* __restrict__ scratch_off, unsigned char* scratch,
           int* keep_count, int* keep_idx, float* keep_scores, int* keep_boxes) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ int s_nvalid, s_kept;
    __shared__ unsigned s_sup;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const VidInfo vi = vid_info(cu_rows, v);
    const int N = vi.T * S;
    int P = 32;
    while (P < N) P <<= 1;
    unsigned long long* keys;
    int2* sbox;
    int2* kept;
    if (P <= kNmsSmemCap) {
        keys = reinterpret_cast<unsigned long long*>(nms_smem);
        sbox = reinterpret_cast<int2*>(nms_smem + (size_t)kNmsSmemCap * 8);
        kept = reinterpret_cast<int2*>(nms_smem + (size_t)kNmsSmemCap * 16);
    } else {
        unsigned char* base = scratch + scratch_off[v];
        keys = reinterpret_cast<unsigned long long*>(base);
        sbox = reinterpret_cast<int2*>(base + (size_t)P * 8);
        kept = reinterpret_cast<int2*>(base + (size_t)P * 16);
    }
    const size_t g0 = (size_t)vi.row0 * S;
    if (tid == 0) { s_nvalid = 0; s_kept = 0; }
    __syncthreads();
    int local_valid = 0;
    for (int i = tid; i < P; i += kNmsThreads) {
        unsigned long long key = 0ull;
        if (i < N) {
            const int lo = boxes_i32[(g0 + i) * 2], hi = boxes_i32[(g0 + i) * 2 + 1];
            if (lo < hi) {
                unsigned u = __float_as_uint(scores[g0 + i]);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                key = ((unsigned long long)u << 32) | (unsigned)i;
                ++local_valid;
            }
        }
        keys[i] = key;
    }
    if (local_valid) atomicAdd(&s_nvalid, local_valid);
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += kNmsThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool first_half = (i & k) == 0;
                    if ((a < b) == first_half && a != b) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int nvalid = s_nvalid;
    for (int i = tid; i < nvalid; i += kNmsThreads) {
        const int idx = (int)(keys[i] & 0xffffffffull);
        sbox[i] = make_int2(boxes_i32[(g0 + idx) * 2], boxes_i32[(g0 + idx) * 2 + 1]);
    }
    __syncthreads();

    for (int base = 0; base < nvalid; base += 32) {
        const int cand = base + lane;
        const bool in_range = cand < nvalid;
        const int2 cb = in_range ? sbox[cand] : make_int2(0, 1);
        const int K = s_kept;
        if (tid == 0) s_sup = 0u;
        __syncthreads();
        bool sup = false;
        for (int k = warp; k < K; k += kNmsThreads / 32) {
            const int2 kb = kept[k];
            if (in_range && min(cb.y, kb.y) > max(cb.x, kb.x)) sup = sup || nms_suppresses(kb, cb, thresh);
            else if (in_range && !(0.0 < thresh)) sup = true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, sup);
        if (lane == 0 && m) atomicOr(&s_sup, m);
        __syncthreads();
        if (warp == 0) {
            const bool alive = in_range && !((s_sup >> lane) & 1u);
            unsigned mask = 0u;       // earlier candidates of this chunk that would suppress me if kept
            for (int i = 0; i < 32; ++i) {
                const int ox = __shfl_sync(0xffffffffu, cb.x, i), oy = __shfl_sync(0xffffffffu, cb.y, i);
                if (i < lane && nms_suppresses(make_int2(ox, oy), cb, thresh)) mask |= 1u << i;
            }
            unsigned keptmask = 0u;
            for (int i = 0; i < 32; ++i) {
                const unsigned mine = (alive && !(mask & keptmask)) ? 1u : 0u;
                keptmask |= __shfl_sync(0xffffffffu, mine, i) << i;
            }
            if ((keptmask >> lane) & 1u) {
                const int pos = K + __popc(keptmask & ((1u << lane) - 1u));
                kept[pos] = cb;
                const int idx = (int)(keys[cand] & 0xffffffffull);
                keep_idx[g0 + pos] = idx;
                keep_scores[g0 + pos] = scores[g0 + idx];
                keep_boxes[(g0 + pos) * 2] = cb.x;
                keep_boxes[(g0 + pos) * 2 + 1] = cb.y;
            }
            if (lane == 0) s_kept = K + __popc(keptmask);
        }
        __syncthreads();
    }
    if (tid == 0) keep_count[v] = s_kept;
}
