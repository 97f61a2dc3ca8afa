// Temporal NMS for videos with more anchors than one CTA's shared memory holds (> 4096; BASELINE.json config 5 has
// 65 536).  Same semantics and tie rule as nms_kernel (decode_nms.cuh); the work is spread over the whole GPU:
//   1. keys (score bits << 32 | anchor index, 0 = dropped) for all anchors, padded to a power of two
//   2. multi-CTA bitonic sort, descending: 4096-key tiles in shared memory, wider strides as global passes
//   3. candidates are processed in visiting order in blocks of 4096:
//      a. nms_filter_kernel  (many CTAs)  every candidate of the block against ALL boxes kept so far -> dead flags
//      b. nms_resolve_kernel (one CTA)    survivors compacted in order, greedy suppression among them in 32-candidate
//                                         chunks (ballot masks, as nms_kernel), kept boxes appended to the outputs
// The kept counter lives on the device; nothing synchronises with the host.
#pragma once
#include "decode_nms.cuh"

constexpr int kNmsTile = 4096;            // keys per CTA in the shared-memory sort stages
constexpr int kNmsBlock = 4096;           // candidates per filter / resolve round

__global__ void __launch_bounds__(256)
nmsl_keys_kernel(const float* __restrict__ scores, const int2* __restrict__ boxes, int N, int P,
                 unsigned long long* __restrict__ keys, int* __restrict__ counters) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) { counters[0] = 0; }                       // kept so far (nvalid is derived from the sorted keys)
    if (i >= P) return;
    unsigned long long key = 0ull;
    if (i < N) {
        const int2 b = boxes[i];
        if (b.x < b.y) {
            unsigned u = __float_as_uint(scores[i]);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            key = ((unsigned long long)u << 32) | (unsigned)i;
        }
    }
    keys[i] = key;
}

__device__ __forceinline__ void bitonic_cmpx(unsigned long long& a, unsigned long long& b, bool desc) {
    if ((a < b) == desc && a != b) { const unsigned long long t = a; a = b; b = t; }
}

// all (k, j) steps with j < kNmsTile for k in [k_lo, k_hi] on one 4096-key tile held in shared memory
__global__ void __launch_bounds__(512)
nmsl_sort_local_kernel(unsigned long long* __restrict__ keys, int k_lo, int k_hi) {
    __shared__ unsigned long long s[kNmsTile];
    const int base = blockIdx.x * kNmsTile, tid = threadIdx.x;
    for (int i = tid; i < kNmsTile; i += 512) s[i] = keys[base + i];
    __syncthreads();
    for (int k = k_lo; k <= k_hi; k <<= 1) {
        for (int j = min(k >> 1, kNmsTile >> 1); j > 0; j >>= 1) {
            for (int p = tid; p < (kNmsTile >> 1); p += 512) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const bool desc = ((base + i) & k) == 0;
                unsigned long long a = s[i], b = s[i | j];
                bitonic_cmpx(a, b, desc);
                s[i] = a; s[i | j] = b;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < kNmsTile; i += 512) keys[base + i] = s[i];
}

// one (k, j) step with j >= kNmsTile, straight on global memory
__global__ void __launch_bounds__(256)
nmsl_sort_global_kernel(unsigned long long* __restrict__ keys, int k, int j, int P) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= (P >> 1)) return;
    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
    unsigned long long a = keys[i], b = keys[i | j];
    const unsigned long long a0 = a;
    bitonic_cmpx(a, b, (i & k) == 0);
    if (a != a0) { keys[i] = a; keys[i | j] = b; }
}

__global__ void __launch_bounds__(256)
nmsl_gather_kernel(const unsigned long long* __restrict__ keys, const int2* __restrict__ boxes, int P,
                   int2* __restrict__ sbox) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const unsigned long long key = keys[i];
    sbox[i] = key ? boxes[(int)(key & 0xffffffffull)] : make_int2(0, 0);     // (0, 0) marks "no candidate"
}

// ---------------------------------------------------------------------------------------------------------
// Suppression stage.  Greedy NMS in visiting order is the unique solution of
//     kept(p)  <=>  valid(p) and no q before p with kept(q) and overlap(q, p) >= thresh,
// so it can be computed as a monotone fixpoint instead of a serial sweep: every undecided candidate looks at the
// candidates BEFORE it that would suppress it -- one of them kept: dead; all of them dead: kept; otherwise wait.  States
// only move undecided -> kept / dead and every decision is final, so rounds need no ordering among threads (a stale
// "undecided" read only postpones a decision), and the number of rounds is the depth of the suppression chains (a few
// dozen), not the number of kept boxes (thousands at 65 536 anchors).
// Who can suppress p = [l, r)?  For thresh > 0 the boxes must overlap (l_q < r, r_q > l) and inter >= thresh * hull
// bounds the partner's length by (r - l) / thresh, hence l_q > l - (r - l) / thresh: with the candidates counting-sorted
// by left end point, the partners of p are ONE contiguous slice of that order, scanned by a warp with coalesced loads.
//   nmsp_count / nmsp_scan / nmsp_scatter   counting sort by left end point (cells 0 .. T)
//   nmsp_round     warp per candidate, a few sweeps per launch; `left[]` counts the still undecided per launch
//   nmsp_cleanup   one CTA, visiting order: decides whatever the fixed number of round launches left undecided
//                  (exactness does not depend on the chains being short)
//   nmsp_emit      kept candidates compacted in visiting order (single-CTA scan), outputs written
// The same exact overlap test as everywhere (nms_suppresses).  thresh <= 0 (every pair suppresses) keeps the serial path.
// ---------------------------------------------------------------------------------------------------------
enum : unsigned char { NMSP_UNDECIDED = 0, NMSP_KEPT = 1, NMSP_DEAD = 2 };

__global__ void __launch_bounds__(256)
nmsp_count_kernel(const int2* __restrict__ sbox, int P, int T, int* __restrict__ cell_cnt, unsigned char* __restrict__ state) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int2 b = sbox[p];
    const bool valid = b.x < b.y;
    state[p] = valid ? NMSP_UNDECIDED : NMSP_DEAD;
    if (valid) atomicAdd(cell_cnt + min(b.x, T), 1);
}

// exclusive scan of cnt[0 .. n-1] into start[0 .. n] (start[n] = total) and cursor[0 .. n-1] = start; one CTA of 1024
__global__ void __launch_bounds__(1024)
nmsp_scan_kernel(const int* __restrict__ cnt, int n, int* __restrict__ start, int* __restrict__ cursor) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += cnt[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        s_warp[lane] = wi - w;                       // exclusive warp offsets
    }
    __syncthreads();
    int run = s_warp[warp] + incl - sum;
    for (int i = i0; i < i1; ++i) { start[i] = run; cursor[i] = run; run += cnt[i]; }
    if (tid == 1023) start[n] = run;
}

__global__ void __launch_bounds__(256)
nmsp_scatter_kernel(const int2* __restrict__ sbox, int P, int T, int* __restrict__ cursor, int2* __restrict__ lbox,
                    int* __restrict__ lrank) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int2 b = sbox[p];
    if (b.x >= b.y) return;
    const int slot = atomicAdd(cursor + min(b.x, T), 1);
    lbox[slot] = b;
    lrank[slot] = p;
}

// fate of candidate p given the states visible now; all lanes of the warp return the same value
__device__ __forceinline__ unsigned char nmsp_decide(int p, int2 cb, const int* __restrict__ cell_start,
                                                     const int2* __restrict__ lbox, const int* __restrict__ lrank,
                                                     const volatile unsigned char* state, int T, double thresh, double slack,
                                                     double inv_thresh, int lane) {
    const int len = cb.y - cb.x;
    const int reach = (int)fmin((double)len * inv_thresh + 2.0, 2.0e9);
    const int lo = max(0, cb.x - reach), hi = min(cb.y, T + 1);            // cells lo .. hi-1 hold every possible partner
    const int i0 = cell_start[lo], i1 = cell_start[hi];
    bool waits = false;
    for (int base = i0; base < i1; base += 32) {
        const int i = base + lane;
        bool hit_kept = false;
        if (i < i1) {
            const int q = lrank[i];
            if (q < p && nms_suppresses(lbox[i], cb, thresh, slack)) {
                const unsigned char st = state[q];
                hit_kept = st == NMSP_KEPT;
                waits = waits || st == NMSP_UNDECIDED;
            }
        }
        if (__any_sync(0xffffffffu, hit_kept)) return NMSP_DEAD;
    }
    return __any_sync(0xffffffffu, waits) ? NMSP_UNDECIDED : NMSP_KEPT;
}

// warp per candidate, `sweeps` passes per launch.  left[0] = undecided candidates after this launch (zeroed by the host
// side memset before the first round, re-counted by every launch into its own slot).
__global__ void __launch_bounds__(256)
nmsp_round_kernel(const int2* __restrict__ sbox, int nvalid_cap, const int* __restrict__ cell_start,
                  const int2* __restrict__ lbox, const int* __restrict__ lrank, unsigned char* state, int T, double thresh,
                  int sweeps, const int* __restrict__ left_prev, int* __restrict__ left_now) {
    if (left_prev != nullptr && *left_prev == 0) return;                    // converged in an earlier launch
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= nvalid_cap) return;
    const volatile unsigned char* vstate = state;
    if (vstate[p] != NMSP_UNDECIDED) return;
    const int2 cb = sbox[p];
    const double slack = fabs(thresh) * 4.440892098500626e-16, inv_thresh = 1.0 / thresh;
    unsigned char st = NMSP_UNDECIDED;
    for (int s = 0; s < sweeps && st == NMSP_UNDECIDED; ++s)
        st = nmsp_decide(p, cb, cell_start, lbox, lrank, vstate, T, thresh, slack, inv_thresh, lane);
    if (lane == 0) {
        if (st != NMSP_UNDECIDED) { state[p] = st; __threadfence(); }
        else atomicAdd(left_now, 1);
    }
}

// one CTA: whatever is still undecided, in visiting order (every earlier candidate is decided when its turn comes)
__global__ void __launch_bounds__(1024)
nmsp_cleanup_kernel(const int2* __restrict__ sbox, int nvalid_cap, const int* __restrict__ cell_start,
                    const int2* __restrict__ lbox, const int* __restrict__ lrank, unsigned char* state, int T, double thresh,
                    const int* __restrict__ left_last) {
    if (*left_last == 0) return;
    __shared__ int s_any, s_kept_hit;
    const int tid = threadIdx.x;
    const double slack = fabs(thresh) * 4.440892098500626e-16, inv_thresh = 1.0 / thresh;
    for (int base = 0; base < nvalid_cap; base += 1024) {
        if (tid == 0) s_any = 0;
        __syncthreads();
        const int mine = base + tid;
        if (mine < nvalid_cap && state[mine] == NMSP_UNDECIDED) atomicOr(&s_any, 1);
        __syncthreads();
        if (!s_any) continue;
        for (int p = base; p < min(base + 1024, nvalid_cap); ++p) {
            if (state[p] != NMSP_UNDECIDED) continue;                       // uniform: every thread reads the same byte
            const int2 cb = sbox[p];
            const int len = cb.y - cb.x;
            const int reach = (int)fmin((double)len * inv_thresh + 2.0, 2.0e9);
            const int lo = max(0, cb.x - reach), hi = min(cb.y, T + 1);
            const int i0 = cell_start[lo], i1 = cell_start[hi];
            if (tid == 0) s_kept_hit = 0;
            __syncthreads();
            for (int i = i0 + tid; i < i1; i += 1024) {
                const int q = lrank[i];
                if (q < p && state[q] == NMSP_KEPT && nms_suppresses(lbox[i], cb, thresh, slack)) s_kept_hit = 1;
            }
            __syncthreads();
            if (tid == 0) state[p] = s_kept_hit ? NMSP_DEAD : NMSP_KEPT;
            __syncthreads();
        }
    }
}

// one CTA: compact the kept candidates in visiting order and write the outputs
__global__ void __launch_bounds__(1024)
nmsp_emit_kernel(const unsigned long long* __restrict__ keys, const int2* __restrict__ sbox, int nvalid_cap,
                 const unsigned char* __restrict__ state, const float* __restrict__ scores, int* __restrict__ keep_count_v,
                 int* __restrict__ keep_idx, float* __restrict__ keep_scores, int* __restrict__ keep_boxes) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nvalid_cap + 1023) / 1024;
    const int i0 = min(nvalid_cap, tid * per), i1 = min(nvalid_cap, i0 + per);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += state[i] == NMSP_KEPT;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        s_warp[lane] = wi - w;
        if (lane == 31) keep_count_v[0] = wi;
    }
    __syncthreads();
    int pos = s_warp[warp] + incl - cnt;
    for (int i = i0; i < i1; ++i) {
        if (state[i] != NMSP_KEPT) continue;
        const int idx = (int)(keys[i] & 0xffffffffull);
        keep_idx[pos] = idx;
        keep_scores[pos] = scores[idx];
        reinterpret_cast<int2*>(keep_boxes)[pos] = sbox[i];
        ++pos;
    }
}

// grid (kNmsBlock / 256, slices): candidate c of the block against the slice's share of the kept boxes
__global__ void __launch_bounds__(256)
nmsl_filter_kernel(const int2* __restrict__ sbox, int block_base, int P, const int2* __restrict__ kept,
                   const int* __restrict__ counters, double thresh, unsigned char* __restrict__ dead) {
    __shared__ int2 sk[256];
    const int c = block_base + blockIdx.x * 256 + threadIdx.x;
    const int K = counters[0];
    const int2 cb = c < P ? sbox[c] : make_int2(0, 0);
    const bool cand = cb.x < cb.y;
    const double slack = fabs(thresh) * 4.440892098500626e-16;
    bool sup = false;
    for (int k0 = blockIdx.y * 256; k0 < K; k0 += gridDim.y * 256) {
        __syncthreads();
        if (k0 + threadIdx.x < K) sk[threadIdx.x] = kept[k0 + threadIdx.x];
        __syncthreads();
        const int n = min(256, K - k0);
        if (cand && !sup)
            for (int q = 0; q < n; ++q)
                if (nms_suppresses(sk[q], cb, thresh, slack)) { sup = true; break; }
    }
    if (sup) dead[c - block_base] = 1;
}

// one CTA: compact the block's live candidates in order, resolve them greedily, append the survivors
__global__ void __launch_bounds__(512)
nmsl_resolve_kernel(const unsigned long long* __restrict__ keys, const int2* __restrict__ sbox, int block_base, int P,
                    unsigned char* __restrict__ dead, const float* __restrict__ scores, double thresh,
                    int* __restrict__ counters, int2* __restrict__ kept, int* __restrict__ keep_idx,
                    float* __restrict__ keep_scores, int* __restrict__ keep_boxes) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    int2* cbox = reinterpret_cast<int2*>(rs_smem);                          // live candidates, visiting order
    int2* lkept = cbox + kNmsBlock;                                         // kept in this block
    int* cpos = reinterpret_cast<int*>(lkept + kNmsBlock);                  // position in the sorted array
    __shared__ int s_warp_cnt[16], s_live, s_kept;
    __shared__ unsigned s_sup, s_rowmask[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double slack = fabs(thresh) * 4.440892098500626e-16;
    // ---- order-preserving compaction of the live candidates (8 per thread, consecutive) ----
    int2 mine[8];
    bool live[8];
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int l = tid * 8 + q, c = block_base + l;
        mine[q] = c < P ? sbox[c] : make_int2(0, 0);
        live[q] = mine[q].x < mine[q].y && dead[l] == 0;
        dead[l] = 0;                                         // ready for the next round
        cnt += live[q];
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp_cnt[warp] = incl;
    if (tid == 0) { s_kept = 0; s_sup = 0u; }
    __syncthreads();
    int off = incl - cnt;
    for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
    if (tid == 511) s_live = off + cnt;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (live[q]) { cbox[off] = mine[q]; cpos[off] = block_base + tid * 8 + q; ++off; }
    __syncthreads();
    const int nlive = s_live;
    // ---- greedy suppression among the live candidates (same scheme as nms_kernel) ----
    for (int base = 0; base < nlive; base += 32) {
        const int cand = base + lane;
        const bool in_range = cand < nlive;
        const int2 cb = in_range ? cbox[cand] : make_int2(0, 1);
        const int K = s_kept;
        bool sup = false;
        for (int k = warp; k < K; k += 16) sup = sup || nms_suppresses(lkept[k], cb, thresh, slack);
        const unsigned m = __ballot_sync(0xffffffffu, sup && in_range);
        if (lane == 0 && m) atomicOr(&s_sup, m);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = warp * 2 + rr;
            const int rx = __shfl_sync(0xffffffffu, cb.x, r), ry = __shfl_sync(0xffffffffu, cb.y, r);
            const bool bit = in_range && r < lane && nms_suppresses(make_int2(rx, ry), cb, thresh, slack);
            const unsigned rm = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) s_rowmask[r] = rm;
        }
        __syncthreads();
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
                const int lpos = K + __popc(keptmask & ((1u << lane) - 1u));
                lkept[lpos] = cb;
                const int gpos = counters[0] + lpos;
                const int idx = (int)(keys[cpos[cand]] & 0xffffffffull);
                kept[gpos] = cb;
                keep_idx[gpos] = idx;
                keep_scores[gpos] = scores[idx];
                reinterpret_cast<int2*>(keep_boxes)[gpos] = cb;
            }
            if (lane == 0) { s_kept = K + __popc(keptmask); s_sup = 0u; }
        }
        __syncthreads();
    }
    if (tid == 0) counters[0] += s_kept;
}

__global__ void nmsl_finish_kernel(const int* __restrict__ counters, int* __restrict__ keep_count_v) {
    keep_count_v[0] = counters[0];
}

constexpr int kNmsResolveSmem = kNmsBlock * (8 + 8 + 4);

constexpr int kNmsScratchPerAnchor = 48;      // bytes of scratch per anchor (anchors rounded up to a power of two)
constexpr int kNmspRoundLaunches = 12, kNmspSweeps = 4;

// host: NMS of ONE video with N > 4096 anchors over T positions, P = N rounded up to a power of two.
// scratch (48 P bytes): keys 8P | boxes in visiting order 8P | [fixpoint path] boxes sorted by left end 8P | their ranks
// 4P | cell counts, starts, cursors 3 x 4 (P + 2) | states P | round counters -- [serial path, thresh <= 0] kept boxes
// 8P | dead flags 4096 | counter.
static cudaError_t launch_nms_large(const float* scores_v, const int* boxes_v, int N, int T, double thresh,
                                    unsigned char* scratch, int* keep_count_v, int* keep_idx_v, float* keep_scores_v,
                                    int* keep_boxes_v, cudaStream_t st) {
    int P = kNmsTile;
    while (P < N) P <<= 1;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(scratch);
    int2* sbox = reinterpret_cast<int2*>(scratch + (size_t)P * 8);
    const int2* boxes = reinterpret_cast<const int2*>(boxes_v);
    static const char tag = 0;
    if (DeviceOnce once_{&tag}) {
        cudaError_t ea = cudaFuncSetAttribute(nmsl_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsResolveSmem);
        if (ea != cudaSuccess) return ea;
    }
    int* counters = reinterpret_cast<int*>(scratch + (size_t)P * 47);          // [0] kept so far (serial path)
    nmsl_keys_kernel<<<(P + 255) / 256, 256, 0, st>>>(scores_v, boxes, N, P, keys, counters);
    nmsl_sort_local_kernel<<<P / kNmsTile, 512, 0, st>>>(keys, 2, kNmsTile);
    for (int k = 2 * kNmsTile; k <= P; k <<= 1) {
        for (int j = k >> 1; j >= kNmsTile; j >>= 1)
            nmsl_sort_global_kernel<<<(P / 2 + 255) / 256, 256, 0, st>>>(keys, k, j, P);
        nmsl_sort_local_kernel<<<P / kNmsTile, 512, 0, st>>>(keys, k, k);
    }
    nmsl_gather_kernel<<<(P + 255) / 256, 256, 0, st>>>(keys, boxes, P, sbox);
    if (thresh > 0.0 && T > 0 && T <= P) {
        // ---- fixpoint path ----
        int2* lbox = reinterpret_cast<int2*>(scratch + (size_t)P * 16);
        int* lrank = reinterpret_cast<int*>(scratch + (size_t)P * 24);
        int* cell_cnt = reinterpret_cast<int*>(scratch + (size_t)P * 28);
        int* cell_start = cell_cnt + (P + 2);
        int* cell_cur = cell_start + (P + 2);
        unsigned char* state = scratch + (size_t)P * 28 + 3 * sizeof(int) * (size_t)(P + 2);      // < 40 P + 24
        int* left = reinterpret_cast<int*>(scratch + (size_t)P * 42);                               // [launches + 1]
        const int cells = T + 1;                                                // left end points 0 .. T
        cudaError_t e = cudaMemsetAsync(cell_cnt, 0, sizeof(int) * (size_t)cells, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(left, 0, sizeof(int) * (kNmspRoundLaunches + 1), st);
        if (e != cudaSuccess) return e;
        nmsp_count_kernel<<<(P + 255) / 256, 256, 0, st>>>(sbox, P, T, cell_cnt, state);
        nmsp_scan_kernel<<<1, 1024, 0, st>>>(cell_cnt, cells, cell_start, cell_cur);
        nmsp_scatter_kernel<<<(P + 255) / 256, 256, 0, st>>>(sbox, P, T, cell_cur, lbox, lrank);
        // candidates are in visiting order with the dropped ones last, so ranks >= N never matter
        for (int r = 0; r < kNmspRoundLaunches; ++r)
            nmsp_round_kernel<<<(N + 7) / 8, 256, 0, st>>>(sbox, N, cell_start, lbox, lrank, state, T, thresh, kNmspSweeps,
                                                          r > 0 ? left + r - 1 : nullptr, left + r);
        nmsp_cleanup_kernel<<<1, 1024, 0, st>>>(sbox, N, cell_start, lbox, lrank, state, T, thresh,
                                               left + kNmspRoundLaunches - 1);
        nmsp_emit_kernel<<<1, 1024, 0, st>>>(keys, sbox, N, state, scores_v, keep_count_v, keep_idx_v, keep_scores_v,
                                            keep_boxes_v);
        return cudaGetLastError();
    }
    // ---- serial path (thresh <= 0: every pair suppresses, no locality to exploit) ----
    int2* kept = reinterpret_cast<int2*>(scratch + (size_t)P * 16);
    unsigned char* dead = scratch + (size_t)P * 24;
    cudaError_t e = cudaMemsetAsync(dead, 0, kNmsBlock, st);
    if (e != cudaSuccess) return e;
    const int n_blocks = (N + kNmsBlock - 1) / kNmsBlock;     // dropped boxes sort last; blocks beyond N hold only key 0
    for (int b = 0; b < n_blocks; ++b) {
        if (b > 0)
            nmsl_filter_kernel<<<dim3(kNmsBlock / 256, 32), 256, 0, st>>>(sbox, b * kNmsBlock, P, kept, counters, thresh,
                                                                        dead);
        nmsl_resolve_kernel<<<1, 512, kNmsResolveSmem, st>>>(keys, sbox, b * kNmsBlock, P, dead, scores_v, thresh, counters, kept,
                                               keep_idx_v, keep_scores_v, keep_boxes_v);
    }
    nmsl_finish_kernel<<<1, 1, 0, st>>>(counters, keep_count_v);
    return cudaGetLastError();
}
