// Nystrom landmark attention core (transformer/nystroformer.py:67-150) on packed variable-length videos.
// The reference front-pads every video with zero rows to a multiple of 64 (:72-75, unmasked).  Those rows are
// never materialised here: a zero row projects to q = k = v = 0, so it contributes nothing to a landmark sum,
// adds exp(0 - max) to the sim3 softmax denominator with a zero value row, and is a zero tap for the value
// convolution; its own output row is dropped (:144).  All kernels below index REAL rows only.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// landmarks: q_land / k_land [V][8][64][64] = mean over each block of `seg` padded rows (nystroformer.py:95-111).
// grid (64 landmarks, V), 256 threads; thread owns 4 consecutive columns of the 1024 q|k columns.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
landmarks_kernel(const float* __restrict__ qkv, const int* __restrict__ cu_rows,
                 float* __restrict__ q_land, float* __restrict__ k_land) {
    const int v = blockIdx.y, j = blockIdx.x;
    const VidInfo vi = vid_info(cu_rows, v);
    const int c4 = threadIdx.x * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int r_lo = j * vi.seg - vi.pad, r_hi = r_lo + vi.seg;      // real-row range of this segment
    if (r_lo < 0) r_lo = 0;
    for (int r = r_lo; r < r_hi; ++r) {
        float4 x = ldg4(qkv + (size_t)(vi.row0 + r) * kQkvCols + c4);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    const float div = (float)vi.seg;
    acc.x /= div; acc.y /= div; acc.z /= div; acc.w /= div;
    const int is_k = c4 >= kInner;
    const int cc = c4 & (kInner - 1);
    const int h = cc >> 6, d = cc & 63;
    float* dst = (is_k ? k_land : q_land) + ((((size_t)v * kHeads + h) * kLandmark + j) * kDimHead + d);
    st4(dst, acc);
}

// ---------------------------------------------------------------------------------------------------------
// attn2 = softmax(q_land k_land^T) per (video, head) (nystroformer.py:117,130) + the two magnitudes the
// pseudo-inverse start value needs (:16-18): stats[v][h] = {max_i sum_j |a_ij|, max_j sum_i |a_ij|}.
// grid (8, V), 256 threads.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn2_kernel(const float* __restrict__ q_land, const float* __restrict__ k_land,
             float* __restrict__ attn2, float* __restrict__ stats) {
    __shared__ __align__(16) float Qs[64 * kLd64];
    __shared__ __align__(16) float Kt[64 * kLd64];
    __shared__ float red[2][8];
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const size_t off = ((size_t)v * kHeads + h) * 4096;
    load64_rowmajor(Qs, q_land + off, 64, tid);
    load64_transposed(Kt, k_land + off, 64, tid);
    __syncthreads();
    float acc[4][4];
    zero44(acc);
    mm64_acc(acc, Qs, Kt, ty, tx);
    __syncthreads();                       // everyone done reading Qs before it is reused for the probabilities
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float mx = fmaxf(fmaxf(acc[i][0], acc[i][1]), fmaxf(acc[i][2], acc[i][3]));
        mx = half_warp_max(mx);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = expf(acc[i][j] - mx); s += acc[i][j]; }
        s = half_warp_sum(s);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = acc[i][j] / s;
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        st4(Qs + (ty * 4 + i) * kLd64 + tx * 4, o);
        st4(attn2 + off + (ty * 4 + i) * 64 + tx * 4, o);
    }
    __syncthreads();
    // row sums (threads 0..63) and column sums (threads 64..127) of |a| = a
    float val = 0.f;
    if (tid < 64) {
        for (int j = 0; j < 64; ++j) val += Qs[tid * kLd64 + j];
    } else if (tid < 128) {
        for (int i = 0; i < 64; ++i) val += Qs[i * kLd64 + (tid - 64)];
    }
    if (tid < 128) {
        val = warp_max(val);
        if ((tid & 31) == 0) red[tid >> 6][(tid >> 5) & 1] = val;
    }
    __syncthreads();
    if (tid == 0) {
        stats[((size_t)v * kHeads + h) * 2 + 0] = fmaxf(red[0][0], red[0][1]);
        stats[((size_t)v * kHeads + h) * 2 + 1] = fmaxf(red[1][0], red[1][1]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// a3v = softmax_over_keys(q_land k^T) v per (video, head) (nystroformer.py:118,130,133): streamed over 64-key
// tiles with a running max / sum, so the (64 x n) kernel is never materialised.  The zero pad keys are folded
// into the start state: logit 0, value 0.
// grid (8, V), 256 threads, dynamic smem 4 tiles.
// ---------------------------------------------------------------------------------------------------------
constexpr int kA3vSmem = 4 * 64 * kLd64 * (int)sizeof(float);

__global__ void __launch_bounds__(256)
a3v_kernel(const float* __restrict__ qkv, const int* __restrict__ cu_rows, const float* __restrict__ q_land,
           float* __restrict__ a3v) {
    extern __shared__ __align__(16) float smem[];
    float* Ql = smem;                    // [landmark][d]
    float* Kt = Ql + 64 * kLd64;         // [d][key]
    float* Ps = Kt + 64 * kLd64;         // [landmark][key]
    float* Vs = Ps + 64 * kLd64;         // [key][d]
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const VidInfo vi = vid_info(cu_rows, v);
    load64_rowmajor(Ql, q_land + ((size_t)v * kHeads + h) * 4096, 64, tid);

    float o[4][4];
    zero44(o);
    float run_max[4], run_sum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        run_max[i] = vi.pad > 0 ? 0.f : -INFINITY;
        run_sum[i] = (float)vi.pad;
    }
    const float* kbase = qkv + (size_t)vi.row0 * kQkvCols + kInner + h * kDimHead;
    const float* vbase = kbase + kInner;
    for (int r0 = 0; r0 < vi.T; r0 += 64) {
        __syncthreads();                 // previous tile fully consumed (also orders the Ql load on the first pass)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int idx = tid + it * 256;
            int r = idx >> 4, c4 = (idx & 15) * 4;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (r0 + r < vi.T) {
                kk = ldg4(kbase + (size_t)(r0 + r) * kQkvCols + c4);
                vv = ldg4(vbase + (size_t)(r0 + r) * kQkvCols + c4);
            }
            Kt[(c4 + 0) * kLd64 + r] = kk.x; Kt[(c4 + 1) * kLd64 + r] = kk.y;
            Kt[(c4 + 2) * kLd64 + r] = kk.z; Kt[(c4 + 3) * kLd64 + r] = kk.w;
            st4(Vs + r * kLd64 + c4, vv);
        }
        __syncthreads();
        float s[4][4];
        zero44(s);
        mm64_acc(s, Ql, Kt, ty, tx);
        const int kvalid = vi.T - r0;    // keys >= kvalid in this tile do not exist
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (tx * 4 + j >= kvalid) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
            mx = half_warp_max(mx);
            const float new_max = fmaxf(run_max[i], mx);
            const float rescale = expf(run_max[i] - new_max);      // exp(-inf) = 0 on the very first tile
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - new_max); ps += s[i][j]; }
            ps = half_warp_sum(ps);
            run_sum[i] = run_sum[i] * rescale + ps;
            run_max[i] = new_max;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= rescale;
            st4(Ps + (ty * 4 + i) * kLd64 + tx * 4, make_float4(s[i][0], s[i][1], s[i][2], s[i][3]));
        }
        __syncthreads();
        mm64_acc(o, Ps, Vs, ty, tx);
    }
    float* dst = a3v + ((size_t)v * kHeads + h) * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        st4(dst + (ty * 4 + i) * 64 + tx * 4,
            make_float4(o[i][0] / run_sum[i], o[i][1] / run_sum[i], o[i][2] / run_sum[i], o[i][3] / run_sum[i]));
}

// ---------------------------------------------------------------------------------------------------------
// Iterative Moore-Penrose pseudo-inverse of attn2 (nystroformer.py:13-28) followed by W = Z * a3v, all five
// 64x64 operands resident in shared memory for the whole 6-iteration chain.
// The start scale uses the maxima over ALL 8 heads of the video (the reference takes torch.max over the whole
// (1,8,64,64) tensor, :16-19) -- per VIDEO here, because the reference processes one video per call.
// grid (8, V), 256 threads, dynamic smem 5 tiles.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPinvSmem = 5 * 64 * kLd64 * (int)sizeof(float);

__device__ __forceinline__ void store_c_minus(float* __restrict__ D, const float (&acc)[4][4], float diag, int ty, int tx) {
    // D = diag * I - acc
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ((tx * 4 + j) == r ? diag : 0.f) - acc[i][j];
        st4(D + r * kLd64 + tx * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

__global__ void __launch_bounds__(256)
pinv_w_kernel(const float* __restrict__ attn2, const float* __restrict__ stats, const float* __restrict__ a3v,
              float* __restrict__ w_out, float* __restrict__ z_out, int iters) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Zs = As + 64 * kLd64;
    float* XZ = Zs + 64 * kLd64;
    float* Ts = XZ + 64 * kLd64;
    float* Us = Ts + 64 * kLd64;
    const int h = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const size_t off = ((size_t)v * kHeads + h) * 4096;

    float mrow = 0.f, mcol = 0.f;
#pragma unroll
    for (int hh = 0; hh < kHeads; ++hh) {
        mrow = fmaxf(mrow, __ldg(stats + ((size_t)v * kHeads + hh) * 2 + 0));
        mcol = fmaxf(mcol, __ldg(stats + ((size_t)v * kHeads + hh) * 2 + 1));
    }
    const float denom = mrow * mcol;
    // As = attn2, Zs = attn2^T / denom
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int idx = tid + it * 256;
        int r = idx >> 4, c4 = (idx & 15) * 4;
        float4 a = ldg4(attn2 + off + r * 64 + c4);
        st4(As + r * kLd64 + c4, a);
        Zs[(c4 + 0) * kLd64 + r] = a.x / denom; Zs[(c4 + 1) * kLd64 + r] = a.y / denom;
        Zs[(c4 + 2) * kLd64 + r] = a.z / denom; Zs[(c4 + 3) * kLd64 + r] = a.w / denom;
    }
    __syncthreads();
    float acc[4][4];
    for (int iter = 0; iter < iters; ++iter) {
        // XZ = A Z ; T = 7I - XZ
        zero44(acc);
        mm64_acc(acc, As, Zs, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st4(XZ + (ty * 4 + i) * kLd64 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        store_c_minus(Ts, acc, 7.f, ty, tx);
        __syncthreads();
        // U = 15I - XZ T
        zero44(acc);
        mm64_acc(acc, XZ, Ts, ty, tx);
        store_c_minus(Us, acc, 15.f, ty, tx);
        __syncthreads();
        // T = 13I - XZ U
        zero44(acc);
        mm64_acc(acc, XZ, Us, ty, tx);
        store_c_minus(Ts, acc, 13.f, ty, tx);      // Ts was last read before the previous barrier
        __syncthreads();
        // Z' = 0.25 Z T   (0.25 is a power of two: scaling before or after the product is bit-identical)
        zero44(acc);
        mm64_acc(acc, Zs, Ts, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st4(Us + (ty * 4 + i) * kLd64 + tx * 4,
                make_float4(0.25f * acc[i][0], 0.25f * acc[i][1], 0.25f * acc[i][2], 0.25f * acc[i][3]));
        __syncthreads();
        float* t = Zs; Zs = Us; Us = t;
    }
    if (z_out != nullptr) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int idx = tid + it * 256;
            int r = idx >> 4, c4 = (idx & 15) * 4;
            st4(z_out + off + r * 64 + c4, lds4(Zs + r * kLd64 + c4));
        }
    }
    // W = Z a3v
    load64_rowmajor(XZ, a3v + off, 64, tid);
    __syncthreads();
    zero44(acc);
    mm64_acc(acc, Zs, XZ, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        st4(w_out + off + (ty * 4 + i) * 64 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
}

// ---------------------------------------------------------------------------------------------------------
// out rows: softmax(q k_land^T) W + depth-wise value convolution, head-merged (nystroformer.py:115,130,133,
// 137-142).  One CTA per (64-row tile of one video, head); grid (n_tiles, 8).  tiles[] = {video, first real row} per tile.
// ---------------------------------------------------------------------------------------------------------
constexpr int kConvHalo = kTaps / 2;                        // 16
constexpr int kConvRows = 64 + 2 * kConvHalo;               // 96
constexpr int kAttnOutSmem = (3 * 64 * kLd64 + kConvRows * 64 + 64) * (int)sizeof(float);

__global__ void __launch_bounds__(256)
attn_out_kernel(const float* __restrict__ qkv, const int* __restrict__ cu_rows, const int2* __restrict__ tiles,
                const float* __restrict__ k_land, const float* __restrict__ w_mat,
                const float* __restrict__ conv_w, float* __restrict__ merged) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                    // q tile [row][d], later the probabilities [row][landmark]
    float* Kt = Qs + 64 * kLd64;         // k_land^T [d][landmark]
    float* Ws = Kt + 64 * kLd64;         // W [landmark][d]
    float* Vs = Ws + 64 * kLd64;         // value rows r0-16 .. r0+79, [row][64]
    float* taps = Vs + kConvRows * 64;
    const int h = blockIdx.y, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int2 tile = tiles[blockIdx.x];
    const int v = tile.x, r0 = tile.y;
    const VidInfo vi = vid_info(cu_rows, v);
    const size_t hoff = ((size_t)v * kHeads + h) * 4096;
    const float* qbase = qkv + (size_t)vi.row0 * kQkvCols + h * kDimHead;
    const float* vbase = qbase + 2 * kInner;

#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int idx = tid + it * 256;
        int r = idx >> 4, c4 = (idx & 15) * 4;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < vi.T) q = ldg4(qbase + (size_t)(r0 + r) * kQkvCols + c4);
        st4(Qs + r * kLd64 + c4, q);
    }
    load64_transposed(Kt, k_land + hoff, 64, tid);
    load64_rowmajor(Ws, w_mat + hoff, 64, tid);
    for (int idx = tid; idx < kConvRows * 16; idx += 256) {
        int r = idx >> 4, c4 = (idx & 15) * 4;
        int rr = r0 - kConvHalo + r;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr >= 0 && rr < vi.T) x = ldg4(vbase + (size_t)rr * kQkvCols + c4);
        st4(Vs + r * 64 + c4, x);
    }
    if (tid < kTaps) taps[tid] = __ldg(conv_w + h * kTaps + tid);
    __syncthreads();

    float s[4][4];
    zero44(s);
    mm64_acc(s, Qs, Kt, ty, tx);
    __syncthreads();                     // Qs is overwritten with the probabilities below
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
        mx = half_warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - mx); sum += s[i][j]; }
        sum = half_warp_sum(sum);
        st4(Qs + (ty * 4 + i) * kLd64 + tx * 4,
            make_float4(s[i][0] / sum, s[i][1] / sum, s[i][2] / sum, s[i][3] / sum));
    }
    __syncthreads();
    float o[4][4];
    zero44(o);
    mm64_acc(o, Qs, Ws, ty, tx);
    // + sum_t taps[t] * v[row + t - 16]: output row i of this thread is Vs row (ty*4 + i + t)
    float cv[4][4];
    zero44(cv);
#pragma unroll
    for (int rr = 0; rr < 4 + kTaps - 1; ++rr) {
        const float4 x = lds4(Vs + (ty * 4 + rr) * 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = rr - i;
            if (t >= 0 && t < kTaps) {
                const float w = taps[t];
                cv[i][0] = fmaf(w, x.x, cv[i][0]); cv[i][1] = fmaf(w, x.y, cv[i][1]);
                cv[i][2] = fmaf(w, x.z, cv[i][2]); cv[i][3] = fmaf(w, x.w, cv[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r < vi.T)
            st4(merged + (size_t)(vi.row0 + r) * kInner + h * kDimHead + tx * 4,
                make_float4(o[i][0] + cv[i][0], o[i][1] + cv[i][1], o[i][2] + cv[i][2], o[i][3] + cv[i][3]));
    }
}
