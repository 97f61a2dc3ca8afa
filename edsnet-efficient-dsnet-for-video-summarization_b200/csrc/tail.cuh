// Scoring tail of DSNet.forward (anchor_based/dsnet.py:105-115): residual + LayerNorm(1024) -> fc1 ->
// D x shared {Linear(128,128), ReLU, LayerNorm(128)} -> multi-scale ROI average pooling -> cls / loc heads.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over 1024 features, one warp per row, two-pass statistics in registers (dsnet.py:89,106; eps 1e-5,
// biased variance as torch.nn.LayerNorm).  In: y = attn_out + x (already summed by the to_out epilogue).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm1024_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ out, int rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* src = y + (size_t)row * kFeat;
    float4 x[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = ldg4(src + (i * 32 + lane) * 4);
        s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
    const float mean = warp_sum(s) / (float)kFeat;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) / (float)kFeat + 1e-5f);
    float* dst = out + (size_t)row * kFeat;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = ldg4(gamma + c), b = ldg4(beta + c);
        st4(dst + c, make_float4(x[i].x * rstd * g.x + b.x, x[i].y * rstd * g.y + b.y,
                                 x[i].z * rstd * g.z + b.z, x[i].w * rstd * g.w + b.w));
    }
}

// Same LayerNorm, but the result is written directly as the row-scaled fp16 hi/lo operand planes (+ inverse row scale)
// that the tcgen05 fc1 GEMM consumes (layout of edsnet_split_f16): saves the fp32 round trip and one launch.
__global__ void __launch_bounds__(256)
layernorm1024_planes_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                            __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv_scale, int rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* src = y + (size_t)row * kFeat;
    float4 x[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = ldg4(src + (i * 32 + lane) * 4);
        s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
    const float mean = warp_sum(s) / (float)kFeat;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) / (float)kFeat + 1e-5f);
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = ldg4(gamma + c), b = ldg4(beta + c);
        x[i] = make_float4(x[i].x * rstd * g.x + b.x, x[i].y * rstd * g.y + b.y, x[i].z * rstd * g.z + b.z,
                           x[i].w * rstd * g.w + b.w);
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(x[i].x), fabsf(x[i].y)), fmaxf(fabsf(x[i].z), fabsf(x[i].w))));
    }
    mx = warp_max(mx);
    int e = 0;
    if (mx > 0.f && mx < INFINITY) e = 14 - ilogbf(mx);
    e = max(-100, min(100, e));
    const float sc = ldexpf(1.f, e);
    if (lane == 0) inv_scale[row] = ldexpf(1.f, -e);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float v[4] = {x[i].x * sc, x[i].y * sc, x[i].z * sc, x[i].w * sc};
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2half_rn(v[j]);
            l[j] = __float2half_rn(v[j] - __half2float(h[j]));
        }
        __half2 hh[2] = {__halves2half2(h[0], h[1]), __halves2half2(h[2], h[3])};
        __half2 ll[2] = {__halves2half2(l[0], l[1]), __halves2half2(l[2], l[3])};
        const size_t o = (size_t)row * kFeat + (i * 32 + lane) * 4;
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(hh);
        *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(ll);
    }
}

// ---------------------------------------------------------------------------------------------------------
// D applications of the ONE shared fc block (dsnet.py:91-96,107-108; eval mode: Dropout = identity).
// 64 rows per CTA; the 128x128 weight (transposed) and the activations stay in shared memory for all D rounds.
// ---------------------------------------------------------------------------------------------------------
constexpr int kLd128 = 132;
constexpr int kFcStackSmem = (128 * kLd128 + 64 * kLd128 + 3 * 128) * (int)sizeof(float);

__global__ void __launch_bounds__(256)
fc_stack_kernel(const float* __restrict__ u_in, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ u_out,
                int rows, int depth) {
    extern __shared__ __align__(16) float smem[];
    float* Wt = smem;                       // [k][n] = w[n][k]
    float* Us = Wt + 128 * kLd128;          // [row][k]
    float* bs = Us + 64 * kLd128;
    float* gs = bs + 128;
    float* es = gs + 128;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int r0 = blockIdx.x * 64;
    for (int idx = tid; idx < 128 * 32; idx += 256) {
        int n = idx >> 5, k4 = (idx & 31) * 4;
        float4 x = ldg4(w + n * 128 + k4);
        Wt[(k4 + 0) * kLd128 + n] = x.x; Wt[(k4 + 1) * kLd128 + n] = x.y;
        Wt[(k4 + 2) * kLd128 + n] = x.z; Wt[(k4 + 3) * kLd128 + n] = x.w;
    }
    for (int idx = tid; idx < 64 * 32; idx += 256) {
        int r = idx >> 5, k4 = (idx & 31) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < rows) x = ldg4(u_in + (size_t)(r0 + r) * kHidden + k4);
        st4(Us + r * kLd128 + k4, x);
    }
    if (tid < 128) { bs[tid] = __ldg(bias + tid); gs[tid] = __ldg(gamma + tid); es[tid] = __ldg(beta + tid); }
    __syncthreads();

    for (int d = 0; d < depth; ++d) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int k4 = 0; k4 < 128; k4 += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = lds4(Us + (ty * 4 + i) * kLd128 + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 b0 = lds4(Wt + (k4 + kk) * kLd128 + tx * 4);
                const float4 b1 = lds4(Wt + (k4 + kk) * kLd128 + 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                    acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                    acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                    acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
                }
            }
        }
        __syncthreads();                    // all reads of Us for this round are done
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                acc[i][j] = fmaxf(acc[i][j] + bs[c], 0.f);
                s += acc[i][j];
            }
            const float mean = half_warp_sum(s) / 128.f;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j] -= mean; q = fmaf(acc[i][j], acc[i][j], q); }
            const float rstd = 1.f / sqrtf(half_warp_sum(q) / 128.f + 1e-5f);
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                o[j] = acc[i][j] * rstd * gs[c] + es[c];
            }
            st4(Us + (ty * 4 + i) * kLd128 + tx * 4, make_float4(o[0], o[1], o[2], o[3]));
            st4(Us + (ty * 4 + i) * kLd128 + 64 + tx * 4, make_float4(o[4], o[5], o[6], o[7]));
        }
        __syncthreads();
    }
    for (int idx = tid; idx < 64 * 32; idx += 256) {
        int r = idx >> 5, k4 = (idx & 31) * 4;
        if (r0 + r < rows) st4(u_out + (size_t)(r0 + r) * kHidden + k4, lds4(Us + r * kLd128 + k4));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Multi-scale ROI average pooling + cls / loc heads (dsnet.py:78-80,111-115).
//   pooled[t][s] = (1/scale_s) * sum_{j = t - scale_s/2}^{t + scale_s/2 - 1} u[j]   (u[j] = 0 outside the video;
//   AvgPool1d(scale, 1, scale//2) with count_include_pad, last of the T+1 outputs dropped)
//   pred_cls = sigmoid(pooled . w_cls + b_cls),  pred_loc = pooled . W_loc + b_loc
// Pooling and the heads are both linear, so the three head projections are taken FIRST, once per feature row
// (d[t] = u[t] . {w_cls, w_loc0, w_loc1}), and the windows then run over 3 channels instead of 128:
// pooled . w = (1/s) sum_j d[j].  The kernel is a pure stream over u (512 B per row in, 12 S bytes per row out):
// one CTA per 128-row tile of one video (tiles[] = {video, first row}) plus a max_scale/2 halo; a warp takes four rows
// at a time (four independent 128-bit loads per lane in flight), the 12 partial dots are reduced with a packed
// butterfly (18 shuffles per 4 rows), then one thread per (row, scale) sums its window out of shared memory.
// ---------------------------------------------------------------------------------------------------------
struct ScaleList { int n; int s[kMaxScales]; };
constexpr int kRoiMaxHalo = 64;                                   // scales <= 128

// FROM_HEADS: `u` is the [rows][4] array of head projections the fc stack already emitted (fc_stack_tc_kernel), the
// first phase is a plain copy.
template <bool FROM_HEADS>
__global__ void __launch_bounds__(256)
roi_pool_heads_kernel(const float* __restrict__ u, const int* __restrict__ cu_rows, const int2* __restrict__ tiles,
                      ScaleList scales, int halo, const float* __restrict__ w_cls, const float* __restrict__ b_cls,
                      const float* __restrict__ w_loc, const float* __restrict__ b_loc,
                      float* __restrict__ pred_cls, float* __restrict__ pred_loc) {
    __shared__ float sd[3][128 + 2 * kRoiMaxHalo + 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int2 tile = tiles[blockIdx.x];
    const int v = tile.x, t0 = tile.y;
    const VidInfo vi = vid_info(cu_rows, v);
    const int nrows = 128 + 2 * halo;
    const float4 wc = ldg4(w_cls + lane * 4);
    const float4 w0 = ldg4(w_loc + lane * 4);
    const float4 w1 = ldg4(w_loc + kHidden + lane * 4);
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
    const int grp = (lane >> 3) & 3;                               // which of the 4 rows this lane ends up holding
    if (FROM_HEADS) {
        for (int r = tid; r < nrows; r += 256) {
            const int t = t0 - halo + r;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < vi.T) d = ldg4(u + (size_t)(vi.row0 + t) * 4);
            sd[0][r] = d.x; sd[1][r] = d.y; sd[2][r] = d.z;
        }
    }
    for (int r8 = warp * 8; !FROM_HEADS && r8 < nrows; r8 += 64) {
        // eight rows per warp and trip: eight independent 128-bit loads per lane in flight
        float4 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = t0 - halo + r8 + k;
            x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < vi.T && r8 + k < nrows) x[k] = ldg4(u + (size_t)(vi.row0 + t) * kHidden + lane * 4);
        }
#pragma unroll
        for (int g4 = 0; g4 < 2; ++g4) {
            const int r4 = r8 + g4 * 4;
            float p[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v = x[g4 * 4 + k];
                p[k * 3 + 0] = fmaf(v.x, wc.x, fmaf(v.y, wc.y, fmaf(v.z, wc.z, v.w * wc.w)));
                p[k * 3 + 1] = fmaf(v.x, w0.x, fmaf(v.y, w0.y, fmaf(v.z, w0.z, v.w * w0.w)));
                p[k * 3 + 2] = fmaf(v.x, w1.x, fmaf(v.y, w1.y, fmaf(v.z, w1.z, v.w * w1.w)));
            }
            // packed butterfly: 12 -> 6 values (lanes 16 apart), 6 -> 3 (8 apart), then plain butterflies over 4, 2, 1;
            // lanes 8 g .. 8 g + 7 end up with the three sums of row r4 + g
            float q[6], r[3];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float mine = b4 ? p[6 + i] : p[i], other = b4 ? p[i] : p[6 + i];
                q[i] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float mine = b3 ? q[3 + i] : q[i], other = b3 ? q[i] : q[3 + i];
                r[i] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
                for (int i = 0; i < 3; ++i) r[i] += __shfl_xor_sync(0xffffffffu, r[i], o);
            }
            if ((lane & 7) == 0 && r4 + grp < nrows) {
                sd[0][r4 + grp] = r[0];
                sd[1][r4 + grp] = r[1];
                sd[2][r4 + grp] = r[2];
            }
        }
    }
    const float bc = __ldg(b_cls), bl0 = __ldg(b_loc), bl1 = __ldg(b_loc + 1);
    __syncthreads();
    const int S = scales.n;
    for (int idx = tid; idx < 128 * S; idx += 256) {
        const int i = idx / S, si = idx - i * S;
        const int t = t0 + i;
        if (t >= vi.T) break;
        const int sc = scales.s[si];
        const int base = i + halo - sc / 2;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < sc; ++j) { a0 += sd[0][base + j]; a1 += sd[1][base + j]; a2 += sd[2][base + j]; }
        const float div = (float)sc;
        const size_t o = (size_t)(vi.row0 + t) * S + si;
        pred_cls[o] = 1.f / (1.f + expf(-(a0 / div + bc)));
        pred_loc[o * 2 + 0] = a1 / div + bl0;
        pred_loc[o * 2 + 1] = a2 / div + bl1;
    }
}
