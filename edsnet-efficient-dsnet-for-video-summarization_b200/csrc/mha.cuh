// Full multi-head attention base model (modules/models.py:12-74), the comparison backbone of BASELINE.json config 4:
// softmax(Q K^T / sqrt(d_k)) V per head, d_k = 128, 8 heads, eval mode (both Dropout(0.5) are the identity).
// fp32 CUDA-core flash attention: the (T x T) score matrix the reference materialises (134 MB at T = 2048) never
// exists; keys stream through shared memory in 64-row tiles with a running max / sum per query row.
// Input qkv [R][3072] = Q | K | V (each [R][1024], head h = columns h*128 .. h*128+127), output y [R][1024] head-merged.
#pragma once
#include "common.cuh"

constexpr int kMhaDk = 128;
constexpr int kMhaFeat = 1024;
constexpr int kMhaQkvCols = 3 * kMhaFeat;
constexpr int kLdD = 132;                 // padded stride of [64][128] tiles
constexpr int kMhaSmem = (64 * kLdD + 128 * kLd64 + 64 * kLd64 + 64 * kLdD) * (int)sizeof(float);

// grid (n_tiles64, 8 heads); tiles[] = {video, first query row}; 256 threads as a 16 x 16 grid, thread (ty, tx) owns
// query rows ty*4..+3 and, of the 64-key score tile, keys tx*4..+3; of the output, columns tx*4..+3 and 64+tx*4..+3.
__global__ void __launch_bounds__(256)
mha_flash_kernel(const float* __restrict__ qkv, const int* __restrict__ cu_rows, const int2* __restrict__ tiles,
                 float* __restrict__ y) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                     // [query][d]
    float* Kt = Qs + 64 * kLdD;           // [d][key]
    float* Ps = Kt + 128 * kLd64;         // [query][key]
    float* Vs = Ps + 64 * kLd64;          // [key][d]
    const int h = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int2 tile = tiles[blockIdx.x];
    const VidInfo vi = vid_info(cu_rows, tile.x);
    const int q0 = tile.y;
    const float* base = qkv + (size_t)vi.row0 * kMhaQkvCols + h * kMhaDk;
    for (int idx = tid; idx < 64 * 32; idx += 256) {
        const int r = idx >> 5, c4 = (idx & 31) * 4;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < vi.T) q = ldg4(base + (size_t)(q0 + r) * kMhaQkvCols + c4);
        st4(Qs + r * kLdD + c4, q);
    }
    float o[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
    float run_max[4], run_sum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { run_max[i] = -INFINITY; run_sum[i] = 0.f; }
    const float sqrt_dk = sqrtf((float)kMhaDk);

    for (int k0 = 0; k0 < vi.T; k0 += 64) {
        __syncthreads();                  // previous tile consumed (and the Q tile stored, first pass)
        for (int idx = tid; idx < 64 * 32; idx += 256) {
            const int r = idx >> 5, c4 = (idx & 31) * 4;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (k0 + r < vi.T) {
                kk = ldg4(base + (size_t)(k0 + r) * kMhaQkvCols + kMhaFeat + c4);
                vv = ldg4(base + (size_t)(k0 + r) * kMhaQkvCols + 2 * kMhaFeat + c4);
            }
            Kt[(c4 + 0) * kLd64 + r] = kk.x; Kt[(c4 + 1) * kLd64 + r] = kk.y;
            Kt[(c4 + 2) * kLd64 + r] = kk.z; Kt[(c4 + 3) * kLd64 + r] = kk.w;
            st4(Vs + r * kLdD + c4, vv);
        }
        __syncthreads();
        float s[4][4];
        zero44(s);
#pragma unroll 4
        for (int k4 = 0; k4 < kMhaDk; k4 += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = lds4(Qs + (ty * 4 + i) * kLdD + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) b[kk] = lds4(Kt + (k4 + kk) * kLd64 + tx * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    s[i][0] = fmaf(av[kk], b[kk].x, s[i][0]); s[i][1] = fmaf(av[kk], b[kk].y, s[i][1]);
                    s[i][2] = fmaf(av[kk], b[kk].z, s[i][2]); s[i][3] = fmaf(av[kk], b[kk].w, s[i][3]);
                }
            }
        }
        const int kvalid = vi.T - k0;
        float rescale[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[i][j] = (tx * 4 + j < kvalid) ? s[i][j] / sqrt_dk : -INFINITY;      // attn / sqrt(d_k), models.py:20
                mx = fmaxf(mx, s[i][j]);
            }
            mx = half_warp_max(mx);
            const float new_max = fmaxf(run_max[i], mx);
            rescale[i] = expf(run_max[i] - new_max);
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - new_max); ps += s[i][j]; }
            ps = half_warp_sum(ps);
            run_sum[i] = run_sum[i] * rescale[i] + ps;
            run_max[i] = new_max;
            st4(Ps + (ty * 4 + i) * kLd64 + tx * 4, make_float4(s[i][0], s[i][1], s[i][2], s[i][3]));
#pragma unroll
            for (int j = 0; j < 8; ++j) o[i][j] *= rescale[i];
        }
        __syncthreads();
#pragma unroll 4
        for (int k4 = 0; k4 < 64; k4 += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = lds4(Ps + (ty * 4 + i) * kLd64 + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 b0 = lds4(Vs + (k4 + kk) * kLdD + tx * 4);
                const float4 b1 = lds4(Vs + (k4 + kk) * kLdD + 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                    o[i][0] = fmaf(av, b0.x, o[i][0]); o[i][1] = fmaf(av, b0.y, o[i][1]);
                    o[i][2] = fmaf(av, b0.z, o[i][2]); o[i][3] = fmaf(av, b0.w, o[i][3]);
                    o[i][4] = fmaf(av, b1.x, o[i][4]); o[i][5] = fmaf(av, b1.y, o[i][5]);
                    o[i][6] = fmaf(av, b1.z, o[i][6]); o[i][7] = fmaf(av, b1.w, o[i][7]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r >= vi.T) continue;
        const float rs = 1.f / run_sum[i];
        float* dst = y + (size_t)(vi.row0 + r) * kMhaFeat + h * kMhaDk;
        st4(dst + tx * 4, make_float4(o[i][0] * rs, o[i][1] * rs, o[i][2] * rs, o[i][3] * rs));
        st4(dst + 64 + tx * 4, make_float4(o[i][4] * rs, o[i][5] * rs, o[i][6] * rs, o[i][7] * rs));
    }
}
