// GoogLeNet pool5 feature extraction (second half): the frame features the scoring path consumes,
// helpers/video_helper.py:27-73 -- torchvision's googlenet without its last two children (dropout, fc), eval mode, then
// feat / (|feat| + 1e-10).  Every convolution (conv + folded BatchNorm(eps 1e-3) + bias) is a product on the tcgen05
// GEMM of gemm_tc.cuh; this file holds what surrounds the products:
//
//   cnn_im2col_planes_kernel   patch gather (kh x kw x C, channel fastest, zero padding) over a VIRTUAL channel concat of
//                              up to four source buffers, ReLU of the producing layer applied on the way, straight into
//                              the GEMM's operand format: fp16 hi / lo planes with one power-of-two scale per row.  The
//                              four branch outputs of an inception module are therefore never concatenated in memory,
//                              and no activation is ever written back "activated".
//   cnn_maxpool_kernel         MaxPool2d(k, stride, pad, ceil_mode=True) over the same kind of input -> dense NHWC fp32
//   cnn_avgpool_l2norm_kernel  AdaptiveAvgPool2d(1) + the reference's L2 normalisation, one CTA per frame
//
// Activations are [pixels][ld] fp32 (NHWC, ld >= channels: GEMM outputs are padded to 128 columns); the input frames may
// be NCHW (general image / pixel / channel strides).
#pragma once
#include "common.cuh"

struct CnnSrc {
    const float* p;
    long long sn;        // stride between images
    int sp, sc;          // stride between pixels / channels
    int col0, ch;        // first column, channels taken from this source
};
struct CnnInput {
    CnnSrc s[4];
    int n_src, relu;
};

__device__ __forceinline__ float cnn_fetch(const CnnInput& in, int img, int pix, int c) {
    int s = 0;
    while (s + 1 < in.n_src && c >= in.s[s].ch) { c -= in.s[s].ch; ++s; }
    const CnnSrc& q = in.s[s];
    const float v = __ldg(q.p + (size_t)img * q.sn + (size_t)pix * q.sp + (size_t)(q.col0 + c) * q.sc);
    return in.relu ? fmaxf(v, 0.f) : v;
}

// four consecutive channels c .. c + 3 (host-checked: every source has channel stride 1, 4-aligned columns / counts /
// strides and a 16-byte aligned base, so a vector never straddles two sources)
__device__ __forceinline__ float4 cnn_fetch4(const CnnInput& in, int img, int pix, int c) {
    int s = 0;
    while (s + 1 < in.n_src && c >= in.s[s].ch) { c -= in.s[s].ch; ++s; }
    const CnnSrc& q = in.s[s];
    float4 v = ldg4(q.p + (size_t)img * q.sn + (size_t)pix * q.sp + (size_t)(q.col0 + c));
    if (in.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
}

// One warp per output pixel; K = kh * kw * C (k = (ky * kw + kx) * C + c), zero padded to kpad (multiple of 64).
// VEC = 4: lanes walk the flattened (tap, channel / 4) index with 16-byte loads and 8-byte plane stores; NCACHE > 0: the
// row has at most 32 * NCACHE vectors and stays in registers between the maximum and the conversion (one read of the
// input instead of two).  VEC = 1: the general scalar form (the network input: 3 channels, NCHW).
template <int VEC, int NCACHE>
__global__ void __launch_bounds__(256)
cnn_im2col_planes_kernel(CnnInput in, int C, int n_img, int H, int W, int kh, int kw, int stride, int pad, int OH, int OW,
                         int kpad, __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv) {
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const long long M = (long long)n_img * OH * OW;
    if (m >= M) return;
    const int img = (int)(m / (OH * OW)), r = (int)(m % (OH * OW));
    const int oy = r / OW, ox = r % OW;
    const int CV = C / VEC, KV = kh * kw * CV;                      // vectors per tap / per row
    auto fetch = [&](int v) -> float4 {                             // vector v of the row (zero outside the image)
        const int tap = v / CV, c = (v - tap * CV) * VEC;
        const int ky = tap / kw, kx = tap - ky * kw;
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (VEC == 4) return cnn_fetch4(in, img, iy * W + ix, c);
        return make_float4(cnn_fetch(in, img, iy * W + ix, c), 0.f, 0.f, 0.f);
    };
    float4 cache[NCACHE > 0 ? NCACHE : 1];
    float mx = 0.f;
    if (NCACHE > 0) {
#pragma unroll
        for (int t = 0; t < NCACHE; ++t) {
            const int v = lane + 32 * t;
            cache[t] = v < KV ? fetch(v) : make_float4(0.f, 0.f, 0.f, 0.f);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(cache[t].x), fabsf(cache[t].y))), fmaxf(fabsf(cache[t].z), fabsf(cache[t].w)));
        }
    } else {
        for (int v = lane; v < KV; v += 32) {
            const float4 x = fetch(v);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
        }
    }
    mx = warp_max(mx);
    int e = 0;
    if (mx > 0.f && mx < INFINITY) e = 14 - ilogbf(mx);
    e = max(-100, min(100, e));
    const float sc = ldexpf(1.f, e);
    if (lane == 0) inv[m] = ldexpf(1.f, -e);
    __half* ph = hi + (size_t)m * kpad;
    __half* pl = lo + (size_t)m * kpad;
    auto put = [&](int v, float4 x) {
        const float a[4] = {x.x * sc, x.y * sc, x.z * sc, x.w * sc};
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            h[j] = __float2half_rn(a[j]);
            l[j] = __float2half_rn(a[j] - __half2float(h[j]));
        }
        if (VEC == 4) {
            *reinterpret_cast<uint2*>(ph + 4 * v) = *reinterpret_cast<uint2*>(h);
            *reinterpret_cast<uint2*>(pl + 4 * v) = *reinterpret_cast<uint2*>(l);
        } else {
            ph[v] = h[0];
            pl[v] = l[0];
        }
    };
    if (NCACHE > 0) {
#pragma unroll
        for (int t = 0; t < NCACHE; ++t) {
            const int v = lane + 32 * t;
            if (v < KV) put(v, cache[t]);
        }
    } else {
        for (int v = lane; v < KV; v += 32) put(v, fetch(v));
    }
    for (int k = KV * VEC + lane; k < kpad; k += 32) { ph[k] = __float2half_rn(0.f); pl[k] = __float2half_rn(0.f); }
}

// Few input channels (the network input: 3 channels, 7 x 7 taps, K = 147): a row is at most 32 * NK values, lane l owns
// k = l, l + 32, ...  The (tap, channel) decomposition of a lane's k's is computed ONCE per warp and reused for the
// PIX output pixels the warp walks; a row lives in registers between its maximum and its conversion (one read of the
// input, no integer division per element: the general scalar kernel spent 4.2 of pool5's 13.2 ms per 256 frames here).
template <int NK, int PIX>
__global__ void __launch_bounds__(256)
cnn_im2col_smallc_kernel(CnnInput in, int C, int n_img, int H, int W, int kh, int kw, int stride, int pad, int OH, int OW,
                         int kpad, __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const long long M = (long long)n_img * OH * OW;
    const long long m0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * PIX;
    const int K = kh * kw * C;
    int dy[NK], dx[NK], ch[NK];
#pragma unroll
    for (int t = 0; t < NK; ++t) {
        const int k = lane + 32 * t;
        const int tap = k / C;
        ch[t] = k - tap * C;
        dy[t] = tap / kw;
        dx[t] = tap - dy[t] * kw;
        if (k >= K) dy[t] = -(1 << 20);                      // never inside the image
    }
    for (int p = 0; p < PIX; ++p) {
        const long long m = m0 + p;
        if (m >= M) return;
        const int img = (int)(m / (OH * OW)), r = (int)(m % (OH * OW));
        const int iy0 = (r / OW) * stride - pad, ix0 = (r % OW) * stride - pad;
        float v[NK];
        float mx = 0.f;
#pragma unroll
        for (int t = 0; t < NK; ++t) {
            const int iy = iy0 + dy[t], ix = ix0 + dx[t];
            v[t] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? cnn_fetch(in, img, iy * W + ix, ch[t]) : 0.f;
            mx = fmaxf(mx, fabsf(v[t]));
        }
        mx = warp_max(mx);
        int e = 0;
        if (mx > 0.f && mx < INFINITY) e = 14 - ilogbf(mx);
        e = max(-100, min(100, e));
        const float sc = ldexpf(1.f, e);
        if (lane == 0) inv[m] = ldexpf(1.f, -e);
        __half* ph = hi + (size_t)m * kpad;
        __half* pl = lo + (size_t)m * kpad;
#pragma unroll
        for (int t = 0; t < NK; ++t) {
            const int k = lane + 32 * t;
            if (k < kpad) {
                const float a = v[t] * sc;
                const __half h = __float2half_rn(a);
                ph[k] = h;
                pl[k] = __float2half_rn(a - __half2float(h));
            }
        }
        for (int k = 32 * NK + lane; k < kpad; k += 32) { ph[k] = __float2half_rn(0.f); pl[k] = __float2half_rn(0.f); }
    }
}

// MaxPool2d(k, stride, pad, ceil_mode=True): thread per (output pixel, VEC channels), channel fastest; windows are clipped
// to the image (the padding never wins a maximum)
template <int VEC>
__global__ void __launch_bounds__(256)
cnn_maxpool_kernel(CnnInput in, int C, int n_img, int H, int W, int k, int stride, int pad, int OH, int OW,
                   float* __restrict__ out) {
    const int CV = C / VEC;
    const long long total = (long long)n_img * OH * OW * CV;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int c = (int)(idx % CV) * VEC;
        const long long m = idx / CV;
        const int img = (int)(m / (OH * OW)), r = (int)(m % (OH * OW));
        const int oy = r / OW, ox = r % OW;
        const int y0 = max(0, oy * stride - pad), y1 = min(H, oy * stride - pad + k);
        const int x0 = max(0, ox * stride - pad), x1 = min(W, ox * stride - pad + k);
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                if (VEC == 4) {
                    const float4 v = cnn_fetch4(in, img, y * W + x, c);
                    best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y);
                    best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
                } else {
                    best.x = fmaxf(best.x, cnn_fetch(in, img, y * W + x, c));
                }
            }
        if (VEC == 4) st4(out + m * C + c, best);
        else out[m * C + c] = best.x;
    }
}

// mean over the HW pixels of every channel, then feat / (|feat|_2 + 1e-10) (video_helper.py:66-72); C <= 1024, one CTA
// per frame, thread <-> channels tid, tid + 256, ...
__global__ void __launch_bounds__(256)
cnn_avgpool_l2norm_kernel(CnnInput in, int C, int HW, float* __restrict__ out) {
    __shared__ float s_part[8];
    const int img = blockIdx.x, tid = threadIdx.x;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = tid + 256 * q;
        if (c < C) {
            float a = 0.f;
            for (int p = 0; p < HW; ++p) a += cnn_fetch(in, img, p, c);
            f[q] = a / (float)HW;
            ss = fmaf(f[q], f[q], ss);
        }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_part[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_part[w];
    const float d = sqrtf(tot) + 1e-10f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = tid + 256 * q;
        if (c < C) out[(size_t)img * C + c] = f[q] / d;
    }
}
