// Shared definitions for the EDSNet anchor-based scoring kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <mutex>
#include <set>
#include <utility>

// Function attributes (dynamic shared-memory opt-in) and device limits are per DEVICE, not per process: a DeviceOnce is
// true exactly once per (current device, tag), so a process that drives several GPUs configures each of them.  It
// holds the registry lock for as long as it lives -- `if (DeviceOnce once{&tag}) { opt-in }` keeps every other host
// thread out until the opt-in has actually run, so none of them can launch the kernel before its attribute is set.
class DeviceOnce {
    static std::mutex& mtx() { static std::mutex m; return m; }
    std::unique_lock<std::mutex> lock_;
    bool first_;
public:
    explicit DeviceOnce(const void* tag) : lock_(mtx()) {
        static std::set<std::pair<int, const void*>> seen;
        int dev = 0;
        cudaGetDevice(&dev);
        first_ = seen.insert({dev, tag}).second;
        if (!first_) lock_.unlock();
    }
    DeviceOnce(const DeviceOnce&) = delete;
    DeviceOnce& operator=(const DeviceOnce&) = delete;
    explicit operator bool() const { return first_; }
};

// Fixed geometry of the reference path (modules/models.py:135, anchor_based/dsnet.py:66-98).
constexpr int kHeads    = 8;      // num_head
constexpr int kDimHead  = 64;     // dim_head
constexpr int kLandmark = 64;     // num_landmarks
constexpr int kInner    = kHeads * kDimHead;      // 512
constexpr int kQkvCols  = 3 * kInner;             // 1536
constexpr int kFeat     = 1024;   // num_feature
constexpr int kHidden   = 128;    // num_hidden
constexpr int kTaps     = 33;     // residual_conv_kernel
constexpr int kPinvIters = 6;     // pinv_iterations
constexpr int kMaxScales = 8;
constexpr int kLd64 = 68;         // padded row stride (floats) of 64x64 smem tiles: 16B aligned, conflict-free as A and B operand

struct VidInfo {
    int row0;   // first packed row of this video
    int T;      // real rows
    int pad;    // zero rows the reference prepends (transformer/nystroformer.py:72-75)
    int n;      // padded length, multiple of 64
    int seg;    // rows per landmark segment = n / 64
};

__device__ __forceinline__ VidInfo vid_info(const int* __restrict__ cu_rows, int v) {
    VidInfo vi;
    vi.row0 = cu_rows[v];
    vi.T = cu_rows[v + 1] - vi.row0;
    vi.pad = (kLandmark - vi.T % kLandmark) % kLandmark;
    vi.n = vi.T + vi.pad;
    vi.seg = vi.n / kLandmark;
    return vi;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// reductions over the 16 lanes that share one 4-row strip in the 16x16 thread layout (tid = ty*16 + tx)
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// C(64x64) += A(64x64, row-major [i][k]) * B(64x64, row-major [k][j]); both in smem with stride kLd64.
// 256 threads, thread (ty,tx) owns rows ty*4..+3, cols tx*4..+3.
__device__ __forceinline__ void mm64_acc(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B,
                                         int ty, int tx) {
#pragma unroll 4
    for (int k4 = 0; k4 < 64; k4 += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = lds4(A + (ty * 4 + i) * kLd64 + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b[kk] = lds4(B + (k4 + kk) * kLd64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                acc[i][0] = fmaf(av[kk], b[kk].x, acc[i][0]);
                acc[i][1] = fmaf(av[kk], b[kk].y, acc[i][1]);
                acc[i][2] = fmaf(av[kk], b[kk].z, acc[i][2]);
                acc[i][3] = fmaf(av[kk], b[kk].w, acc[i][3]);
            }
        }
    }
}

__device__ __forceinline__ void zero44(float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// copy a 64x64 row-major global tile (ld_g floats) into smem [r][c] (stride kLd64); 256 threads
__device__ __forceinline__ void load64_rowmajor(float* __restrict__ S, const float* __restrict__ G, int ld_g, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int idx = tid + it * 256;          // 1024 float4
        int r = idx >> 4, c4 = (idx & 15) * 4;
        st4(S + r * kLd64 + c4, ldg4(G + (size_t)r * ld_g + c4));
    }
}
// same tile stored transposed: S[c][r] = G[r][c]
__device__ __forceinline__ void load64_transposed(float* __restrict__ S, const float* __restrict__ G, int ld_g, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        int idx = tid + it * 256;
        int r = idx >> 4, c4 = (idx & 15) * 4;
        float4 v = ldg4(G + (size_t)r * ld_g + c4);
        S[(c4 + 0) * kLd64 + r] = v.x;
        S[(c4 + 1) * kLd64 + r] = v.y;
        S[(c4 + 2) * kLd64 + r] = v.z;
        S[(c4 + 3) * kLd64 + r] = v.w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Dropout mask: Philox4x32-10, counter = (row, layer, offset_lo, offset_hi), key = (seed_lo, seed_hi); the 128 output
// bits decide the 128 hidden columns of one (row, layer): bit set = kept (p = 1/2, survivors scaled by 2).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ uint4 dropout_words(unsigned long long seed, unsigned long long offset, int row, int layer) {
    return philox4x32_10(make_uint4((uint32_t)row, (uint32_t)layer, (uint32_t)offset, (uint32_t)(offset >> 32)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

