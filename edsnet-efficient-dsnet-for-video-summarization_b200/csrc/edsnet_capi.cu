// C ABI of libedsnet_b200.so -- see include/edsnet_b200.h for the contract of every entry point.
#include "../../include/edsnet_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_f32.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "fc_stack_tc.cuh"
#include "attn_tc.cuh"
#include "pinv_tc.cuh"
#include "mha.cuh"
#include "mha_tc.cuh"
#include "nystrom.cuh"
#include "tail.cuh"
#include "decode_nms.cuh"
#include "nms_large.cuh"
#include "summary.cuh"
#include "eval.cuh"
#include "kts.cuh"
#include "train.cuh"
#include "cnn.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return EDSNET_E_CUDA;
}
#define CU_CHECK(expr, where)                                   \
    do {                                                        \
        cudaError_t e__ = (expr);                               \
        if (e__ != cudaSuccess) return cuda_fail(e__, where);   \
    } while (0)

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- optional per-stage CUDA-event timing (bench / profiling only; off by default) ----
enum Stage : int { ST_SPLIT = 0, ST_QKV, ST_LANDMARKS, ST_ATTN2, ST_A3V, ST_PINV, ST_CONV, ST_ATTN_OUT, ST_TO_OUT, ST_LN,
                   ST_FC1, ST_FC_STACK, ST_ROI, ST_DECODE, ST_NMS,
                   // training step (train.cuh)
                   ST_T_WSPLIT, ST_T_LOSS, ST_T_ROI_BWD, ST_T_FC_BWD, ST_T_SPLIT, ST_T_GEMM_DW, ST_T_GEMM_DX, ST_T_LN_BWD,
                   ST_T_ATTN_PREP, ST_T_ATTN_ROWS, ST_T_PINV_BWD, ST_T_ATTN2_BWD, ST_T_ATTN_KEYS, ST_T_FINISH, ST_T_ADAM,
                   ST_COUNT };
const char* const kStageNames[ST_COUNT] = {"split_f16", "to_qkv_gemm", "landmarks", "attn2_softmax", "a3v_stream",
                                           "pinv_w", "value_conv", "attn_out", "to_out_gemm", "layernorm1024",
                                           "fc1_gemm", "fc_stack", "roi_pool_heads", "decode_boxes", "nms",
                                           "train_weight_planes", "train_loss_grad", "train_roi_heads_bwd",
                                           "train_fc_stack_bwd", "train_split_planes", "train_gemm_dw", "train_gemm_dx",
                                           "train_ln1024_bwd", "train_attn_prep", "train_attn_bwd_rows", "train_pinv_bwd",
                                           "train_attn2_bwd", "train_attn_bwd_keys", "train_dqkv_finish", "train_adam"};
struct StageEvents { int stage; cudaEvent_t e0, e1; };
bool g_stage_timing = false;
std::vector<StageEvents> g_stage_events;

struct StageScope {
    cudaStream_t st;
    bool on;
    StageEvents ev;
    StageScope(int stage, cudaStream_t s) : st(s), on(g_stage_timing) {
        if (!on) return;
        ev.stage = stage;
        cudaEventCreate(&ev.e0);
        cudaEventCreate(&ev.e1);
        cudaEventRecord(ev.e0, st);
    }
    ~StageScope() {
        if (!on) return;
        cudaEventRecord(ev.e1, st);
        g_stage_events.push_back(ev);
    }
};

int check_cfg(const edsnet_config* cfg) {
    if (!cfg) return fail(EDSNET_E_ARG, "config is NULL");
    if (cfg->n_scales < 1 || cfg->n_scales > EDSNET_MAX_SCALES)
        return fail(EDSNET_E_ARG, "n_scales must be in 1..8");
    for (int i = 0; i < cfg->n_scales; ++i) {
        const int s = cfg->scales[i];
        if (s < 2 || s > 128) return fail(EDSNET_E_ARG, "anchor scale out of range 2..128");
        // anchor_based/dsnet.py:113-115: with an odd scale AvgPool1d yields T outputs, [:-1] leaves T-1 and the
        // reference's .view(seq_len, num_scales) raises.  Same failure surface here.
        if (s & 1) return fail(EDSNET_E_ARG, "odd anchor scale: the reference's view() fails for odd scales");
    }
    if (cfg->fc_depth < 0 || cfg->fc_depth > 64) return fail(EDSNET_E_ARG, "fc_depth out of range 0..64");
    if (cfg->precision < EDSNET_PREC_FP32 || cfg->precision > EDSNET_PREC_FP16X2)
        return fail(EDSNET_E_ARG, "unknown precision");
    if (cfg->base_model != EDSNET_BASE_NYSTROM && cfg->base_model != EDSNET_BASE_ATTENTION)
        return fail(EDSNET_E_UNSUPPORTED, "base model outside the accelerated path (nystromformer, attention)");
    return EDSNET_OK;
}

int check_batch(const edsnet_batch* b) {
    if (!b) return fail(EDSNET_E_ARG, "batch is NULL");
    if (b->n_videos < 1 || b->total_rows < 1 || b->max_rows < 1) return fail(EDSNET_E_ARG, "empty batch");
    if (!b->cu_rows || !b->tiles64 || !b->tiles128) return fail(EDSNET_E_ARG, "batch tables are NULL");
    if (b->cu_rows_host && (b->cu_rows_host[0] != 0 || b->cu_rows_host[b->n_videos] != b->total_rows))
        return fail(EDSNET_E_ARG, "cu_rows_host does not describe this batch");
    if (b->n_videos > 65535) return fail(EDSNET_E_ARG, "at most 65535 videos per call");
    return EDSNET_OK;
}

ScaleList make_scales(const edsnet_config* cfg, int* halo) {
    ScaleList sl;
    sl.n = cfg->n_scales;
    int mx = 0;
    for (int i = 0; i < kMaxScales; ++i) {
        sl.s[i] = i < cfg->n_scales ? cfg->scales[i] : 0;
        if (sl.s[i] > mx) mx = sl.s[i];
    }
    *halo = mx / 2;
    return sl;
}

template <typename K>
cudaError_t opt_in_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

int gemm_dispatch(int precision, int epilogue, const float* A, const void* A16, const float* B, const void* B16,
                  float* C, int M, int N, int K, const float* bias, const float* res, int qcols, cudaStream_t st,
                  int stage = ST_QKV, float* aux = nullptr, const float* aux2 = nullptr, const float* aux3 = nullptr,
                  int passes = 0) {
    // passes: tcgen05 MMA passes over the hi/lo operand planes; 0 = what the precision implies (fp16x3 -> 3, fp16x2 -> 2,
    // fp16 -> 1).  edsnet_forward asks for 3 where the fp16x2 mode keeps full split precision (fc1).
    if (passes == 0) passes = precision == EDSNET_PREC_FP16X3 ? 3 : (precision == EDSNET_PREC_FP16X2 ? 2 : 1);
    StageScope scope(stage, st);
    if (M < 1 || N < 1 || K < 1) return fail(EDSNET_E_ARG, "gemm: empty problem");
    if (epilogue < 0 || epilogue > EPI_LN_FOLD) return fail(EDSNET_E_ARG, "gemm: unknown epilogue");
    if (epilogue == EPI_RES_LNPLANES &&
        (precision == EDSNET_PREC_FP32 || N != kFeat || !bias || !res || !aux || !aux2 || !aux3))
        return fail(EDSNET_E_ARG, "gemm: the LayerNorm-plane epilogue needs a tcgen05 precision, N = 1024 and its operands");
    if (epilogue == EPI_RES_LNPLANES && ((reinterpret_cast<uintptr_t>(res) & 31u) || (reinterpret_cast<uintptr_t>(C) & 31u)))
        return fail(EDSNET_E_ARG, "gemm: the LayerNorm-plane epilogue reads the residual with 32-byte loads: res and C must be 32-byte aligned");
    if (epilogue == EPI_LN_FOLD && (precision == EDSNET_PREC_FP32 || K != kFeat || !bias || !aux || !aux2))
        return fail(EDSNET_E_ARG, "gemm: the LayerNorm-fold epilogue needs a tcgen05 precision, K = 1024 and its operands");
    if (epilogue == 4 && (precision == EDSNET_PREC_FP32 || !aux || N != kQkvCols))
        return fail(EDSNET_E_ARG, "gemm: the plane epilogue needs a tcgen05 precision, N = 1536 and the scale output");
    if (((epilogue == 2 || epilogue == 3) && !bias) || (epilogue == 3 && !res)) return fail(EDSNET_E_ARG, "gemm: epilogue operand is NULL");
    GemmEpiArgs ep{bias, res, N, qcols, nullptr, nullptr, aux, aux2, aux3};
    if (precision == EDSNET_PREC_FP32) {
        if (!A || !B || !C) return fail(EDSNET_E_ARG, "gemm: NULL operand");
        if (K % kGemmBK || N % 4) return fail(EDSNET_E_ARG, "gemm fp32: K must be a multiple of 16, N of 4");
        cudaError_t e;
        switch (epilogue) {
            case 0: e = launch_sgemm_nt<EPI_NONE>(A, K, B, K, C, N, M, N, K, ep, st); break;
            case 1: e = launch_sgemm_nt<EPI_QSCALE>(A, K, B, K, C, N, M, N, K, ep, st); break;
            case 2: e = launch_sgemm_nt<EPI_BIAS>(A, K, B, K, C, N, M, N, K, ep, st); break;
            default: e = launch_sgemm_nt<EPI_BIAS_RES>(A, K, B, K, C, N, M, N, K, ep, st); break;
        }
        CU_CHECK(e, "sgemm_nt_kernel");
        return EDSNET_OK;
    }
    if (!A16 || !B16 || !C) return fail(EDSNET_E_ARG, "gemm tcgen05: fp16 operand planes are NULL");
    ep.a_scale = split_scales(A16, M, K);
    ep.b_scale = split_scales(B16, N, K);
    std::string msg;
    cudaError_t e = launch_gemm_tc(passes, epilogue,
                                   reinterpret_cast<const __half*>(A16), reinterpret_cast<const __half*>(B16),
                                   C, M, N, K, ep, st, &msg);
    if (e != cudaSuccess) {
        g_err = "gemm_tc: " + (msg.empty() ? std::string(cudaGetErrorString(e)) : msg);
        return EDSNET_E_CUDA;
    }
    return EDSNET_OK;
}

// qkv: fp32 [R][1536] for EDSNET_PREC_FP32; for the tcgen05 precisions the hi plane [R][1536] fp16 followed by the lo
// plane, with qkv_inv [R][24] (gemm_tc.cuh EPI_QKV_PLANES).
// Key-range / row-range splits of the per-(video, head) attention kernels: with few videos in the batch (one long
// video: BASELINE config 5) a grid of (heads, videos) CTAs would leave most of the 148 SMs idle.
int a3v_split_cap(int n_videos) {                   // capacity the workspace is sized for
    return 4 * n_videos >= 74 ? 1 : (148 + 4 * n_videos - 1) / (4 * n_videos);
}
// Only when the longest video has at least kSplitMinRows rows: below that every video is summed in ONE fixed order
// whatever batch it travels in, which is what keeps a packed batch bit-identical to one-video-per-call scoring.
constexpr int kSplitMinRows = 2048;
int a3v_splits(int n_videos, int max_rows) {
    if (max_rows < kSplitMinRows) return 1;
    return std::max(1, std::min(a3v_split_cap(n_videos), (max_rows + 63) / 64));
}
int attn_out_splits(int n_videos, int max_rows) {
    if (8 * n_videos >= 148 || max_rows < kSplitMinRows) return 1;
    return std::max(1, std::min((296 + 8 * n_videos - 1) / (8 * n_videos), (max_rows + 127) / 128));
}

int nystrom_core_impl(int precision, const edsnet_batch* b, const float* qkv, const float* qkv_inv,
                      const float* conv_w, float* q_land, float* k_land, float* attn2, float* stats, float* a3v,
                      float* zmat, float* wmat, float* merged, cudaStream_t st, void* merged16 = nullptr,
                      float* a3_part = nullptr, bool merged_lo = true) {
    // merged16 != nullptr (tcgen05 precisions, edsnet_forward): `merged` only carries the attention part and the sum
    // with the value convolution leaves as the to_out operand planes in merged16 (edsnet_split_f16 layout, 512 columns)
    const bool tcp = precision != EDSNET_PREC_FP32;
    const __half* p_hi = reinterpret_cast<const __half*>(qkv);
    const __half* p_lo = p_hi + (size_t)b->total_rows * kQkvCols;
    CUtensorMap map_hi, map_lo;
    if (tcp) {
        if (!qkv_inv) return fail(EDSNET_E_ARG, "nystrom_core: tcgen05 precision needs the plane scales");
        std::string msg;
        if (!tc::make_map(&map_hi, p_hi, (uint64_t)b->total_rows, kQkvCols, 64, 64, &msg) ||
            !tc::make_map(&map_lo, p_lo, (uint64_t)b->total_rows, kQkvCols, 64, 64, &msg))
            return fail(EDSNET_E_CUDA, "nystrom_core: " + msg);
        static const char tc_tag = 0;
        if (DeviceOnce once_{&tc_tag}) {
            CU_CHECK(opt_in_smem(tc::attn_out_tc_kernel, tc::kAoSmemBytes), "smem opt-in attn_out_tc");
            CU_CHECK(opt_in_smem(tc::a3v_tc_kernel, tc::kA3SmemBytes), "smem opt-in a3v_tc");
            CU_CHECK(opt_in_smem(tc::value_conv_kernel, tc::kConvSmemBytes), "smem opt-in value_conv");
            CU_CHECK(opt_in_smem(tc::value_conv_tc_kernel, tc::kCvSmemBytes), "smem opt-in value_conv_tc");
            CU_CHECK(opt_in_smem(tc::pinv_w_tc_kernel, tc::kPinvTcSmemBytes), "smem opt-in pinv_w_tc");
            CU_CHECK(opt_in_smem(tc::pinv_w_tc2_kernel, tc::kPinv2SmemBytes), "smem opt-in pinv_w_tc2");
        }
    }
    static const char f32_tag = 0;
    if (DeviceOnce once_{&f32_tag}) {
        CU_CHECK(opt_in_smem(a3v_kernel, kA3vSmem), "smem opt-in a3v");
        CU_CHECK(opt_in_smem(pinv_w_kernel, kPinvSmem), "smem opt-in pinv");
        CU_CHECK(opt_in_smem(attn_out_kernel, kAttnOutSmem), "smem opt-in attn_out");
    }
    const int V = b->n_videos;
    {
        StageScope scope(ST_LANDMARKS, st);
        // zmat is not touched before the pinv kernel writes Z: its first word is the a3v kernel's work-item counter
        if (tcp) tc::landmarks_planes_kernel<<<dim3(kLandmark, V), 256, 0, st>>>(p_hi, p_lo, qkv_inv, b->cu_rows, q_land, k_land,
                                                                              reinterpret_cast<unsigned*>(zmat));
        else landmarks_kernel<<<dim3(kLandmark, V), 256, 0, st>>>(qkv, b->cu_rows, q_land, k_land);
        CU_CHECK(cudaGetLastError(), "landmarks_kernel");
    }
    // EDSNET_ATTN2_VARIANT=1: the stand-alone CUDA-core kernel also on the tcgen05 path (cross-check); by default the
    // a3v kernel produces attn2 and its magnitudes itself
    static const int attn2_variant = [] { const char* e = getenv("EDSNET_ATTN2_VARIANT"); return e ? atoi(e) : 0; }();
    const bool attn2_fused = tcp && attn2_variant == 0;
    if (!attn2_fused) {
        StageScope scope(ST_ATTN2, st);
        attn2_kernel<<<dim3(kHeads, V), 256, 0, st>>>(q_land, k_land, attn2, stats);
        CU_CHECK(cudaGetLastError(), "attn2_kernel");
    }
    {
        StageScope scope(ST_A3V, st);
        if (tcp) {
            // a3_part (edsnet_forward / edsnet_train_forward pass it): room for a3v_split_cap(V) key ranges per (video, head)
            const int z = a3_part ? a3v_splits(V, b->max_rows) : 1;
            // EDSNET_A3V_VARIANT=1: one CTA per work item instead of the persistent kernel (cross-check)
            static const int a3v_variant = [] { const char* e = getenv("EDSNET_A3V_VARIANT"); return e ? atoi(e) : 0; }();
            const int n_items = (kHeads / 2) * V * z;
            unsigned* ticket = (zmat != nullptr && a3v_variant == 0) ? reinterpret_cast<unsigned*>(zmat) : nullptr;
            tc::a3v_tc_kernel<<<ticket ? std::min(n_items, tc::num_sms()) : n_items, tc::kA3Threads, tc::kA3SmemBytes, st>>>(
                map_hi, map_lo, qkv_inv, b->cu_rows, q_land, a3v, a3_part, k_land, attn2_fused ? attn2 : nullptr, stats,
                ticket, V, z);
            if (z > 1) {
                CU_CHECK(cudaGetLastError(), "a3v_tc_kernel");
                tc::a3v_merge_kernel<<<dim3(kHeads, V), 256, 0, st>>>(a3_part, z, a3v);
            }
        } else {
            a3v_kernel<<<dim3(kHeads, V), 256, kA3vSmem, st>>>(qkv, b->cu_rows, q_land, a3v);
        }
        CU_CHECK(cudaGetLastError(), "a3v_kernel");
    }
    {
        StageScope scope(ST_PINV, st);
        // EDSNET_PINV_VARIANT=1: the round-1 chain (two heads per CTA, four products per iteration), kept as a cross-check
        static const int pinv_variant = [] { const char* e = getenv("EDSNET_PINV_VARIANT"); return e ? atoi(e) : 0; }();
        if (tcp && pinv_variant == 1) tc::pinv_w_tc_kernel<<<dim3(kHeads / 2, V), 256, tc::kPinvTcSmemBytes, st>>>(attn2, stats, a3v, wmat, zmat, kPinvIters);
        else if (tcp) tc::pinv_w_tc2_kernel<<<dim3(kHeads, V), 256, tc::kPinv2SmemBytes, st>>>(attn2, stats, a3v, wmat, zmat, kPinvIters);
        else pinv_w_kernel<<<dim3(kHeads, V), 256, kPinvSmem, st>>>(attn2, stats, a3v, wmat, zmat, kPinvIters);
        CU_CHECK(cudaGetLastError(), "pinv_w_kernel");
    }
    if (tcp) {
        {
            StageScope scope(ST_ATTN_OUT, st);
            tc::attn_out_tc_kernel<<<dim3(kHeads, V, attn_out_splits(V, b->max_rows)), 320, tc::kAoSmemBytes, st>>>(map_hi, map_lo, qkv_inv, b->cu_rows,
                                                                                  k_land, wmat, merged,
                                                                                  merged16 ? stats : nullptr);
            CU_CHECK(cudaGetLastError(), "attn_out_tc_kernel");
        }
        StageScope scope(ST_CONV, st);
        if (merged16) {
            // stats [V][8][2] is dead after the pinv kernel: slot 0 of every (video, head) now holds max|W|
            __half* m_hi = static_cast<__half*>(merged16);
            __half* m_lo = m_hi + (size_t)b->total_rows * kInner;
            float* m_inv = reinterpret_cast<float*>(m_lo + (size_t)b->total_rows * kInner);
            {
                // v planes in boxes of 32 rows: the 160-row window is 5 boxes
                CUtensorMap map32_hi, map32_lo;
                std::string msg;
                if (!tc::make_map(&map32_hi, p_hi, (uint64_t)b->total_rows, kQkvCols, 64, 32, &msg) ||
                    !tc::make_map(&map32_lo, p_lo, (uint64_t)b->total_rows, kQkvCols, 64, 32, &msg))
                    return fail(EDSNET_E_CUDA, "nystrom_core: " + msg);
                tc::value_conv_tc_kernel<<<std::min(b->n_tiles128, tc::num_sms()), 320, tc::kCvSmemBytes, st>>>(
                    map32_hi, map32_lo, qkv_inv, b->cu_rows, reinterpret_cast<const int2*>(b->tiles128), conv_w, merged,
                    stats, m_hi, m_lo, m_inv, merged_lo ? 1 : 0, b->n_tiles128);
            }
        } else {
            tc::value_conv_kernel<<<dim3(b->n_tiles128, 4), 256, tc::kConvSmemBytes, st>>>(
                p_hi, p_lo, qkv_inv, b->cu_rows, reinterpret_cast<const int2*>(b->tiles128), conv_w, merged);
        }
        CU_CHECK(cudaGetLastError(), "value_conv_kernel");
    } else {
        StageScope scope(ST_ATTN_OUT, st);
        attn_out_kernel<<<dim3(b->n_tiles64, kHeads), 256, kAttnOutSmem, st>>>(
            qkv, b->cu_rows, reinterpret_cast<const int2*>(b->tiles64), k_land, wmat, conv_w, merged);
        CU_CHECK(cudaGetLastError(), "attn_out_kernel");
    }
    return EDSNET_OK;
}

// heads_out != nullptr (tcgen05 precisions, edsnet_forward): the stack emits the three head projections per row and
// u_out may be nullptr
int fc_stack_impl(const edsnet_config* cfg, const edsnet_weights* w, const float* u_in, float* u_out, int rows,
                  cudaStream_t st, float* heads_out = nullptr) {
    StageScope scope(ST_FC_STACK, st);
    if (cfg->precision != EDSNET_PREC_FP32) {
        // tensor-core version (fp16 hi/lo split, fp32-grade) for both tcgen05 precisions: the block is too small
        // and too error-sensitive for a single-pass variant to be worth having
        if (!w->fcb_w16) return fail(EDSNET_E_ARG, "fc_stack: tcgen05 precision needs fcb_w16 (edsnet_split_f16)");
        CU_CHECK(launch_fc_stack_tc(u_in, w->fcb_w16, w->fcb_b, w->fcb_ln_w, w->fcb_ln_b, u_out, rows, cfg->fc_depth,
                                    st, w->cls_w, w->loc_w, heads_out), "fc_stack_tc_kernel");
        return EDSNET_OK;
    }
    static const char fcs_tag = 0;
    if (DeviceOnce once_{&fcs_tag}) CU_CHECK(opt_in_smem(fc_stack_kernel, kFcStackSmem), "smem opt-in fc_stack");
    fc_stack_kernel<<<(rows + 63) / 64, 256, kFcStackSmem, st>>>(u_in, w->fcb_w, w->fcb_b, w->fcb_ln_w,
                                                                  w->fcb_ln_b, u_out, rows, cfg->fc_depth);
    CU_CHECK(cudaGetLastError(), "fc_stack_kernel");
    return EDSNET_OK;
}

int roi_impl(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* b, const float* u,
             float* pred_cls, float* pred_loc, cudaStream_t st, bool from_heads = false) {
    int halo = 0;
    ScaleList sl = make_scales(cfg, &halo);
    if (halo > kRoiMaxHalo) return fail(EDSNET_E_UNSUPPORTED, "roi_pool_heads: anchor scale above 128");
    StageScope scope(ST_ROI, st);
    if (from_heads)
        roi_pool_heads_kernel<true><<<b->n_tiles128, 256, 0, st>>>(u, b->cu_rows, reinterpret_cast<const int2*>(b->tiles128),
                                                                   sl, halo, w->cls_w, w->cls_b, w->loc_w, w->loc_b,
                                                                   pred_cls, pred_loc);
    else
        roi_pool_heads_kernel<false><<<b->n_tiles128, 256, 0, st>>>(u, b->cu_rows, reinterpret_cast<const int2*>(b->tiles128),
                                                                    sl, halo, w->cls_w, w->cls_b, w->loc_w, w->loc_b,
                                                                    pred_cls, pred_loc);
    CU_CHECK(cudaGetLastError(), "roi_pool_heads_kernel");
    return EDSNET_OK;
}

}  // namespace

extern "C" {

const char* edsnet_last_error(void) { return g_err.c_str(); }
int edsnet_abi_version(void) { return EDSNET_ABI_VERSION; }

size_t edsnet_workspace_bytes(const edsnet_config* cfg, int32_t total_rows, int32_t n_videos,
                              edsnet_workspace_layout* layout) {
    edsnet_workspace_layout L;
    std::memset(&L, 0, sizeof(L));
    const size_t R = (size_t)(total_rows > 0 ? total_rows : 0), V = (size_t)(n_videos > 0 ? n_videos : 0);
    const size_t head_mat = V * kHeads * 4096 * sizeof(float);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const bool mha = cfg && cfg->base_model == EDSNET_BASE_ATTENTION;
    L.qkv = take(R * (mha ? kMhaQkvCols : kQkvCols) * sizeof(float));
    L.yn = L.qkv;                                    // LayerNorm output reuses the (dead) qkv region
    L.q_land = take(head_mat);
    L.k_land = take(head_mat);
    L.attn2 = take(head_mat);
    L.stats = take(V * kHeads * 2 * sizeof(float));
    L.qkv_inv = take(R * 24 * sizeof(float));
    L.a3v = take(head_mat);
    L.zmat = take(head_mat);
    L.wmat = take(head_mat);
    L.merged = take(R * (mha ? kMhaFeat : kInner) * sizeof(float));
    const bool tcp = cfg && cfg->precision != EDSNET_PREC_FP32;
    // tcgen05 precisions, Nystrom base: y leaves to_out as the fc1 operand planes (+ 4 bytes of scale per row)
    L.y = take(tcp ? split_f16_bytes(R, kFeat) : R * kFeat * sizeof(float));
    L.u0 = take(R * kHidden * sizeof(float));
    L.u1 = take(R * kHidden * sizeof(float));
    L.x16 = off;
    if (cfg && cfg->precision != EDSNET_PREC_FP32) take(split_f16_bytes(R, kFeat));
    L.zeros = take(kFeat * sizeof(float));
    L.mha16 = off;
    if (tcp && mha) L.mha16 = take(R * kMhaQkvCols * 4);           // q | k | v operand planes of the attention base (hi, lo)
    L.a3_part = off;
    if (tcp && !mha) L.a3_part = take(V * kHeads * (size_t)a3v_split_cap((int)V) * kLandmark * tc::kA3PartLd * sizeof(float));
    L.zstat = off;
    L.xstat = off;
    if (tcp && !mha) {
        L.zstat = take(R * 32 * sizeof(float));
        L.xstat = take(R * 2 * sizeof(float));
    }
    L.total = off;
    if (layout) *layout = L;
    return L.total;
}

size_t edsnet_split_f16_bytes(int64_t rows, int64_t cols) {
    if (rows < 0 || cols < 0) return 0;
    return split_f16_bytes((size_t)rows, (size_t)cols);
}

int edsnet_split_f16(const float* src, void* dst_hi_lo, int64_t rows, int64_t cols, void* stream) {
    if (!src || !dst_hi_lo || rows < 1 || rows > (1 << 30)) return fail(EDSNET_E_ARG, "split_f16: bad argument");
    if (cols != 128 && cols != 512 && cols != 1024)
        return fail(EDSNET_E_ARG, "split_f16: cols must be 128, 512 or 1024");
    StageScope scope(ST_SPLIT, static_cast<cudaStream_t>(stream));
    cudaError_t e = launch_split_f16(src, dst_hi_lo, (int)rows, (int)cols, static_cast<cudaStream_t>(stream));
    CU_CHECK(e, "split_f16_kernel");
    return EDSNET_OK;
}

int edsnet_debug_tc_status(int32_t reset) {
    int flag = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&flag, tc::g_timeout_flag, sizeof(int));
    if (e != cudaSuccess) { cuda_fail(e, "tc status read"); return -1; }
    if (reset && flag) {
        const int zero = 0;
        e = cudaMemcpyToSymbol(tc::g_timeout_flag, &zero, sizeof(int));
        if (e != cudaSuccess) { cuda_fail(e, "tc status reset"); return -1; }
    }
    return flag;
}

int edsnet_debug_stage_timing(int32_t enable) {
    for (auto& ev : g_stage_events) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); }
    g_stage_events.clear();
    g_stage_timing = enable != 0;
    return EDSNET_OK;
}

int edsnet_debug_stage_count(void) { return ST_COUNT; }
const char* edsnet_debug_stage_name(int32_t stage) {
    return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : "";
}

int edsnet_debug_stage_times(double* ms_sum, int32_t* launches, int32_t n) {
    if (!ms_sum || !launches || n < ST_COUNT) return fail(EDSNET_E_ARG, "stage_times: need ST_COUNT slots");
    for (int i = 0; i < n; ++i) { ms_sum[i] = 0.0; launches[i] = 0; }
    for (auto& ev : g_stage_events) {
        CU_CHECK(cudaEventSynchronize(ev.e1), "stage event sync");
        float ms = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, ev.e0, ev.e1), "stage event elapsed");
        ms_sum[ev.stage] += ms;
        launches[ev.stage] += 1;
    }
    for (auto& ev : g_stage_events) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); }
    g_stage_events.clear();
    return EDSNET_OK;
}

int edsnet_debug_set_tc_variant(int32_t variant) {
    if (variant < 0 || variant > 9 || variant == 1 || variant == 2) return fail(EDSNET_E_ARG, "tc variant must be 0 or 3..9");
    tc::variant_ref() = variant;
    return EDSNET_OK;
}

int edsnet_gemm(int32_t precision, int32_t epilogue, const float* A, const void* A16, const float* B,
                const void* B16, float* C, int32_t M, int32_t N, int32_t K, const float* bias, const float* res,
                int32_t qcols, void* stream) {
    if (epilogue > EPI_QKV_PLANES) return fail(EDSNET_E_ARG, "gemm: unknown epilogue");
    return gemm_dispatch(precision, epilogue, A, A16, B, B16, C, M, N, K, bias, res, qcols,
                         static_cast<cudaStream_t>(stream));
}

int edsnet_nystrom_core(int32_t precision, const edsnet_batch* batch, const float* qkv, const float* qkv_inv,
                        const float* res_conv_w, float* q_land,
                        float* k_land, float* attn2, float* stats, float* a3v, float* zmat, float* wmat,
                        float* merged, void* stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (!qkv || !res_conv_w || !q_land || !k_land || !attn2 || !stats || !a3v || !wmat || !merged)
        return fail(EDSNET_E_ARG, "nystrom_core: NULL operand");
    if (precision < EDSNET_PREC_FP32 || precision > EDSNET_PREC_FP16X2) return fail(EDSNET_E_ARG, "unknown precision");
    return nystrom_core_impl(precision, batch, qkv, qkv_inv, res_conv_w, q_land, k_land, attn2, stats, a3v, zmat, wmat,
                             merged, static_cast<cudaStream_t>(stream));
}

int edsnet_fc_stack(const edsnet_config* cfg, const edsnet_weights* w, const float* u_in, float* u_out,
                    int32_t rows, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!w || !u_in || !u_out || rows < 1) return fail(EDSNET_E_ARG, "fc_stack: bad argument");
    return fc_stack_impl(cfg, w, u_in, u_out, rows, static_cast<cudaStream_t>(stream));
}

int edsnet_roi_pool_heads(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch,
                          const float* u, float* pred_cls, float* pred_loc, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!w || !u || !pred_cls || !pred_loc) return fail(EDSNET_E_ARG, "roi_pool_heads: NULL operand");
    return roi_impl(cfg, w, batch, u, pred_cls, pred_loc, static_cast<cudaStream_t>(stream));
}

int edsnet_forward(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch, const float* x,
                   float* pred_cls, float* pred_loc, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!w || !x || !pred_cls || !pred_loc || !workspace) return fail(EDSNET_E_ARG, "forward: NULL operand");
    // the kernels read x and the workspace with 32-byte vector loads (rows are 4 KB, so any row of an aligned buffer is fine)
    if ((reinterpret_cast<uintptr_t>(x) & 31u) || (reinterpret_cast<uintptr_t>(workspace) & 31u))
        return fail(EDSNET_E_ARG, "forward: x and workspace must be 32-byte aligned");
    edsnet_workspace_layout L;
    const size_t need = edsnet_workspace_bytes(cfg, batch->total_rows, batch->n_videos, &L);
    if (workspace_bytes < need) return fail(EDSNET_E_WORKSPACE, "forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const int R = batch->total_rows;
    // the two-pass mode is defined (and measured against the 1e-3 bar) for the Nystrom base only
    const int prec = (cfg->precision == EDSNET_PREC_FP16X2 && cfg->base_model == EDSNET_BASE_ATTENTION)
                         ? EDSNET_PREC_FP16X3 : cfg->precision;
    const void* x16 = nullptr;
    if (prec != EDSNET_PREC_FP32) {
        const bool nys = cfg->base_model == EDSNET_BASE_NYSTROM;
        // LayerNorm(1024) folded into the to_out / fc1 epilogues: Nystrom base, split-fp16 precision (the single-pass
        // mode has no digits to spare for the fold's  acc - mean wgsum  and keeps the LayerNorm kernel)
        const bool fold = nys && (prec == EDSNET_PREC_FP16X3 || prec == EDSNET_PREC_FP16X2);
        if ((nys && (!w->to_qkv_w16 || !w->to_out_w16)) || (!fold && !w->fc1_w16) || !w->fcb_w16)
            return fail(EDSNET_E_ARG, "forward: tcgen05 precision needs the fp16 weight planes (edsnet_split_f16)");
        if (fold && (!w->fc1_fold_w16 || !w->fc1_fold_wgsum || !w->fc1_fold_b || !w->to_out_bc || !w->to_out_bounds))
            return fail(EDSNET_E_ARG, "forward: fp16x3 / fp16x2 need the LayerNorm-folded fc1 operands (fc1_fold_*, to_out_bounds)");
        {
            // operand planes of x; Nystrom base: also (mean, max|.|) per row for the to_out epilogue
            StageScope scope(ST_SPLIT, st);
            CU_CHECK(launch_split_f16(x, ws + L.x16, R, kFeat, st, fold ? reinterpret_cast<float2*>(ws + L.xstat) : nullptr,
                                      /*write_lo=*/!(nys && prec == EDSNET_PREC_FP16X2)),
                     "split_f16_kernel");
        }
        x16 = ws + L.x16;
    }
    if (cfg->base_model == EDSNET_BASE_ATTENTION) {
        // full multi-head attention base (modules/models.py:46-65): Q|K|V projection, flash attention, output
        // projection (bias-free) + residual x
        if (!w->mha_qkv_w || !w->mha_fc_w || (prec != EDSNET_PREC_FP32 && (!w->mha_qkv_w16 || !w->mha_fc_w16)))
            return fail(EDSNET_E_ARG, "forward: attention base needs the mha_* weights");
        rc = gemm_dispatch(prec, EPI_NONE, x, x16, w->mha_qkv_w, w->mha_qkv_w16, F(L.qkv), R, kMhaQkvCols, kFeat,
                           nullptr, nullptr, 0, st);
        if (rc) return rc;
        if (prec == EDSNET_PREC_FP32) {
            static const char mha_tag = 0;
            if (DeviceOnce once_{&mha_tag}) CU_CHECK(opt_in_smem(mha_flash_kernel, kMhaSmem), "smem opt-in mha_flash");
            StageScope scope(ST_A3V, st);
            mha_flash_kernel<<<dim3(batch->n_tiles64, kHeads), 256, kMhaSmem, st>>>(
                F(L.qkv), batch->cu_rows, reinterpret_cast<const int2*>(batch->tiles64), F(L.merged));
            CU_CHECK(cudaGetLastError(), "mha_flash_kernel");
        } else {
            // tensor-core attention core: Q | K | V as operand planes (one scale per row and 128-column head slice), then
            // flash attention with both contractions on tcgen05 (mha_tc.cuh)
            static const char mhat_tag = 0;
            if (DeviceOnce once_{&mhat_tag}) CU_CHECK(opt_in_smem(tc::mha_tc_kernel, tc::kMtSmemBytes), "smem opt-in mha_tc");
            __half* p_hi = reinterpret_cast<__half*>(ws + L.mha16);
            __half* p_lo = p_hi + (size_t)R * kMhaQkvCols;
            {
                StageScope scope(ST_LANDMARKS, st);
                tc::mha_planes_kernel<<<(R + 7) / 8, 256, 0, st>>>(F(L.qkv), p_hi, p_lo, F(L.qkv_inv), R);
                CU_CHECK(cudaGetLastError(), "mha_planes_kernel");
            }
            CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo;
            std::string msg;
            if (!tc::make_map(&mq_hi, p_hi, (uint64_t)R, kMhaQkvCols, 64, 128, &msg) ||
                !tc::make_map(&mq_lo, p_lo, (uint64_t)R, kMhaQkvCols, 64, 128, &msg) ||
                !tc::make_map(&mk_hi, p_hi, (uint64_t)R, kMhaQkvCols, 64, 64, &msg) ||
                !tc::make_map(&mk_lo, p_lo, (uint64_t)R, kMhaQkvCols, 64, 64, &msg))
                return fail(EDSNET_E_CUDA, "mha: " + msg);
            StageScope scope(ST_A3V, st);
            tc::mha_tc_kernel<<<dim3(batch->n_tiles128, kHeads), 320, tc::kMtSmemBytes, st>>>(
                mq_hi, mq_lo, mk_hi, mk_lo, F(L.qkv_inv), batch->cu_rows, reinterpret_cast<const int2*>(batch->tiles128),
                F(L.merged));
            CU_CHECK(cudaGetLastError(), "mha_tc_kernel");
        }
        const void* m16 = nullptr;
        if (prec != EDSNET_PREC_FP32) {
            rc = edsnet_split_f16(F(L.merged), ws + L.x16, R, kMhaFeat, stream);
            if (rc) return rc;
            m16 = ws + L.x16;
        }
        CU_CHECK(cudaMemsetAsync(ws + L.zeros, 0, kFeat * sizeof(float), st), "zero bias");
        rc = gemm_dispatch(prec, EPI_BIAS_RES, F(L.merged), m16, w->mha_fc_w, w->mha_fc_w16, F(L.y), R, kFeat, kMhaFeat,
                           F(L.zeros), x, 0, st, ST_TO_OUT);
        if (rc) return rc;
    } else {
        // 1. qkv = x Wqkv^T, q pre-scaled by 1/8                                   (nystroformer.py:82-91)
        //    (tcgen05: written as fp16 operand planes + per-(row, head) scales, see attn_tc.cuh)
        rc = gemm_dispatch(prec, prec == EDSNET_PREC_FP32 ? EPI_QSCALE : EPI_QKV_PLANES, x, x16, w->to_qkv_w, w->to_qkv_w16,
                           F(L.qkv), R, kQkvCols, kFeat, nullptr, nullptr, kInner, st, ST_QKV, F(L.qkv_inv));
        if (rc) return rc;
        // 2. landmark attention core -> merged heads                                (nystroformer.py:95-142)
        // (tcgen05: the x16 planes are dead after step 1; the front of that region takes merged (R x 512) as planes)
        void* merged16_out = prec != EDSNET_PREC_FP32 ? static_cast<void*>(ws + L.x16) : nullptr;
        rc = nystrom_core_impl(prec, batch, F(L.qkv), F(L.qkv_inv), w->res_conv_w, F(L.q_land), F(L.k_land), F(L.attn2), F(L.stats),
                               F(L.a3v), F(L.zmat), F(L.wmat), F(L.merged), st, merged16_out,
                               prec != EDSNET_PREC_FP32 ? F(L.a3_part) : nullptr, /*merged_lo=*/prec != EDSNET_PREC_FP16X2);
        if (rc) return rc;
        // 3. y = merged Wout^T + b + x                                              (nystroformer.py:143, dsnet.py:105)
        //    (fp16x3: z = y - mean(x row) - mean(bias) leaves as the fc1 operand planes with the row sums LayerNorm needs; step 4
        //    then is ONE GEMM, fc1(LN(y)) = rstd (z (W o gamma)^T - mean(z) rowsum(W o gamma)) + (W beta + b))
        const void* merged16 = merged16_out;
        if (prec == EDSNET_PREC_FP16X3 || prec == EDSNET_PREC_FP16X2) {
            rc = gemm_dispatch(prec, EPI_RES_LNPLANES, nullptr, merged16, nullptr, w->to_out_w16, F(L.y), R, kFeat, kInner,
                               w->to_out_bc, x, 0, st, ST_TO_OUT, F(L.zstat), F(L.xstat), w->to_out_bounds);
            if (rc) return rc;
            // fc1 keeps all three passes in the fp16x2 mode too: it is the projection whose operand rounding the five
            // LayerNorm blocks behind it amplify most (profiles/r02_precision_probe.log), and N = 128 makes it HBM bound
            rc = gemm_dispatch(prec, EPI_LN_FOLD, nullptr, ws + L.y, nullptr, w->fc1_fold_w16, F(L.u0), R, kHidden, kFeat,
                               w->fc1_fold_b, nullptr, 0, st, ST_FC1, F(L.zstat), w->fc1_fold_wgsum, nullptr, 3);
            if (rc) return rc;
        } else {
            rc = gemm_dispatch(prec, EPI_BIAS_RES, F(L.merged), merged16, w->to_out_w, w->to_out_w16, F(L.y), R, kFeat,
                               kInner, w->to_out_b, x, 0, st, ST_TO_OUT);
            if (rc) return rc;
        }
    }
    // 4. LayerNorm(1024) -> fc1                                                 (dsnet.py:106)
    const void* yn16 = nullptr;
    const bool ln_folded = (prec == EDSNET_PREC_FP16X3 || prec == EDSNET_PREC_FP16X2) && cfg->base_model == EDSNET_BASE_NYSTROM;
    if (!ln_folded) {
    {
        StageScope scope(ST_LN, st);
        if (prec != EDSNET_PREC_FP32) {
            // LayerNorm output goes straight into the operand planes of fc1 (no fp32 yn, no separate split)
            __half* hi = reinterpret_cast<__half*>(ws + L.x16);
            __half* lo = hi + (size_t)R * kFeat;
            layernorm1024_planes_kernel<<<(R + 7) / 8, 256, 0, st>>>(F(L.y), w->ln_w, w->ln_b, hi, lo,
                                                                     reinterpret_cast<float*>(lo + (size_t)R * kFeat), R);
            yn16 = ws + L.x16;
        } else {
            layernorm1024_kernel<<<(R + 7) / 8, 256, 0, st>>>(F(L.y), w->ln_w, w->ln_b, F(L.yn), R);
        }
        CU_CHECK(cudaGetLastError(), "layernorm1024_kernel");
    }
    rc = gemm_dispatch(prec, EPI_BIAS, F(L.yn), yn16, w->fc1_w, w->fc1_w16, F(L.u0), R, kHidden, kFeat, w->fc1_b,
                       nullptr, 0, st, ST_FC1);
    if (rc) return rc;
    }
    // 5. shared fc block x depth                                                (dsnet.py:107-108)
    //    (tcgen05: the stack's last layer emits the three head projections per row into the u1 region, 16 B per row,
    //    instead of the 512-byte rows; pooling and heads are linear, so the windows then run over those)
    const bool heads_fused = prec != EDSNET_PREC_FP32;
    rc = fc_stack_impl(cfg, w, F(L.u0), heads_fused ? nullptr : F(L.u1), R, st, heads_fused ? F(L.u1) : nullptr);
    if (rc) return rc;
    // 6. ROI pooling + heads                                                    (dsnet.py:110-115)
    return roi_impl(cfg, w, batch, F(L.u1), pred_cls, pred_loc, st, heads_fused);
}

size_t edsnet_kts_scratch_bytes(int32_t n) { return n > 0 ? kts_video_bytes(n) : 0; }

int edsnet_kts(const edsnet_batch* batch, const edsnet_kts_video* videos, const float* x, int32_t ncp_cap,
               int32_t m_fixed, double vmax, int32_t desc_rate, int32_t lmin, int32_t lmax, int32_t* n_cps,
               int32_t* cps, double* objective, void* scratch, void* stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (!videos || !n_cps || !cps || !scratch) return fail(EDSNET_E_ARG, "kts: NULL operand");
    if (lmin < 1 || lmax < lmin || desc_rate < 1) return fail(EDSNET_E_ARG, "kts: need 1 <= lmin <= lmax, desc_rate >= 1");
    const int nmax = batch->max_rows;
    const size_t dp_smem = 2 * (size_t)(nmax + 1) * sizeof(double);
    if (nmax > 12800) return fail(EDSNET_E_UNSUPPORTED, "kts: more than 12800 frames in one video");
    static_assert(sizeof(edsnet_kts_video) == sizeof(KtsVideo), "video record layout");
    const KtsVideo* vids = reinterpret_cast<const KtsVideo*>(videos);
    unsigned char* scr = static_cast<unsigned char*>(scratch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int V = batch->n_videos;
    if (x != nullptr) {
        const unsigned t16 = (unsigned)((nmax + 15) / 16);
        kts_gram_kernel<<<dim3(t16, t16, V), 256, 0, st>>>(x, vids, scr);
        CU_CHECK(cudaGetLastError(), "kts_gram_kernel");
    }
    kts_prefix_kernel<<<V, 256, 0, st>>>(vids, scr);
    CU_CHECK(cudaGetLastError(), "kts_prefix_kernel");
    kts_scatter_kernel<<<dim3((unsigned)(((size_t)nmax * nmax + 255) / 256), V), 256, 0, st>>>(vids, scr);
    CU_CHECK(cudaGetLastError(), "kts_scatter_kernel");
    static const char dp_tag = 0;
    if (DeviceOnce once_{&dp_tag}) CU_CHECK(opt_in_smem(kts_dp_kernel, 2 * (12800 + 1) * (int)sizeof(double)), "smem opt-in kts_dp");
    kts_dp_kernel<<<V, 1024, dp_smem, st>>>(vids, scr, ncp_cap, m_fixed, vmax, desc_rate, lmin, lmax, batch->cu_rows,
                                           n_cps, cps, objective);
    CU_CHECK(cudaGetLastError(), "kts_dp_kernel");
    return EDSNET_OK;
}

int edsnet_eval_metrics(const edsnet_batch* batch, const int64_t* cu_frames, const uint8_t* summary,
                        const edsnet_eval_truth* truth, const float* x, double* fscore, double* diversity,
                        double* user_f1, int32_t* counts, void* stream) {
    int rc = check_batch(batch);
    if (rc) return rc;
    if (!truth || !truth->cu_users || !truth->user_off || !truth->user_frames || !truth->user_summ || !truth->metric)
        return fail(EDSNET_E_ARG, "eval_metrics: truth tables are NULL");
    if (!cu_frames || !summary || !x || !fscore || !diversity || !user_f1 || !counts)
        return fail(EDSNET_E_ARG, "eval_metrics: NULL operand");
    EvalTruth tr{truth->cu_users, reinterpret_cast<const long long*>(truth->user_off), truth->user_frames,
                 truth->user_summ, truth->metric};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    eval_metrics_kernel<<<batch->n_videos, 256, 0, st>>>(batch->cu_rows, reinterpret_cast<const long long*>(cu_frames),
                                                        summary, tr, x, fscore, diversity, user_f1, counts);
    CU_CHECK(cudaGetLastError(), "eval_metrics_kernel");
    return EDSNET_OK;
}

int edsnet_keyshot_summary(const edsnet_config* cfg, const edsnet_batch* batch, const edsnet_shots* shots,
                           const int32_t* keep_count, const float* keep_scores, const int32_t* keep_boxes,
                           float* pos_scores, float* frame_scores, int32_t* seg_scores, uint8_t* picked,
                           uint8_t* summary, void* dp_scratch, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!shots || !shots->cu_seg || !shots->cps || !shots->nfps || !shots->picks || !shots->cu_frames ||
        !shots->capacity || !shots->gcd || !shots->dp_off)
        return fail(EDSNET_E_ARG, "keyshot_summary: shot tables are NULL");
    // keep_count == NULL: pos_scores is the input (per-position scores), keep_scores / keep_boxes are not read
    if ((keep_count && (!keep_scores || !keep_boxes)) || !pos_scores || !frame_scores || !seg_scores || !picked ||
        !summary || !dp_scratch)
        return fail(EDSNET_E_ARG, "keyshot_summary: NULL operand");
    ShotTables sh{shots->cu_seg, shots->cps, shots->nfps, shots->picks,
                  reinterpret_cast<const long long*>(shots->cu_frames), shots->capacity, shots->gcd,
                  reinterpret_cast<const long long*>(shots->dp_off)};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    keyshot_summary_kernel<<<batch->n_videos, 256, 0, st>>>(batch->cu_rows, cfg->n_scales, sh, keep_count, keep_scores,
                                                           keep_boxes, pos_scores, frame_scores, seg_scores, picked,
                                                           summary, static_cast<unsigned char*>(dp_scratch));
    CU_CHECK(cudaGetLastError(), "keyshot_summary_kernel");
    return EDSNET_OK;
}

// ---- GoogLeNet pool5 feature extraction: what surrounds the convolution products (cnn.cuh) ----
namespace {
int cnn_input(const edsnet_cnn_input* in, CnnInput* out, int* channels) {
    if (!in) return fail(EDSNET_E_ARG, "cnn: input is NULL");
    if (in->n_src < 1 || in->n_src > 4) return fail(EDSNET_E_ARG, "cnn: 1..4 source buffers");
    int C = 0;
    for (int i = 0; i < in->n_src; ++i) {
        const edsnet_cnn_src& q = in->src[i];
        if (!q.p || q.channels < 1 || q.col0 < 0) return fail(EDSNET_E_ARG, "cnn: bad source buffer");
        out->s[i] = CnnSrc{q.p, (long long)q.image_stride, q.pixel_stride, q.channel_stride, q.col0, q.channels};
        C += q.channels;
    }
    for (int i = in->n_src; i < 4; ++i) out->s[i] = CnnSrc{nullptr, 0, 0, 0, 0, 0};
    out->n_src = in->n_src;
    out->relu = in->relu ? 1 : 0;
    *channels = C;
    return EDSNET_OK;
}
// 16-byte vector path: channel-contiguous sources whose columns, counts, strides and bases are multiples of four floats
bool cnn_vec4(const CnnInput& ci, int C) {
    if (C % 4) return false;
    for (int i = 0; i < ci.n_src; ++i) {
        const CnnSrc& q = ci.s[i];
        if (q.sc != 1 || (q.ch & 3) || (q.col0 & 3) || (q.sp & 3) || (q.sn & 3) || (reinterpret_cast<uintptr_t>(q.p) & 15))
            return false;
    }
    return true;
}
int pool_out(int H, int k, int stride, int pad) {           // torch, ceil_mode=True
    int o = (H + 2 * pad - k + stride - 1) / stride + 1;
    if ((o - 1) * stride >= H + pad) --o;                   // the last window must start inside the image or its left padding
    return o;
}
}  // namespace

int edsnet_cnn_im2col(const edsnet_cnn_input* in, int32_t n_img, int32_t H, int32_t W, int32_t kh, int32_t kw,
                      int32_t stride, int32_t pad, int32_t kpad, void* planes, void* stream) {
    CnnInput ci;
    int C = 0;
    int rc = cnn_input(in, &ci, &C);
    if (rc) return rc;
    if (n_img < 1 || H < 1 || W < 1 || kh < 1 || kw < 1 || stride < 1 || pad < 0 || !planes)
        return fail(EDSNET_E_ARG, "cnn_im2col: bad geometry");
    if (kpad % 64 != 0 || kpad < kh * kw * C) return fail(EDSNET_E_ARG, "cnn_im2col: kpad must be a multiple of 64 >= kh * kw * C");
    const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
    if (OH < 1 || OW < 1) return fail(EDSNET_E_ARG, "cnn_im2col: empty output");
    const long long M = (long long)n_img * OH * OW;
    if (M > 0x7fffffffLL) return fail(EDSNET_E_ARG, "cnn_im2col: too many output pixels");
    __half* hi = static_cast<__half*>(planes);
    __half* lo = hi + (size_t)M * kpad;
    float* inv = reinterpret_cast<float*>(lo + (size_t)M * kpad);
    const unsigned blocks = (unsigned)((M + 7) / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cnn_vec4(ci, C)) {
        if (kh * kw * C <= 1024)
            cnn_im2col_planes_kernel<4, 8><<<blocks, 256, 0, st>>>(ci, C, n_img, H, W, kh, kw, stride, pad, OH, OW, kpad, hi, lo, inv);
        else
            cnn_im2col_planes_kernel<4, 0><<<blocks, 256, 0, st>>>(ci, C, n_img, H, W, kh, kw, stride, pad, OH, OW, kpad, hi, lo, inv);
    } else if (kh * kw * C <= 256) {
        // few channels (the network input): eight pixels per warp, the row in registers
        const unsigned blocks8 = (unsigned)((M + 63) / 64);
        cnn_im2col_smallc_kernel<8, 8><<<blocks8, 256, 0, st>>>(ci, C, n_img, H, W, kh, kw, stride, pad, OH, OW, kpad, hi, lo, inv);
    } else {
        cnn_im2col_planes_kernel<1, 0><<<blocks, 256, 0, st>>>(ci, C, n_img, H, W, kh, kw, stride, pad, OH, OW, kpad, hi, lo, inv);
    }
    CU_CHECK(cudaGetLastError(), "cnn_im2col_planes_kernel");
    return EDSNET_OK;
}

int edsnet_cnn_maxpool(const edsnet_cnn_input* in, int32_t n_img, int32_t H, int32_t W, int32_t k, int32_t stride,
                       int32_t pad, float* out, void* stream) {
    CnnInput ci;
    int C = 0;
    int rc = cnn_input(in, &ci, &C);
    if (rc) return rc;
    if (n_img < 1 || H < 1 || W < 1 || k < 1 || stride < 1 || pad < 0 || 2 * pad > k || !out)
        return fail(EDSNET_E_ARG, "cnn_maxpool: bad geometry");
    const int OH = pool_out(H, k, stride, pad), OW = pool_out(W, k, stride, pad);
    const long long total = (long long)n_img * OH * OW * C;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
    if (cnn_vec4(ci, C) && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        cnn_maxpool_kernel<4><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(ci, C, n_img, H, W, k, stride, pad, OH, OW, out);
    else
        cnn_maxpool_kernel<1><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(ci, C, n_img, H, W, k, stride, pad, OH, OW, out);
    CU_CHECK(cudaGetLastError(), "cnn_maxpool_kernel");
    return EDSNET_OK;
}

int edsnet_cnn_avgpool_l2norm(const edsnet_cnn_input* in, int32_t n_img, int32_t HW, float* out, void* stream) {
    CnnInput ci;
    int C = 0;
    int rc = cnn_input(in, &ci, &C);
    if (rc) return rc;
    if (n_img < 1 || HW < 1 || C > 1024 || !out) return fail(EDSNET_E_ARG, "cnn_avgpool_l2norm: bad geometry (C <= 1024)");
    cnn_avgpool_l2norm_kernel<<<n_img, 256, 0, static_cast<cudaStream_t>(stream)>>>(ci, C, HW, out);
    CU_CHECK(cudaGetLastError(), "cnn_avgpool_l2norm_kernel");
    return EDSNET_OK;
}

int edsnet_forward_launches(const edsnet_config* cfg) {
    if (!cfg) return -1;
    // qkv, 5 x nystrom core (4 on the tcgen05 path: attn2 is produced by the a3v kernel), to_out, layernorm, fc1, fc stack,
    // roi+heads; tcgen05 modes add the operand split of x
    // (and of merged for the attention base) and run the value convolution as its own kernel; fp16x3 with the Nystrom
    // base has no LayerNorm kernel (folded into to_out's and fc1's epilogues)
    if (cfg->base_model == EDSNET_BASE_ATTENTION) return cfg->precision == EDSNET_PREC_FP32 ? 7 : 10;
    return cfg->precision == EDSNET_PREC_FP32 ? 11 : (cfg->precision == EDSNET_PREC_FP16 ? 12 : 11);
}

int edsnet_decode_boxes(const edsnet_config* cfg, const edsnet_batch* batch, const float* pred_loc,
                        float* boxes_f32, int32_t* boxes_i32, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!pred_loc || (!boxes_f32 && !boxes_i32)) return fail(EDSNET_E_ARG, "decode_boxes: NULL operand");
    const long long max_n = (long long)batch->max_rows * cfg->n_scales;
    int halo = 0;
    ScaleList sl = make_scales(cfg, &halo);
    decode_boxes_kernel<<<dim3((unsigned)((max_n + 255) / 256), batch->n_videos), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(pred_loc, batch->cu_rows, sl, boxes_f32, boxes_i32);
    CU_CHECK(cudaGetLastError(), "decode_boxes_kernel");
    return EDSNET_OK;
}

int edsnet_decode_nms(const edsnet_config* cfg, const edsnet_batch* batch, const float* pred_cls,
                      const float* pred_loc, double nms_thresh, float* boxes_f32, int32_t* boxes_i32,
                      int32_t* keep_count, int32_t* keep_idx, float* keep_scores, int32_t* keep_boxes,
                      const int64_t* nms_scratch_off, void* nms_scratch, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!pred_cls || !pred_loc || !boxes_i32 || !keep_count || !keep_idx || !keep_scores || !keep_boxes)
        return fail(EDSNET_E_ARG, "decode_nms: NULL operand");
    const long long max_n = (long long)batch->max_rows * cfg->n_scales;
    if (max_n > kNmsSmemCap && (!nms_scratch || (!nms_scratch_off && !batch->cu_rows_host)))
        return fail(EDSNET_E_ARG, "decode_nms: a video has more than 4096 anchors, scratch is required");
    const int skip_large = (max_n > kNmsSmemCap && batch->cu_rows_host) ? 1 : 0;
    if (max_n > (1ll << 30)) return fail(EDSNET_E_UNSUPPORTED, "decode_nms: video too long");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int halo = 0;
    ScaleList sl = make_scales(cfg, &halo);
    // shared memory per CTA: 24 bytes per anchor of the largest video, rounded up to a power of two (<= 4096 anchors)
    int cap = 32;
    while (cap < max_n && cap < kNmsSmemCap) cap <<= 1;
    const int nms_smem = cap * 24;
    // the largest size (4096 anchors x 24 B = 96 KB) once per device covers every launch
    static const char nms_tag = 0;
    if (DeviceOnce once_{&nms_tag}) CU_CHECK(opt_in_smem(nms_kernel, kNmsSmemCap * 24), "smem opt-in nms");
    {
        StageScope scope(ST_DECODE, st);
        decode_boxes_kernel<<<dim3((unsigned)((max_n + 255) / 256), batch->n_videos), 256, 0, st>>>(
            pred_loc, batch->cu_rows, sl, boxes_f32, boxes_i32);
        CU_CHECK(cudaGetLastError(), "decode_boxes_kernel");
    }
    {
        StageScope scope(ST_NMS, st);
        nms_kernel<<<batch->n_videos, kNmsThreads, nms_smem, st>>>(
            pred_cls, boxes_i32, batch->cu_rows, cfg->n_scales, nms_thresh, cap, skip_large,
            reinterpret_cast<const long long*>(nms_scratch_off), static_cast<unsigned char*>(nms_scratch), keep_count,
            keep_idx, keep_scores, keep_boxes);
        CU_CHECK(cudaGetLastError(), "nms_kernel");
        if (skip_large) {
            // videos with more than 4096 anchors: multi-CTA sort + blocked suppression, one video after the other;
            // scratch offsets follow the same rule as the caller's table (48 bytes per anchor, power of two)
            size_t off = 0;
            for (int v = 0; v < batch->n_videos; ++v) {
                const long long row0 = batch->cu_rows_host[v];
                const long long n = (long long)(batch->cu_rows_host[v + 1] - row0) * cfg->n_scales;
                if (n <= kNmsSmemCap) continue;
                size_t P = kNmsTile;
                while ((long long)P < n) P <<= 1;
                const size_t g0 = (size_t)row0 * cfg->n_scales;
                CU_CHECK(launch_nms_large(pred_cls + g0, boxes_i32 + 2 * g0, (int)n, (int)(n / cfg->n_scales), nms_thresh,
                                          static_cast<unsigned char*>(nms_scratch) + off, keep_count + v, keep_idx + g0,
                                          keep_scores + g0, keep_boxes + 2 * g0, st), "nms_large");
                off += (size_t)kNmsScratchPerAnchor * P;
            }
        }
    }
    return EDSNET_OK;
}

}  // extern "C"

// =====================================================================================================================
// Training step (BASELINE.json config 3; anchor_based/train.py:110-128): train-mode forward, losses, backward, Adam.
// =====================================================================================================================
namespace {

// column-maximum slots of all transposed splits of one backward: 2 x 128 + 128 + 1024 + 1024 + 1024 + 512 + 512 + 1536 + 1024
constexpr int kCmaxSlots = 8192;

size_t round64(size_t x) { return (x + 63) & ~(size_t)63; }
size_t planes_bytes(size_t rows, size_t cols) { return rows * cols * 4 + rows * 4; }

void train_layout(const edsnet_config* cfg, int32_t total_rows, int32_t n_videos, edsnet_train_layout* out) {
    edsnet_train_layout L;
    std::memset(&L, 0, sizeof(L));
    const size_t R = (size_t)(total_rows > 0 ? total_rows : 0), V = (size_t)(n_videos > 0 ? n_videos : 0);
    const size_t D = (size_t)(cfg && cfg->fc_depth > 0 ? cfg->fc_depth : 0), S = (size_t)(cfg ? cfg->n_scales : 1);
    const size_t Rp = round64(R), K5 = round64(D * R);
    const size_t head_mat = V * kHeads * 4096 * sizeof(float);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    // ---- forward: weight planes of the step, activations kept for the backward ----
    L.w_qkv16 = take(planes_bytes(kQkvCols, kFeat));
    L.w_out16 = take(planes_bytes(kFeat, kInner));
    L.w_fc116 = take(planes_bytes(kHidden, kFeat));
    L.w_fcb16 = take(planes_bytes(kHidden, kHidden));
    L.qkv16 = take(R * kQkvCols * 4);
    L.qkv_inv = take(R * 24 * sizeof(float));
    L.q_land = take(head_mat);
    L.k_land = take(head_mat);
    L.attn2 = take(head_mat);
    L.stats = take(V * kHeads * 2 * sizeof(float));
    L.a3v = take(head_mat);
    L.zmat = take(head_mat);
    L.wmat = take(head_mat);
    L.a3_part = take(V * kHeads * (size_t)a3v_split_cap((int)V) * kLandmark * tc::kA3PartLd * sizeof(float));
    L.merged = take(R * kInner * sizeof(float));
    L.y = take(R * kFeat * sizeof(float));
    L.yn = take(R * kFeat * sizeof(float));
    L.uin = take(std::max<size_t>(D, 1) * R * kHidden * sizeof(float));
    L.hs = take(std::max<size_t>(D, 1) * R * kHidden * sizeof(float));
    L.u_last = take(R * kHidden * sizeof(float));
    L.heads = take(R * 4 * sizeof(float));
    // ---- backward ----
    L.qkv_f32 = take(R * kQkvCols * sizeof(float));
    L.dqkv = take(R * kQkvCols * sizeof(float));
    L.m3 = take(V * kHeads * 64 * sizeof(float));
    L.l3 = take(V * kHeads * 64 * sizeof(float));
    L.acc0 = off;                                    // dW | dkl | dql | column maxima: zeroed by one memset at the start of the backward
    L.dw_att = take(head_mat);
    L.dkl = take(head_mat);
    L.dql = take(head_mat);
    L.cmax = take(kCmaxSlots * sizeof(unsigned));
    L.acc_bytes = off - L.acc0;
    L.db_att = take(head_mat);
    L.da2 = take(head_mat);
    L.dc_part = take(V * kHeads * sizeof(float));
    L.zhist = take(head_mat * (kPinvIters * 4 + 1));
    L.g = take(R * 4 * sizeof(float));
    L.d_logit = take(R * S * sizeof(float));
    L.das = take(std::max<size_t>(D, 1) * R * kHidden * sizeof(float));
    L.du0 = take(R * kHidden * sizeof(float));
    L.dyn = take(R * kFeat * sizeof(float));
    L.dy = take(R * kFeat * sizeof(float));
    L.dmerged = take(R * kInner * sizeof(float));
    // operand-plane scratch of the GEMMs (A side / B side), reused launch after launch
    L.t_a = take(std::max(std::max(planes_bytes(kQkvCols, Rp), planes_bytes(kHidden, K5)), planes_bytes(R, kFeat)));
    L.t_b = take(std::max(std::max(planes_bytes(kFeat, Rp), planes_bytes(kHidden, K5)),
                          std::max(planes_bytes(kFeat, kHidden), planes_bytes(kInner, kFeat))));
    // second pair for the weight-gradient products that run on the side stream next to the dX chain
    L.t_c = take(std::max(std::max(planes_bytes(kHidden, K5), planes_bytes(kHidden, Rp)), planes_bytes(kFeat, Rp)));
    L.t_d = take(std::max(std::max(planes_bytes(kHidden, K5), planes_bytes(kFeat, Rp)), planes_bytes(kInner, Rp)));
    L.total = off;
    *out = L;
}

// events that fork / join the backward's side stream: one set per device, created on first use
struct SideEvents { cudaEvent_t fork, fc, ln, hist, done; };
SideEvents* side_events() {
    static std::mutex m;
    static std::map<int, SideEvents> pool;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m);
    auto it = pool.find(dev);
    if (it == pool.end()) {
        SideEvents ev;
        cudaEvent_t* e[5] = {&ev.fork, &ev.fc, &ev.ln, &ev.hist, &ev.done};
        for (auto* p : e)
            if (cudaEventCreateWithFlags(p, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        it = pool.emplace(dev, ev).first;
    }
    return &it->second;
}

SplitTJob split_t_job(const float* src, int rows, int cols, int ld, void* planes, int* cta, unsigned** cmax) {
    SplitTJob j;
    j.src = src;
    j.cmax = *cmax;
    *cmax += cols;
    j.rows = rows;
    j.cols = cols;
    j.ld = ld;
    j.kp = (int)round64((size_t)rows);
    j.hi = static_cast<__half*>(planes);
    j.lo = j.hi + (size_t)cols * j.kp;
    j.inv = reinterpret_cast<float*>(j.lo + (size_t)cols * j.kp);
    j.cta0 = *cta;
    *cta += (cols + 31) / 32;
    return j;
}

int launch_split_t(const SplitTJob* jobs, int n, int ctas, cudaStream_t st) {
    StageScope scope(ST_T_SPLIT, st);
    SplitTJobs js;
    js.n = n;
    int max_rows = 1, max_kp = 64;
    for (int i = 0; i < 4; ++i) {
        js.j[i] = jobs[i < n ? i : 0];
        max_rows = std::max(max_rows, js.j[i].rows);
        max_kp = std::max(max_kp, js.j[i].kp);
    }
    split_t_colmax_kernel<<<dim3(ctas, (max_rows + 255) / 256), 256, 0, st>>>(js);
    CU_CHECK(cudaGetLastError(), "split_t_colmax_kernel");
    split_t_tiles_kernel<<<dim3(ctas, max_kp / 64), 256, 0, st>>>(js);
    CU_CHECK(cudaGetLastError(), "split_t_tiles_kernel");
    return EDSNET_OK;
}

int check_train_cfg(const edsnet_config* cfg) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (cfg->base_model != EDSNET_BASE_NYSTROM)
        return fail(EDSNET_E_UNSUPPORTED, "training kernels cover the Nystrom base (the hot path) only");
    if (cfg->fc_depth < 1) return fail(EDSNET_E_UNSUPPORTED, "training kernels need fc_depth >= 1");
    return EDSNET_OK;
}

}  // namespace

extern "C" {

size_t edsnet_train_workspace_bytes(const edsnet_config* cfg, int32_t total_rows, int32_t n_videos,
                                    edsnet_train_layout* layout) {
    edsnet_train_layout L;
    train_layout(cfg, total_rows, n_videos, &L);
    if (layout) *layout = L;
    return L.total;
}

int edsnet_train_launches(const edsnet_config* cfg, int32_t* forward, int32_t* backward) {
    if (!cfg) return fail(EDSNET_E_ARG, "config is NULL");
    if (forward) *forward = 4 + 1 + 1 + 5 + 1 + 1 + 1 + 1 + 1 + 1 + 1;       // weight planes, split, qkv, core (attn2 inside a3v), ..., roi
    if (backward) *backward = 37;
    return EDSNET_OK;
}

int edsnet_dropout_mask(uint64_t seed, uint64_t offset, int32_t rows, int32_t depth, uint8_t* out, void* stream) {
    if (!out || rows < 1 || depth < 1) return fail(EDSNET_E_ARG, "dropout_mask: bad argument");
    dropout_mask_kernel<<<(rows * depth + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, offset, rows, depth, out);
    CU_CHECK(cudaGetLastError(), "dropout_mask_kernel");
    return EDSNET_OK;
}

int edsnet_split_f16_t(const float* src, int64_t rows, int64_t cols, void* dst, void* stream) {
    if (!src || !dst || rows < 1 || cols < 1 || rows > (1 << 30) || cols > (1 << 24))
        return fail(EDSNET_E_ARG, "split_f16_t: bad argument");
    // the column maxima are gathered in the cols x 4 scratch bytes behind the planes
    const size_t kp = round64((size_t)rows);
    unsigned* cm = reinterpret_cast<unsigned*>(static_cast<unsigned char*>(dst) + planes_bytes((size_t)cols, kp));
    CU_CHECK(cudaMemsetAsync(cm, 0, (size_t)cols * sizeof(unsigned), static_cast<cudaStream_t>(stream)), "split_f16_t scratch");
    int cta = 0;
    SplitTJob j = split_t_job(src, (int)rows, (int)cols, (int)cols, dst, &cta, &cm);
    return launch_split_t(&j, 1, cta, static_cast<cudaStream_t>(stream));
}

int edsnet_train_forward(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch, const float* x,
                         int32_t dropout, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float* pred_cls,
                         float* pred_loc, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_train_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!w || !x || !pred_cls || !pred_loc || !workspace) return fail(EDSNET_E_ARG, "train_forward: NULL operand");
    edsnet_train_layout L;
    train_layout(cfg, batch->total_rows, batch->n_videos, &L);
    if (workspace_bytes < L.total) return fail(EDSNET_E_WORKSPACE, "train_forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const int R = batch->total_rows;
    const int P3 = EDSNET_PREC_FP16X3;              // every projection with three split-fp16 passes: fp32-grade
    // operand planes of this step's weights (they change every step)
    {
        StageScope scope(ST_T_WSPLIT, st);
        CU_CHECK(launch_split_f16(w->to_qkv_w, ws + L.w_qkv16, kQkvCols, kFeat, st), "split to_qkv.weight");
        CU_CHECK(launch_split_f16(w->to_out_w, ws + L.w_out16, kFeat, kInner, st), "split to_out.weight");
        CU_CHECK(launch_split_f16(w->fc1_w, ws + L.w_fc116, kHidden, kFeat, st), "split fc1.weight");
        CU_CHECK(launch_split_f16(w->fcb_w, ws + L.w_fcb16, kHidden, kHidden, st), "split fc_block.0.weight");
    }
    {
        StageScope scope(ST_SPLIT, st);
        CU_CHECK(launch_split_f16(x, ws + L.t_a, R, kFeat, st), "split x");
    }
    rc = gemm_dispatch(P3, EPI_QKV_PLANES, nullptr, ws + L.t_a, nullptr, ws + L.w_qkv16, F(L.qkv16), R, kQkvCols, kFeat,
                       nullptr, nullptr, kInner, st, ST_QKV, F(L.qkv_inv));
    if (rc) return rc;
    rc = nystrom_core_impl(P3, batch, F(L.qkv16), F(L.qkv_inv), w->res_conv_w, F(L.q_land), F(L.k_land), F(L.attn2),
                           F(L.stats), F(L.a3v), F(L.zmat), F(L.wmat), F(L.merged), st, nullptr, F(L.a3_part));
    if (rc) return rc;
    {
        StageScope scope(ST_SPLIT, st);
        CU_CHECK(launch_split_f16(F(L.merged), ws + L.t_a, R, kInner, st), "split merged");
    }
    rc = gemm_dispatch(P3, EPI_BIAS_RES, nullptr, ws + L.t_a, nullptr, ws + L.w_out16, F(L.y), R, kFeat, kInner,
                       w->to_out_b, x, 0, st, ST_TO_OUT);
    if (rc) return rc;
    {
        StageScope scope(ST_LN, st);
        layernorm1024_kernel<<<(R + 7) / 8, 256, 0, st>>>(F(L.y), w->ln_w, w->ln_b, F(L.yn), R);
        CU_CHECK(cudaGetLastError(), "layernorm1024_kernel");
    }
    {
        StageScope scope(ST_SPLIT, st);
        CU_CHECK(launch_split_f16(F(L.yn), ws + L.t_a, R, kFeat, st), "split LayerNorm output");
    }
    rc = gemm_dispatch(P3, EPI_BIAS, nullptr, ws + L.t_a, nullptr, ws + L.w_fc116, F(L.uin), R, kHidden, kFeat, w->fc1_b,
                       nullptr, 0, st, ST_FC1);
    if (rc) return rc;
    {
        StageScope scope(ST_FC_STACK, st);
        CU_CHECK(launch_fc_stack_tc(F(L.uin), ws + L.w_fcb16, w->fcb_b, w->fcb_ln_w, w->fcb_ln_b, F(L.u_last), R,
                                    cfg->fc_depth, st, w->cls_w, w->loc_w, F(L.heads), F(L.hs), dropout ? 1 : 0, seed,
                                    offset, reinterpret_cast<const unsigned long long*>(offset_dev)),
                 "fc_stack_tc_kernel<train>");
    }
    return roi_impl(cfg, w, batch, F(L.heads), pred_cls, pred_loc, st, true);
}

int edsnet_loss_grad(const edsnet_config* cfg, const edsnet_batch* batch, const float* pred_cls, const float* pred_loc,
                     const int32_t* cls_label, const float* loc_label, float lambda_reg, float scale, float* d_logit,
                     float* d_loc, float* loss_out, void* stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!pred_cls || !pred_loc || !cls_label || !loc_label || !d_logit || !d_loc || !loss_out)
        return fail(EDSNET_E_ARG, "loss_grad: NULL operand");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageScope scope(ST_T_LOSS, st);
    loss_grad_kernel<<<batch->n_videos, 256, 0, st>>>(pred_cls, pred_loc, cls_label, loc_label, batch->cu_rows,
                                                     cfg->n_scales, lambda_reg, scale, d_logit, d_loc, loss_out);
    CU_CHECK(cudaGetLastError(), "loss_grad_kernel");
    return EDSNET_OK;
}

int edsnet_train_backward(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch, const float* x,
                          const float* pred_cls, const float* d_cls, const float* d_loc, int32_t d_cls_is_logit_grad,
                          int32_t dropout, const edsnet_grads* grads, void* workspace, size_t workspace_bytes,
                          void* stream, void* side_stream) {
    int rc = check_train_cfg(cfg);
    if (rc) return rc;
    rc = check_batch(batch);
    if (rc) return rc;
    if (!w || !x || !pred_cls || !d_cls || !d_loc || !grads || !workspace)
        return fail(EDSNET_E_ARG, "train_backward: NULL operand");
    const float* const* gp = reinterpret_cast<const float* const*>(grads);
    for (size_t i = 0; i < sizeof(edsnet_grads) / sizeof(float*); ++i)
        if (!gp[i]) return fail(EDSNET_E_ARG, "train_backward: a gradient buffer is NULL");
    edsnet_train_layout L;
    train_layout(cfg, batch->total_rows, batch->n_videos, &L);
    if (workspace_bytes < L.total) return fail(EDSNET_E_WORKSPACE, "train_backward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const int R = batch->total_rows, V = batch->n_videos, D = cfg->fc_depth, S = cfg->n_scales;
    const int Rp = (int)round64((size_t)R), K5 = (int)round64((size_t)D * R);
    const int P3 = EDSNET_PREC_FP16X3;
    const int2* t64 = reinterpret_cast<const int2*>(batch->tiles64);
    const int2* t128 = reinterpret_cast<const int2*>(batch->tiles128);
    static const char bwd_tag = 0;
    if (DeviceOnce once_{&bwd_tag}) {
        CU_CHECK(opt_in_smem(fc_stack_bwd_kernel, kFcBwdSmem), "smem opt-in fc_stack_bwd");
        CU_CHECK(opt_in_smem(attn_bwd_rows_kernel, kAttnBwdRowsSmem), "smem opt-in attn_bwd_rows");
        CU_CHECK(opt_in_smem(pinv_bwd_kernel, kPinvBwdSmem), "smem opt-in pinv_bwd");
        CU_CHECK(opt_in_smem(pinv_hist_kernel, kPinvHistSmem), "smem opt-in pinv_hist");
        CU_CHECK(opt_in_smem(attn2_bwd_kernel, 4 * 64 * kLd64 * (int)sizeof(float)), "smem opt-in attn2_bwd");
        CU_CHECK(opt_in_smem(attn_bwd_keys_kernel, kAttnBwdKeysSmem), "smem opt-in attn_bwd_keys");
    }
    CU_CHECK(cudaMemsetAsync(ws + L.acc0, 0, L.acc_bytes, st), "zero attention accumulators");
    // Side stream (optional): what nothing downstream of the dX chain waits for -- the forward chain of the pseudo-inverse
    // (needed only by pinv_bwd) and the three weight-gradient products of the tail -- runs next to the chain.
    cudaStream_t ss = side_stream ? static_cast<cudaStream_t>(side_stream) : st;
    const bool par = ss != st;
    SideEvents* ev = par ? side_events() : nullptr;
    if (par && !ev) return fail(EDSNET_E_CUDA, "train_backward: could not create the side-stream events");
    if (par) {
        CU_CHECK(cudaEventRecord(ev->fork, st), "fork event");
        CU_CHECK(cudaStreamWaitEvent(ss, ev->fork, 0), "fork wait");
    }
    {
        StageScope scope(ST_T_PINV_BWD, ss);
        pinv_hist_kernel<<<dim3(kHeads, V), kPinvBwdThreads, kPinvHistSmem, ss>>>(F(L.attn2), F(L.stats), F(L.zhist), kPinvIters);
        CU_CHECK(cudaGetLastError(), "pinv_hist_kernel");
    }
    if (par) CU_CHECK(cudaEventRecord(ev->hist, ss), "hist event");
    // 1. d logits
    const float* d_logit = d_cls;
    if (!d_cls_is_logit_grad) {
        StageScope scope(ST_T_LOSS, st);
        sigmoid_bwd_kernel<<<(R * S + 255) / 256, 256, 0, st>>>(pred_cls, d_cls, F(L.d_logit), R * S);
        CU_CHECK(cudaGetLastError(), "sigmoid_bwd_kernel");
        d_logit = F(L.d_logit);
    }
    // 2. pooling + heads
    {
        int halo = 0;
        ScaleList sl = make_scales(cfg, &halo);
        StageScope scope(ST_T_ROI_BWD, st);
        roi_heads_bwd_kernel<<<batch->n_tiles128, 256, (size_t)S * 3 * (kRoiBwdRows + 1) * sizeof(float), st>>>(
            d_logit, d_loc, batch->cu_rows, t128, sl, halo, F(L.g), grads->cls_b, grads->loc_b);
        CU_CHECK(cudaGetLastError(), "roi_heads_bwd_kernel");
    }
    // 3. shared fc block x D
    {
        StageScope scope(ST_T_FC_BWD, st);
        FcBwdGrads fg{grads->fcb_b, grads->fcb_ln_w, grads->fcb_ln_b, grads->fc1_b, grads->cls_w, grads->loc_w};
        fc_stack_bwd_kernel<<<(R + 63) / 64, 256, kFcBwdSmem, st>>>(F(L.g), F(L.hs), F(L.uin), F(L.das), F(L.du0), w->fcb_w,
                                                                   w->fcb_ln_w, w->fcb_ln_b, w->cls_w, w->loc_w, R, D,
                                                                   dropout ? 2.f : 1.f, fg);
        CU_CHECK(cudaGetLastError(), "fc_stack_bwd_kernel");
    }
    SplitTJob jobs[4];
    int cta;
    unsigned* cm = reinterpret_cast<unsigned*>(ws + L.cmax);
    if (par) {
        CU_CHECK(cudaEventRecord(ev->fc, st), "fc event");
        CU_CHECK(cudaStreamWaitEvent(ss, ev->fc, 0), "fc wait");
    }
    // 4. d fc_block.0.weight = sum_l da_l^T u_l  (one product with K = D x rows)                      [side stream]
    cta = 0;
    jobs[0] = split_t_job(F(L.das), D * R, kHidden, kHidden, ws + L.t_c, &cta, &cm);
    jobs[1] = split_t_job(F(L.uin), D * R, kHidden, kHidden, ws + L.t_d, &cta, &cm);
    rc = launch_split_t(jobs, 2, cta, ss);
    if (rc) return rc;
    rc = gemm_dispatch(P3, EPI_NONE, nullptr, ws + L.t_c, nullptr, ws + L.t_d, grads->fcb_w, kHidden, kHidden, K5, nullptr,
                       nullptr, 0, ss, ST_T_GEMM_DW);
    if (rc) return rc;
    // 5. d fc1.weight = du0^T LN(y)                                                                    [side stream]
    cta = 0;
    jobs[0] = split_t_job(F(L.du0), R, kHidden, kHidden, ws + L.t_c, &cta, &cm);
    jobs[1] = split_t_job(F(L.yn), R, kFeat, kFeat, ws + L.t_d, &cta, &cm);
    rc = launch_split_t(jobs, 2, cta, ss);
    if (rc) return rc;
    rc = gemm_dispatch(P3, EPI_NONE, nullptr, ws + L.t_c, nullptr, ws + L.t_d, grads->fc1_w, kHidden, kFeat, Rp, nullptr,
                       nullptr, 0, ss, ST_T_GEMM_DW);
    if (rc) return rc;
    // 6. d LN(y) = du0 W1
    {
        StageScope scope(ST_T_SPLIT, st);
        CU_CHECK(launch_split_f16(F(L.du0), ws + L.t_a, R, kHidden, st), "split du0");
    }
    cta = 0;
    jobs[0] = split_t_job(w->fc1_w, kHidden, kFeat, kFeat, ws + L.t_b, &cta, &cm);
    rc = launch_split_t(jobs, 1, cta, st);
    if (rc) return rc;
    rc = gemm_dispatch(P3, EPI_NONE, nullptr, ws + L.t_a, nullptr, ws + L.t_b, F(L.dyn), R, kFeat, kHidden, nullptr, nullptr,
                       0, st, ST_T_GEMM_DX);
    if (rc) return rc;
    // 7. LayerNorm(1024)
    {
        StageScope scope(ST_T_LN_BWD, st);
        ln1024_bwd_kernel<<<(R + 63) / 64, 256, 0, st>>>(F(L.y), F(L.dyn), w->ln_w, F(L.dy), R, grads->ln_w, grads->ln_b,
                                                        grads->to_out_b);
        CU_CHECK(cudaGetLastError(), "ln1024_bwd_kernel");
    }
    if (par) {
        CU_CHECK(cudaEventRecord(ev->ln, st), "ln event");
        CU_CHECK(cudaStreamWaitEvent(ss, ev->ln, 0), "ln wait");
    }
    // 8. d to_out.weight = dy^T merged                                                                 [side stream]
    cta = 0;
    jobs[0] = split_t_job(F(L.dy), R, kFeat, kFeat, ws + L.t_c, &cta, &cm);
    jobs[1] = split_t_job(F(L.merged), R, kInner, kInner, ws + L.t_d, &cta, &cm);
    rc = launch_split_t(jobs, 2, cta, ss);
    if (rc) return rc;
    rc = gemm_dispatch(P3, EPI_NONE, nullptr, ws + L.t_c, nullptr, ws + L.t_d, grads->to_out_w, kFeat, kInner, Rp, nullptr,
                       nullptr, 0, ss, ST_T_GEMM_DW);
    if (rc) return rc;
    if (par) CU_CHECK(cudaEventRecord(ev->done, ss), "side done event");
    // 9. d merged = dy Wout
    {
        StageScope scope(ST_T_SPLIT, st);
        CU_CHECK(launch_split_f16(F(L.dy), ws + L.t_a, R, kFeat, st), "split dy");
    }
    cta = 0;
    jobs[0] = split_t_job(w->to_out_w, kFeat, kInner, kInner, ws + L.t_b, &cta, &cm);
    rc = launch_split_t(jobs, 1, cta, st);
    if (rc) return rc;
    rc = gemm_dispatch(P3, EPI_NONE, nullptr, ws + L.t_a, nullptr, ws + L.t_b, F(L.dmerged), R, kInner, kFeat, nullptr,
                       nullptr, 0, st, ST_T_GEMM_DX);
    if (rc) return rc;
    // 10. attention block
    {
        StageScope scope(ST_T_ATTN_PREP, st);
        const __half* p_hi = reinterpret_cast<const __half*>(ws + L.qkv16);
        qkv_planes_to_f32_kernel<<<(unsigned)(((size_t)R * (kQkvCols / 4) + 255) / 256), 256, 0, st>>>(
            p_hi, p_hi + (size_t)R * kQkvCols, F(L.qkv_inv), F(L.qkv_f32), R);
        CU_CHECK(cudaGetLastError(), "qkv_planes_to_f32_kernel");
        a3_stats_kernel<<<dim3(kHeads, V), 256, 0, st>>>(F(L.qkv_f32), batch->cu_rows, F(L.q_land), F(L.m3), F(L.l3));
        CU_CHECK(cudaGetLastError(), "a3_stats_kernel");
    }
    {
        StageScope scope(ST_T_ATTN_ROWS, st);
        attn_bwd_rows_kernel<<<dim3(batch->n_tiles64, kHeads), 256, kAttnBwdRowsSmem, st>>>(
            F(L.qkv_f32), F(L.dmerged), batch->cu_rows, t64, F(L.k_land), F(L.wmat), w->res_conv_w, F(L.dqkv), F(L.dw_att),
            F(L.dkl), grads->res_conv_w);
        CU_CHECK(cudaGetLastError(), "attn_bwd_rows_kernel");
    }
    if (par) CU_CHECK(cudaStreamWaitEvent(st, ev->hist, 0), "hist wait");
    {
        StageScope scope(ST_T_PINV_BWD, st);
        pinv_bwd_kernel<<<dim3(kHeads, V), kPinvBwdThreads, kPinvBwdSmem, st>>>(F(L.attn2), F(L.stats), F(L.a3v), F(L.dw_att), F(L.zhist),
                                                                   F(L.db_att), F(L.da2), F(L.dc_part), kPinvIters);
        CU_CHECK(cudaGetLastError(), "pinv_bwd_kernel");
    }
    {
        StageScope scope(ST_T_ATTN2_BWD, st);
        attn2_bwd_kernel<<<dim3(kHeads, V), 256, 4 * 64 * kLd64 * sizeof(float), st>>>(
            F(L.attn2), F(L.stats), F(L.da2), F(L.dc_part), F(L.q_land), F(L.k_land), F(L.dql), F(L.dkl));
        CU_CHECK(cudaGetLastError(), "attn2_bwd_kernel");
    }
    {
        StageScope scope(ST_T_ATTN_KEYS, st);
        attn_bwd_keys_kernel<<<dim3(batch->n_tiles64, kHeads), 256, kAttnBwdKeysSmem, st>>>(
            F(L.qkv_f32), batch->cu_rows, t64, F(L.q_land), F(L.a3v), F(L.db_att), F(L.m3), F(L.l3), F(L.dqkv), F(L.dql));
        CU_CHECK(cudaGetLastError(), "attn_bwd_keys_kernel");
    }
    {
        StageScope scope(ST_T_FINISH, st);
        dqkv_finish_kernel<<<dim3(batch->n_tiles64, kFinishSplit), 256, 0, st>>>(batch->cu_rows, t64, F(L.dql), F(L.dkl), F(L.dqkv));
        CU_CHECK(cudaGetLastError(), "dqkv_finish_kernel");
    }
    // 11. d to_qkv.weight = dqkv^T x
    cta = 0;
    jobs[0] = split_t_job(F(L.dqkv), R, kQkvCols, kQkvCols, ws + L.t_a, &cta, &cm);
    jobs[1] = split_t_job(x, R, kFeat, kFeat, ws + L.t_b, &cta, &cm);
    rc = launch_split_t(jobs, 2, cta, st);
    if (rc) return rc;
    rc = gemm_dispatch(P3, EPI_NONE, nullptr, ws + L.t_a, nullptr, ws + L.t_b, grads->to_qkv_w, kQkvCols, kFeat, Rp, nullptr,
                       nullptr, 0, st, ST_T_GEMM_DW);
    if (rc) return rc;
    if (par) CU_CHECK(cudaStreamWaitEvent(st, ev->done, 0), "side join");
    return EDSNET_OK;
}

int edsnet_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || n > (1ll << 31) - 256 || step < 1)
        return fail(EDSNET_E_ARG, "adam_step: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageScope scope(ST_T_ADAM, st);
    const double bc1 = 1.0 - std::pow((double)beta1, (double)step), bc2 = 1.0 - std::pow((double)beta2, (double)step);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (int)n, lr, beta1, beta2, eps,
                                                            weight_decay, (float)bc1, (float)std::sqrt(bc2), grad_scale);
    CU_CHECK(cudaGetLastError(), "adam_kernel");
    return EDSNET_OK;
}

}  // extern "C"
