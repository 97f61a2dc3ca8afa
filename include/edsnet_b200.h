/*
 * edsnet_b200 -- C ABI of the B200 (sm_100a) implementation of EDSNet's anchor-based scoring path.
 *
 * The reference (ashish2506prasad/EDSNet-Efficient-DSNet-for-Video-Summarization) is pure Python/PyTorch and has
 * no FFI layer; its seam for this path is the Python class `DSNet` (src/anchor_based/dsnet.py:65-153).  The
 * entry points below are what a binding for that class needs; `edsnet_b200/dsnet.py` is that binding (ctypes) and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only.  Every pointer marked [dev] is a CUDA device pointer on the
 * current device; `stream` is a cudaStream_t passed as void*.  No entry point allocates, frees or synchronises;
 * all work is enqueued on `stream` and is re-entrant per stream as long as the workspaces differ.  Return value:
 * 0 = enqueued, non-zero = error (EDSNET_E_*), message via edsnet_last_error() (thread-local).
 * There is no CPU fallback: without a CUDA device every compute entry point returns EDSNET_E_CUDA.
 *
 * Packed variable-length batches: the feature rows of all videos are concatenated, x is [total_rows][1024] fp32,
 * cu_rows[v] .. cu_rows[v+1] are the rows of video v.  One reference call `model(x[None])` (dsnet.py:100-115) is
 * the special case n_videos == 1.
 */
#ifndef EDSNET_B200_H
#define EDSNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDSNET_ABI_VERSION 14

enum {
    EDSNET_OK = 0,
    EDSNET_E_ARG = 1,          /* bad argument (null pointer, odd anchor scale, too many scales, ...) */
    EDSNET_E_CUDA = 2,         /* CUDA runtime error, text in edsnet_last_error() */
    EDSNET_E_WORKSPACE = 3,    /* workspace too small */
    EDSNET_E_UNSUPPORTED = 4   /* configuration outside the accelerated path */
};

/* arithmetic of the dense projections (to_qkv, to_out, fc1) */
enum {
    EDSNET_PREC_FP32 = 0,      /* CUDA-core FFMA, fp32 operands                       (<= 1e-5 vs reference) */
    EDSNET_PREC_FP16X3 = 1,    /* tcgen05 fp16 hi/lo split, 3 MMA passes, fp32 accum  (<= 1e-5 vs reference) */
    EDSNET_PREC_FP16 = 2,      /* tcgen05 fp16 single pass, fp32 accumulate           (~ 1e-3 vs reference)  */
    EDSNET_PREC_FP16X2 = 3     /* tcgen05, to_qkv and to_out with TWO passes (activations rounded to fp16's 11 bits,
                                * weights split hi/lo: A_hi.[B_hi|B_lo] is one N = 256 instruction), fc1 / fc block /
                                * attention core as in FP16X3                          (<= 5e-4 vs reference) */
};

/* base model in front of the shared scoring tail (modules/models.py:118-147) */
enum {
    EDSNET_BASE_NYSTROM = 0,   /* NystromAttention, the accelerated hot path                                      */
    EDSNET_BASE_ATTENTION = 1  /* full multi-head attention (AttentionExtractor), comparison config 4 only        */
};

#define EDSNET_MAX_SCALES 8

/* Model hyper-parameters that are not baked in.  Baked in (reference call sites): num_feature 1024, num_hidden
 * 128, num_head 8, dim_head 64, 64 landmarks, 6 pinv iterations, 33-tap value conv (modules/models.py:134-135). */
typedef struct {
    int32_t fc_depth;                         /* applications of the shared fc block (dsnet.py:67,96) */
    int32_t n_scales;                         /* len(anchor_scales), 1..8                              */
    int32_t scales[EDSNET_MAX_SCALES];        /* even, 2..128 (odd scales crash the reference's .view) */
    int32_t precision;                        /* EDSNET_PREC_*                                         */
    int32_t base_model;                       /* EDSNET_BASE_*                                         */
} edsnet_config;

/* Weights in the reference's own state-dict layout (row-major (out, in) as nn.Linear stores them), fp32 [dev].
 * Names: DSNet.state_dict() keys, dsnet.py:66-98 / transformer/nystroformer.py:52-65. */
typedef struct {
    const float* to_qkv_w;     /* base_model.to_qkv.weight      (1536, 1024) */
    const float* to_out_w;     /* base_model.to_out.0.weight    (1024, 512)  */
    const float* to_out_b;     /* base_model.to_out.0.bias      (1024)       */
    const float* res_conv_w;   /* base_model.res_conv.weight    (8, 1, 33, 1)*/
    const float* ln_w;         /* layer_norm.weight             (1024)       */
    const float* ln_b;         /* layer_norm.bias               (1024)       */
    const float* fc1_w;        /* fc1.weight                    (128, 1024)  */
    const float* fc1_b;        /* fc1.bias                      (128)        */
    const float* fcb_w;        /* fc_block.0.weight             (128, 128)   */
    const float* fcb_b;        /* fc_block.0.bias               (128)        */
    const float* fcb_ln_w;     /* fc_block.3.weight             (128)        */
    const float* fcb_ln_b;     /* fc_block.3.bias               (128)        */
    const float* cls_w;        /* fc_cls.0.weight               (1, 128)     */
    const float* cls_b;        /* fc_cls.0.bias                 (1)          */
    const float* loc_w;        /* fc_loc.0.weight               (2, 128)     */
    const float* loc_b;        /* fc_loc.0.bias                 (2)          */
    /* fp16 hi/lo splits of the three projection weights, produced by edsnet_split_f16; required for the
     * tcgen05 precisions, ignored for EDSNET_PREC_FP32.  Layout: see edsnet_split_f16. */
    const void* to_qkv_w16;    /* planes of (1536, 1024) */
    const void* to_out_w16;    /* planes of (1024, 512)  */
    const void* fc1_w16;       /* planes of (128, 1024)  */
    const void* fcb_w16;       /* planes of (128, 128)   */
    /* EDSNET_BASE_ATTENTION only (modules/models.py:33-44, all bias-free): */
    const float* mha_qkv_w;    /* rows 0..1023 base_model.Q.weight, 1024..2047 K.weight, 2048..3071 V.weight (3072, 1024) */
    const float* mha_fc_w;     /* base_model.fc.0.weight        (1024, 1024) */
    const void* mha_qkv_w16;   /* planes of (3072, 1024) */
    const void* mha_fc_w16;    /* planes of (1024, 1024) */
    /* EDSNET_PREC_FP16X3, Nystrom base: LayerNorm(1024) is folded into fc1 (dsnet.py:105-106),
     *   fc1(LN(y)) = rstd (y (W o gamma)^T - mean(y) rowsum(W o gamma)) + (W beta + b),
     * so the forward needs these derived operands instead of fc1_w16 (required; edsnet_b200/dsnet.py builds them): */
    const void* fc1_fold_w16;      /* planes of fc1.weight * layer_norm.weight[None, :]   (128, 1024) */
    const float* fc1_fold_wgsum;   /* row sums of that product                            (128)       */
    const float* fc1_fold_b;       /* fc1.weight @ layer_norm.bias + fc1.bias             (128)       */
    const float* to_out_bc;        /* to_out.bias - mean(to_out.bias): LayerNorm ignores a row constant, and a
                                    * large common bias would otherwise cost the statistics digits      (1024)      */
    const float* to_out_bounds;    /* { max_n sum_k |to_out.weight[n,k]|, max_n |to_out_bc[n]| }, rounded up (2) */
} edsnet_weights;

/* A packed batch of videos.  Arrays [dev] int32 unless noted.  Tile tables are built by the host (see
 * edsnet_b200/plan.py): one {video, first_row} pair per 64-row (attention) / 128-row (pooling) tile of a video. */
typedef struct {
    int32_t n_videos;
    int32_t total_rows;        /* sum of T over the batch == cu_rows[n_videos] */
    int32_t max_rows;          /* max T in the batch */
    const int32_t* cu_rows;    /* [n_videos + 1] */
    const int32_t* tiles64;    /* [n_tiles64][2] */
    int32_t n_tiles64;
    const int32_t* tiles128;   /* [n_tiles128][2] */
    int32_t n_tiles128;
    const int32_t* cu_rows_host; /* optional HOST copy of cu_rows: lets edsnet_decode_nms spread videos with more than
                                  * 4096 anchors over the whole GPU (without it they run on one CTA each) */
} edsnet_batch;

/* Byte offsets of the intermediates inside the forward workspace (for tests / profiling). */
typedef struct {
    size_t qkv;        /* fp32 mode: [rows][1536] q (pre-scaled by 1/8) | k | v; tcgen05: fp16 hi plane, lo plane */
    size_t q_land;     /* [videos][8][64][64]                                         */
    size_t k_land;
    size_t attn2;      /* softmax(q_land k_land^T)                                    */
    size_t stats;      /* [videos][8][2]                                              */
    size_t qkv_inv;    /* tcgen05 precisions: [rows][24] inverse scales of the q|k|v planes */
    size_t a3v;        /* softmax(q_land k^T) v                                       */
    size_t zmat;       /* pseudo-inverse of attn2                                     */
    size_t wmat;       /* zmat a3v                                                    */
    size_t merged;     /* [rows][512] head-merged attention output + value conv ([rows][1024] for the attention base);
                        * tcgen05 precisions, Nystrom base: attention part only, the sum with the value conv leaves as
                        * the to_out operand planes at x16 */
    size_t y;          /* [rows][1024] to_out + bias + x; EDSNET_PREC_FP16X3, Nystrom base: fc1 operand planes of
                        * z = y - mean(x row) (edsnet_split_f16 layout), LayerNorm being folded into fc1 */
    size_t yn;         /* [rows][1024] LayerNorm(y)           (aliases qkv)           */
    size_t u0;         /* [rows][128] fc1 output                                      */
    size_t u1;         /* [rows][128] after the fc stack; tcgen05 precisions: [rows][4] head projections
                        * (u . w_cls, u . w_loc[0], u . w_loc[1], 0), the hidden rows are not written */
    size_t x16;        /* tcgen05 precisions: operand planes of the current GEMM's A   */
    size_t zeros;      /* [1024] zero bias (attention base: its projections have no bias) */
    size_t mha16;      /* attention base, tcgen05 precisions: Q | K | V operand planes [rows][3072] hi, lo (scales in qkv_inv) */
    size_t a3_part;    /* tcgen05 precisions, few videos: per (video, head, key range) un-normalised rows of
                        * softmax(q_land k^T) v with their running max / sum, [..][64][66], merged into a3v */
    size_t zstat;      /* EDSNET_PREC_FP16X3, Nystrom base: [rows][16][2] (sum z, sum z^2) per 64-column slot */
    size_t xstat;      /* EDSNET_PREC_FP16X3, Nystrom base: [rows][2] (mean, max|.|) of the input rows */
    size_t total;
} edsnet_workspace_layout;

const char* edsnet_last_error(void);
int edsnet_abi_version(void);

/* Workspace sizing: fills *layout (may be NULL) and returns the bytes edsnet_forward needs. */
size_t edsnet_workspace_bytes(const edsnet_config* cfg, int32_t total_rows, int32_t n_videos,
                              edsnet_workspace_layout* layout);

/* DSNet.forward (dsnet.py:100-115) over a packed batch: x [dev][total_rows][1024] fp32 ->
 * pred_cls [dev][total_rows][S] (sigmoid scores), pred_loc [dev][total_rows][S][2] (centre, log-width offsets).
 * x and workspace must be 32-byte aligned (EDSNET_E_ARG otherwise). */
int edsnet_forward(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch,
                   const float* x, float* pred_cls, float* pred_loc,
                   void* workspace, size_t workspace_bytes, void* stream);

/* DSNet.predict's decode (dsnet.py:146-153: get_anchors + offset2bbox + cw2lr), evaluate.py:26 clip/round and
 * bbox_helper.nms (helpers/bbox_helper.py:80-118), per video.
 *   boxes_f32 [dev][total_rows*S][2] float32 left/right (may be NULL), boxes_i32 [dev][total_rows*S][2],
 *   keep_count [dev][n_videos]; keep_idx / keep_scores / keep_boxes are [dev] arrays of total_rows*S entries,
 *   the kept proposals of video v (descending score) start at entry cu_rows[v]*S; keep_idx is the flat
 *   anchor index t*S+s inside the video.
 *   Videos with more than 4096 anchors need scratch: 48 bytes per anchor, anchors rounded up to a power of two,
 *   video after video in batch order; nms_scratch_off [dev][n_videos] are those byte offsets (only read when
 *   batch->cu_rows_host is NULL).  Both may be NULL when no video is that long. */
int edsnet_decode_nms(const edsnet_config* cfg, const edsnet_batch* batch, const float* pred_cls,
                      const float* pred_loc, double nms_thresh, float* boxes_f32, int32_t* boxes_i32,
                      int32_t* keep_count, int32_t* keep_idx, float* keep_scores, int32_t* keep_boxes,
                      const int64_t* nms_scratch_off, void* nms_scratch, void* stream);

/* Shot structure of the videos of a batch, for edsnet_keyshot_summary.  All arrays [dev].
 * dp_scratch layout per video (byte offset dp_off[v]): two int32 rows of (capacity[v] + 1), then
 * n_seg x ceil((capacity[v] + 1) / 32) uint32 decision words. */
typedef struct {
    const int32_t* cu_seg;      /* [n_videos + 1] first shot of every video                                   */
    const int32_t* cps;         /* [total_seg][2] change points: first / last frame of the shot, inclusive     */
    const int32_t* nfps;        /* [total_seg] frames per shot (the knapsack weights)                          */
    const int32_t* picks;       /* [total_rows] frame index of every sub-sampled position, aligned with cu_rows */
    const int64_t* cu_frames;   /* [n_videos + 1] prefix sum of n_frames                                       */
    const int32_t* capacity;    /* [n_videos] int(n_frames * 0.15) / gcd[v]                                    */
    const int32_t* gcd;         /* [n_videos] a common divisor of capacity and all shot weights (1 always works) */
    const int64_t* dp_off;      /* [n_videos] byte offsets into dp_scratch                                     */
} edsnet_shots;

/* evaluate.py:29 / infer.py:35: vsumm_helper.bbox2summary (helpers/vsumm_helper.py:101-116, 53-98, 26-45) on the
 * device, from the kept proposals edsnet_decode_nms wrote.  With keep_count == NULL the per-position scores are taken
 * from pos_scores as an INPUT instead (vsumm_helper.get_keyshot_summ on given scores: the ground-truth targets of
 * anchor_based/train.py:79) and keep_scores / keep_boxes are not read.  Outputs: pos_scores [total_rows], frame_scores
 * [total_frames], seg_scores [total_seg], picked [total_seg] (0/1), summary [total_frames] (0/1). */
int edsnet_keyshot_summary(const edsnet_config* cfg, const edsnet_batch* batch, const edsnet_shots* shots,
                           const int32_t* keep_count, const float* keep_scores, const int32_t* keep_boxes,
                           float* pos_scores, float* frame_scores, int32_t* seg_scores, uint8_t* picked,
                           uint8_t* summary, void* dp_scratch, void* stream);

/* Ground-truth user summaries of the videos of a batch, for edsnet_eval_metrics.  All arrays [dev]. */
typedef struct {
    const int32_t* cu_users;    /* [n_videos + 1] first user row of every video                                  */
    const int64_t* user_off;    /* [total_users] byte offset of the user's 0/1 row inside user_summ, multiple of 4 */
    const int32_t* user_frames; /* [n_videos] frames per user row (user_summary.shape[1])                         */
    const uint8_t* user_summ;   /* 0/1 bytes, rows padded to a multiple of 4 bytes                                */
    const int32_t* metric;      /* [n_videos] 0 = 'avg' (TVSum keys), 1 = 'max' (evaluate.py:31)                  */
} edsnet_eval_truth;

/* evaluate.py:31-37: vsumm_helper.get_summ_f1score (helpers/vsumm_helper.py:142-172, f1_score :8-23) and
 * get_summ_diversity (:119-139) of downsample_summ (:48-50), per video, on the device.
 *   summary [total_frames] 0/1 and cu_frames [n_videos + 1] as written by / given to edsnet_keyshot_summary;
 *   x [total_rows][1024] the batch's features.  Outputs: fscore [n_videos] float64 (bit-exact: integer counts +
 *   the reference's float64 operations), diversity [n_videos] float64 (float64 accumulation; the reference sums
 *   float32 products: equal to ~1e-6 relative), user_f1 [total_users] float64, counts [n_videos][2] =
 *   {frames selected within the users' frame count, selected feature rows}. */
int edsnet_eval_metrics(const edsnet_batch* batch, const int64_t* cu_frames, const uint8_t* summary,
                        const edsnet_eval_truth* truth, const float* x, double* fscore, double* diversity,
                        double* user_f1, int32_t* counts, void* stream);

/* One video of an edsnet_kts call: its block inside the scratch buffer and its rows in the packed features. */
typedef struct {
    int64_t scratch_off;       /* byte offset, multiple of 256; the block holds edsnet_kts_scratch_bytes(n) bytes */
    int32_t row0;              /* first packed feature row (== cu_rows[v])                                        */
    int32_t n;                 /* frames (== T of the video)                                                      */
} edsnet_kts_video;

/* Kernel temporal segmentation (kts/cpd_auto.py:6-33, kts/cpd_nonlin.py:4-92), the shot boundaries the reference
 * computes from the sub-sampled features (helpers/video_helper.py:109-126) before it scores a video.
 *   videos [dev][n_videos]; x [dev][total_rows][1024] or NULL: with NULL every video's float32 kernel matrix
 *   (n x n, row-major) must already sit at the start of its scratch block (np.matmul(features, features.T) of the
 *   caller); ncp_cap < 0: at most n - 1 change points (video_helper.py:118); m_fixed >= 0: exactly that many
 *   (cpd_nonlin) instead of cpd_auto's penalised choice; vmax / desc_rate / lmin / lmax as cpd_auto / cpd_nonlin.
 *   Outputs: n_cps [dev][n_videos]; cps [dev][total_rows] (video v's change points, ascending, from entry cu_rows[v]);
 *   objective [dev][total_rows] float64 (may be NULL): the objective for 0 .. n_cps[v] change points.
 * For a given kernel matrix the change points are the reference's bit for bit (same float32 / float64 operations in
 * the same order, first minimum); K = X X^T itself is computed in float32 in a fixed order, which the reference's
 * BLAS does not promise.  n is limited to 12 800 frames (two DP rows in shared memory). */
size_t edsnet_kts_scratch_bytes(int32_t n);
int edsnet_kts(const edsnet_batch* batch, const edsnet_kts_video* videos, const float* x, int32_t ncp_cap,
               int32_t m_fixed, double vmax, int32_t desc_rate, int32_t lmin, int32_t lmax, int32_t* n_cps,
               int32_t* cps, double* objective, void* scratch, void* stream);

/* ---- GoogLeNet pool5 feature extraction (helpers/video_helper.py:27-73), the step in front of the scoring path ----
 * The convolutions are products on the tcgen05 GEMM (edsnet_gemm, epilogue 2 = + bias, with BatchNorm folded into
 * weights and bias by the host); the entry points below are what surrounds them.  Activations are fp32
 * [pixels][ld] buffers (NHWC; ld >= channels); an input is a VIRTUAL channel concatenation of up to four such
 * buffers (the four branches of an inception module are never concatenated in memory) with the producing layers'
 * ReLU applied while reading.  General strides, so the network input may be NCHW. */
typedef struct {
    const float* p;            /* [dev] first element                                                       */
    int64_t image_stride;      /* elements between images                                                   */
    int32_t pixel_stride;      /* elements between pixels (row-major y * W + x)                             */
    int32_t channel_stride;    /* elements between channels                                                 */
    int32_t col0;              /* first channel column taken from this buffer                               */
    int32_t channels;          /* channels taken                                                            */
} edsnet_cnn_src;
typedef struct {
    edsnet_cnn_src src[4];
    int32_t n_src;             /* 1..4                                                                      */
    int32_t relu;              /* 1: max(x, 0) applied to every value read                                  */
} edsnet_cnn_input;

/* Patch gather of a kh x kw convolution (stride, zero padding `pad`) into the GEMM operand format: row m = output
 * pixel (image, oy, ox), column k = (ky * kw + kx) * C + c with C = the input's total channels, zero padded to kpad (a
 * multiple of 64).  planes: hi [M][kpad] fp16 | lo [M][kpad] fp16 | inverse row scales [M] fp32 =
 * edsnet_split_f16_bytes(M, kpad) bytes, M = n_img * OH * OW, OH = (H + 2 pad - kh) / stride + 1. */
int edsnet_cnn_im2col(const edsnet_cnn_input* in, int32_t n_img, int32_t H, int32_t W, int32_t kh, int32_t kw,
                      int32_t stride, int32_t pad, int32_t kpad, void* planes, void* stream);
/* MaxPool2d(k, stride, pad, ceil_mode=True) -> out [dev][n_img * OH * OW][C] fp32; OH as torch computes it. */
int edsnet_cnn_maxpool(const edsnet_cnn_input* in, int32_t n_img, int32_t H, int32_t W, int32_t k, int32_t stride,
                       int32_t pad, float* out, void* stream);
/* AdaptiveAvgPool2d(1) over HW pixels, then feat / (|feat|_2 + 1e-10) -> out [dev][n_img][C] fp32, C <= 1024. */
int edsnet_cnn_avgpool_l2norm(const edsnet_cnn_input* in, int32_t n_img, int32_t HW, float* out, void* stream);

/* decode only (what DSNet.predict returns, dsnet.py:146-153, plus the evaluate.py:26 clip/round):
 * boxes_f32 [dev][total_rows*S][2] (may be NULL), boxes_i32 [dev][total_rows*S][2] (may be NULL). */
int edsnet_decode_boxes(const edsnet_config* cfg, const edsnet_batch* batch, const float* pred_loc,
                        float* boxes_f32, int32_t* boxes_i32, void* stream);

/* Number of kernel launches one edsnet_forward call enqueues for this configuration (bench bookkeeping). */
int edsnet_forward_launches(const edsnet_config* cfg);

/* fp32 (rows, cols), cols in {128, 512, 1024} -> operand planes for the tcgen05 precisions.  Every row is first scaled
 * by a power of two that puts its largest magnitude in [2^14, 2^15) (keeps hi AND lo in fp16's normal range).
 * dst layout, edsnet_split_f16_bytes(rows, cols) bytes: hi plane [rows][cols] fp16 (hi = fp16(x 2^s)) |
 * lo plane [rows][cols] fp16 (lo = fp16(x 2^s - hi)) | inverse scales [rows] fp32 (2^-s). */
size_t edsnet_split_f16_bytes(int64_t rows, int64_t cols);
int edsnet_split_f16(const float* src, void* dst_hi_lo, int64_t rows, int64_t cols, void* stream);

/* ---- training step (BASELINE.json config 3) ------------------------------------------------------------------------
 * What the reference's loop body runs per video (anchor_based/train.py:110-128): `model(seq)` in train() mode,
 * calc_cls_loss + calc_loc_loss (anchor_based/losses.py:5-57), `loss.backward()`, `optimizer.step()` (Adam, :53-55) -- as
 * CUDA kernels on a packed batch of videos.  Nystrom base only, fc_depth >= 1.  All projections and their dW / dX
 * products run on tcgen05 with three split-fp16 passes; the 64-wide attention backward is fp32 on CUDA cores.
 * The input features get no gradient (the reference's features are data). */

/* Gradient buffers, fp32 [dev], one per parameter, same shapes as the weights of the same name.  They must be ZERO on
 * entry to edsnet_train_backward (the small ones are accumulated atomically); on exit they hold d loss / d parameter. */
typedef struct {
    float* to_qkv_w;   float* to_out_w;  float* to_out_b;  float* res_conv_w;
    float* ln_w;       float* ln_b;      float* fc1_w;     float* fc1_b;
    float* fcb_w;      float* fcb_b;     float* fcb_ln_w;  float* fcb_ln_b;
    float* cls_w;      float* cls_b;     float* loc_w;     float* loc_b;
} edsnet_grads;

/* Byte offsets inside the training workspace (tests / profiling).  The forward fills the first group and the backward
 * reads it, so ONE workspace must be passed to both calls of a step. */
typedef struct {
    size_t w_qkv16, w_out16, w_fc116, w_fcb16;   /* operand planes of the step's weights (edsnet_split_f16 layout)       */
    size_t qkv16, qkv_inv;                       /* q | k | v operand planes [rows][1536] hi, lo; scales [rows][24]      */
    size_t q_land, k_land, attn2, stats, a3v, zmat, wmat;   /* as in edsnet_workspace_layout                             */
    size_t a3_part;                              /* key-range partials of a3v (few, long videos), as in the forward     */
    size_t merged;                               /* [rows][512] fp32: attention + value convolution, head-merged         */
    size_t y;                                    /* [rows][1024] to_out + bias + x                                      */
    size_t yn;                                   /* [rows][1024] LayerNorm(y)                                           */
    size_t uin;                                  /* [depth][rows][128] input of every application of the fc block       */
    size_t hs;                                   /* [depth][rows][128] Dropout(ReLU(Linear)) of every application        */
    size_t u_last;                               /* [rows][128] final hidden rows                                        */
    size_t heads;                                /* [rows][4] head projections                                          */
    size_t qkv_f32;                              /* [rows][1536] fp32 copy of q/8 | k | v (backward)                     */
    size_t dqkv;                                 /* [rows][1536] gradient of the to_qkv output                           */
    size_t m3, l3;                               /* [videos][8][64] row maximum / normaliser of softmax(q_land k^T)      */
    size_t acc0, acc_bytes;                      /* dw_att | dkl | dql | cmax, zeroed at the start of the backward      */
    size_t dw_att, dkl, dql, db_att, da2;        /* [videos][8][64][64] each                                            */
    size_t cmax;                                 /* column maxima of the transposed operand splits (uint bit patterns)   */
    size_t dc_part;                              /* [videos][8] gradient of the pseudo-inverse start scale, per head     */
    size_t zhist;                                /* [videos][8][6 * 4 + 1][64][64] Z_k, A Z_k, T2_k, T3_k of the pinv chain, Z */
    size_t g;                                    /* [rows][4] gradient of the head projections                           */
    size_t d_logit;                              /* [rows][S]                                                            */
    size_t das;                                  /* [depth][rows][128] gradient of every Linear output of the fc block   */
    size_t du0, dyn, dy, dmerged;                /* [rows][128], [rows][1024], [rows][1024], [rows][512]                 */
    size_t t_a, t_b;                             /* operand-plane scratch of the backward GEMMs                          */
    size_t t_c, t_d;                             /* the same for the weight-gradient products on the side stream         */
    size_t total;
} edsnet_train_layout;

size_t edsnet_train_workspace_bytes(const edsnet_config* cfg, int32_t total_rows, int32_t n_videos,
                                    edsnet_train_layout* layout);
/* Kernel launches of one forward / one backward call (bench bookkeeping). */
int edsnet_train_launches(const edsnet_config* cfg, int32_t* forward, int32_t* backward);

/* DSNet.forward in train() mode (anchor_based/dsnet.py:100-115 with the Dropout(0.5) of the shared fc block, :91-95,
 * active when dropout != 0) keeping every activation the backward needs.  The mask is Philox4x32-10 with counter
 * (row, layer, offset) and key seed: bit c of the 128 output bits keeps hidden column c; edsnet_dropout_mask writes it
 * out ([depth][rows][128] bytes, 1 = kept) for tests.  offset_dev (may be NULL): one uint64 in DEVICE memory that is added
 * to `offset` when the kernel runs, so that a captured CUDA graph draws a fresh mask on every replay.  Only the fp32 weight
 * pointers of `w` are read. */
int edsnet_train_forward(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch, const float* x,
                         int32_t dropout, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float* pred_cls,
                         float* pred_loc, void* workspace, size_t workspace_bytes, void* stream);
int edsnet_dropout_mask(uint64_t seed, uint64_t offset, int32_t rows, int32_t depth, uint8_t* out, void* stream);

/* anchor_based/losses.py:5-57 combined as anchor_based/train.py:119-123, per video of the batch:
 *   cls_label [dev][total_rows][S] int32 (1 positive, -1 negative, 0 ignored), loc_label [dev][total_rows][S][2] fp32.
 * Outputs: loss_out [dev][n_videos][3] = {cls + lambda_reg * loc, cls, loc}; d_logit [total_rows][S] and d_loc
 * [total_rows][S][2] = scale * d loss / d (logit before the sigmoid | offset), in closed form (no division by p or
 * 1 - p).  scale is typically 1 / (videos of the optimiser step). */
int edsnet_loss_grad(const edsnet_config* cfg, const edsnet_batch* batch, const float* pred_cls, const float* pred_loc,
                     const int32_t* cls_label, const float* loc_label, float lambda_reg, float scale, float* d_logit,
                     float* d_loc, float* loss_out, void* stream);

/* loss.backward(): parameter gradients from d loss / d outputs.  d_cls is either the gradient with respect to the
 * logits (d_cls_is_logit_grad != 0, what edsnet_loss_grad writes) or with respect to pred_cls after the sigmoid (what
 * torch.autograd hands over); pred_cls = the forward's output; dropout = the forward's flag; workspace = the forward's.
 * side_stream (may be NULL): a second stream of the same device on which the work nothing else waits for -- the fp32
 * re-run of the pseudo-inverse chain and the three weight-gradient products of the tail -- runs next to the dX chain;
 * it is forked from and joined back into `stream` with events inside the call (capturable into a CUDA graph), so the
 * caller never synchronises it.  The events are per device: one call at a time per device when a side stream is given. */
int edsnet_train_backward(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch, const float* x,
                          const float* pred_cls, const float* d_cls, const float* d_loc, int32_t d_cls_is_logit_grad,
                          int32_t dropout, const edsnet_grads* grads, void* workspace, size_t workspace_bytes,
                          void* stream, void* side_stream);

/* torch.optim.Adam(lr, betas, eps, weight_decay) (anchor_based/train.py:53-55) on flat fp32 buffers of n values; the
 * gradient is multiplied by grad_scale first (1 / world size after a summing all-reduce); step counts from 1. */
int edsnet_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream);

/* fp32 (rows, cols) -> operand planes of the TRANSPOSE: hi [cols][kp] fp16 | lo [cols][kp] fp16 | inverse scales [cols]
 * fp32, kp = rows rounded up to 64 (zero padded), every output row scaled like edsnet_split_f16 does.  The operand format
 * of every dW = dY^T X product of the backward.  dst needs edsnet_split_f16_bytes(cols, kp) + 4 * cols bytes (the last
 * 4 * cols are scratch for the column maxima). */
int edsnet_split_f16_t(const float* src, int64_t rows, int64_t cols, void* dst, void* stream);

/* ---- stage-level entry points (tests, per-kernel timing, ncu) ---- */

/* C[M][N] = A[M][K] . B[N][K]^T, epilogue: 0 none, 1 first qcols columns * 1/8, 2 + bias, 3 + bias + res
 * (4, the q|k|v plane epilogue, is internal to edsnet_forward). */
int edsnet_gemm(int32_t precision, int32_t epilogue, const float* A, const void* A16, const float* B,
                const void* B16, float* C, int32_t M, int32_t N, int32_t K, const float* bias,
                const float* res, int32_t qcols, void* stream);
/* qkv -> merged: landmarks, three softmax kernels, pseudo-inverse, aggregation, value conv (nystroformer.py).
 * precision selects the CUDA-core (EDSNET_PREC_FP32) or the tcgen05 split-fp16 kernels for the row stages. */
int edsnet_nystrom_core(int32_t precision, const edsnet_batch* batch, const float* qkv, const float* qkv_inv,
                        const float* res_conv_w, float* q_land,
                        float* k_land, float* attn2, float* stats, float* a3v, float* zmat, float* wmat,
                        float* merged, void* stream);
/* u0 -> u1: fc_depth applications of the shared block. */
int edsnet_fc_stack(const edsnet_config* cfg, const edsnet_weights* w, const float* u_in, float* u_out,
                    int32_t rows, void* stream);
/* u1 -> pred_cls / pred_loc. */
int edsnet_roi_pool_heads(const edsnet_config* cfg, const edsnet_weights* w, const edsnet_batch* batch,
                          const float* u, float* pred_cls, float* pred_loc, void* stream);

/* ---- diagnostics ---- */

/* Synchronous.  Returns 1 if a tcgen05 GEMM pipeline wait timed out since the last reset (the kernel then drained
 * with undefined results instead of hanging), 0 if not, -1 on CUDA error.  reset != 0 clears the flag. */
int edsnet_debug_tc_status(int32_t reset);
/* Per-stage CUDA-event timing of every kernel the entry points launch (bench / profiling; off by default, adds two
 * event records per launch, not thread safe).  enable != 0 starts a fresh recording, 0 stops and discards.
 * edsnet_debug_stage_times synchronises on the recorded events, fills ms_sum[stage] / launches[stage] for
 * stage < edsnet_debug_stage_count() and clears the recording. */
int edsnet_debug_stage_timing(int32_t enable);
int edsnet_debug_stage_count(void);
const char* edsnet_debug_stage_name(int32_t stage);
int edsnet_debug_stage_times(double* ms_sum, int32_t* launches, int32_t n);
/* Variant of the tcgen05 GEMM (profiling): 0 = default, 3 = two double-buffered accumulators for every K,
 * 4 = CTA-pair kernel (tcgen05.mma.cta_group::2). */
int edsnet_debug_set_tc_variant(int32_t variant);

#ifdef __cplusplus
}
#endif
#endif
