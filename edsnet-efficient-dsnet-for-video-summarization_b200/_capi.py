"""ctypes binding of libedsnet_b200.so (the C ABI declared in include/edsnet_b200.h).

The library is the product: if it is missing or fails to load this module raises, it never
falls back to another implementation.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``python -m edsnet_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "csrc", "libedsnet_b200.so")

EDSNET_ABI_VERSION = 14
EDSNET_MAX_SCALES = 8

OK, E_ARG, E_CUDA, E_WORKSPACE, E_UNSUPPORTED = 0, 1, 2, 3, 4
PREC_FP32, PREC_FP16X3, PREC_FP16, PREC_FP16X2 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "fp16x3": PREC_FP16X3, "fp16": PREC_FP16, "fp16x2": PREC_FP16X2}
BASE_MODELS = {"nystromformer": 0, "attention": 1}


class Config(C.Structure):
    _fields_ = [("fc_depth", C.c_int32), ("n_scales", C.c_int32),
                ("scales", C.c_int32 * EDSNET_MAX_SCALES), ("precision", C.c_int32), ("base_model", C.c_int32)]


WEIGHT_FIELDS = ("to_qkv_w", "to_out_w", "to_out_b", "res_conv_w", "ln_w", "ln_b", "fc1_w", "fc1_b",
                 "fcb_w", "fcb_b", "fcb_ln_w", "fcb_ln_b", "cls_w", "cls_b", "loc_w", "loc_b",
                 "to_qkv_w16", "to_out_w16", "fc1_w16", "fcb_w16",
                 "mha_qkv_w", "mha_fc_w", "mha_qkv_w16", "mha_fc_w16",
                 "fc1_fold_w16", "fc1_fold_wgsum", "fc1_fold_b", "to_out_bc", "to_out_bounds")


class Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in WEIGHT_FIELDS]


class Batch(C.Structure):
    _fields_ = [("n_videos", C.c_int32), ("total_rows", C.c_int32), ("max_rows", C.c_int32),
                ("cu_rows", C.c_void_p), ("tiles64", C.c_void_p), ("n_tiles64", C.c_int32),
                ("tiles128", C.c_void_p), ("n_tiles128", C.c_int32), ("cu_rows_host", C.c_void_p)]


LAYOUT_FIELDS = ("qkv", "q_land", "k_land", "attn2", "stats", "qkv_inv", "a3v", "zmat", "wmat", "merged", "y", "yn",
                 "u0", "u1", "x16", "zeros", "mha16", "a3_part", "zstat", "xstat", "total")


class WorkspaceLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in LAYOUT_FIELDS]


GRAD_FIELDS = ("to_qkv_w", "to_out_w", "to_out_b", "res_conv_w", "ln_w", "ln_b", "fc1_w", "fc1_b",
               "fcb_w", "fcb_b", "fcb_ln_w", "fcb_ln_b", "cls_w", "cls_b", "loc_w", "loc_b")


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GRAD_FIELDS]


TRAIN_LAYOUT_FIELDS = ("w_qkv16", "w_out16", "w_fc116", "w_fcb16", "qkv16", "qkv_inv", "q_land", "k_land", "attn2", "stats",
                       "a3v", "zmat", "wmat", "a3_part", "merged", "y", "yn", "uin", "hs", "u_last", "heads", "qkv_f32", "dqkv", "m3",
                       "l3", "acc0", "acc_bytes", "dw_att", "dkl", "dql", "db_att", "da2", "cmax", "dc_part", "zhist", "g", "d_logit",
                       "das", "du0", "dyn", "dy", "dmerged", "t_a", "t_b", "t_c", "t_d", "total")


class TrainLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in TRAIN_LAYOUT_FIELDS]


class Shots(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cu_seg", "cps", "nfps", "picks", "cu_frames", "capacity", "gcd", "dp_off")]


class EvalTruth(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cu_users", "user_off", "user_frames", "user_summ", "metric")]


class CnnSrc(C.Structure):
    _fields_ = [("p", C.c_void_p), ("image_stride", C.c_int64), ("pixel_stride", C.c_int32),
                ("channel_stride", C.c_int32), ("col0", C.c_int32), ("channels", C.c_int32)]


class CnnInput(C.Structure):
    _fields_ = [("src", CnnSrc * 4), ("n_src", C.c_int32), ("relu", C.c_int32)]


# every symbol include/edsnet_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "edsnet_last_error": (C.c_char_p, []),
    "edsnet_abi_version": (C.c_int, []),
    "edsnet_workspace_bytes": (C.c_size_t, [C.POINTER(Config), C.c_int32, C.c_int32, C.POINTER(WorkspaceLayout)]),
    "edsnet_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(Batch), _P, _P, _P, _P,
                                 C.c_size_t, _P]),
    "edsnet_decode_nms": (C.c_int, [C.POINTER(Config), C.POINTER(Batch), _P, _P, C.c_double, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _P]),
    "edsnet_keyshot_summary": (C.c_int, [C.POINTER(Config), C.POINTER(Batch), C.POINTER(Shots), _P, _P, _P, _P, _P, _P,
                                         _P, _P, _P, _P]),
    "edsnet_eval_metrics": (C.c_int, [C.POINTER(Batch), _P, _P, C.POINTER(EvalTruth), _P, _P, _P, _P, _P, _P]),
    "edsnet_kts_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "edsnet_kts": (C.c_int, [C.POINTER(Batch), _P, _P, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                             _P, _P, _P, _P, _P]),
    "edsnet_cnn_im2col": (C.c_int, [C.POINTER(CnnInput), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, _P, _P]),
    "edsnet_cnn_maxpool": (C.c_int, [C.POINTER(CnnInput), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     _P, _P]),
    "edsnet_cnn_avgpool_l2norm": (C.c_int, [C.POINTER(CnnInput), C.c_int32, C.c_int32, _P, _P]),
    "edsnet_decode_boxes": (C.c_int, [C.POINTER(Config), C.POINTER(Batch), _P, _P, _P, _P]),
    "edsnet_forward_launches": (C.c_int, [C.POINTER(Config)]),
    "edsnet_split_f16_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "edsnet_split_f16": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P]),
    "edsnet_gemm": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P,
                              C.c_int32, _P]),
    "edsnet_nystrom_core": (C.c_int, [C.c_int32, C.POINTER(Batch), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "edsnet_fc_stack": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), _P, _P, C.c_int32, _P]),
    "edsnet_roi_pool_heads": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(Batch), _P, _P, _P, _P]),
    "edsnet_train_workspace_bytes": (C.c_size_t, [C.POINTER(Config), C.c_int32, C.c_int32, C.POINTER(TrainLayout)]),
    "edsnet_train_launches": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "edsnet_train_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(Batch), _P, C.c_int32, C.c_uint64,
                                       C.c_uint64, _P, _P, _P, _P, C.c_size_t, _P]),
    "edsnet_dropout_mask": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, _P, _P]),
    "edsnet_loss_grad": (C.c_int, [C.POINTER(Config), C.POINTER(Batch), _P, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P,
                                   _P]),
    "edsnet_train_backward": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(Batch), _P, _P, _P, _P, C.c_int32,
                                        C.c_int32, C.POINTER(Grads), _P, C.c_size_t, _P, _P]),
    "edsnet_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_int64, C.c_float, _P]),
    "edsnet_split_f16_t": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "edsnet_debug_tc_status": (C.c_int, [C.c_int32]),
    "edsnet_debug_stage_timing": (C.c_int, [C.c_int32]),
    "edsnet_debug_stage_count": (C.c_int, []),
    "edsnet_debug_stage_name": (C.c_char_p, [C.c_int32]),
    "edsnet_debug_stage_times": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32]),
    "edsnet_debug_set_tc_variant": (C.c_int, [C.c_int32]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)       # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if handle.edsnet_abi_version() != EDSNET_ABI_VERSION:
        raise RuntimeError("libedsnet_b200.so ABI version mismatch: rebuild the extension")
    _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().edsnet_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


class EdsnetError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"edsnet_b200 error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != OK:
        raise EdsnetError(rc, last_error())


def raise_on_tc_timeout() -> None:
    """Call at a point where the device work of interest has been synchronised (after .cpu(), stream.synchronize()).
    A tcgen05 pipeline wait that gave up (about one second without progress: preemption, a debugger, a hung peer) makes
    the kernel drain with undefined results instead of hanging; the flag it leaves is read AND cleared here, so one
    stall fails exactly the call it corrupted and does not poison later ones."""
    rc = lib().edsnet_debug_tc_status(1)
    if rc > 0:
        raise EdsnetError(E_CUDA, "a tcgen05 pipeline wait timed out; the results of this call are undefined "
                                  "(flag cleared, the call can be repeated)")
    if rc < 0:
        raise EdsnetError(E_CUDA, last_error())


def make_config(scales, fc_depth: int, precision: int, base_model: int = 0) -> Config:
    scales = [int(s) for s in scales]
    if not 1 <= len(scales) <= EDSNET_MAX_SCALES:
        raise ValueError(f"1..{EDSNET_MAX_SCALES} anchor scales supported, got {len(scales)}")
    cfg = Config()
    cfg.fc_depth = int(fc_depth)
    cfg.n_scales = len(scales)
    for i, s in enumerate(scales):
        cfg.scales[i] = s
    cfg.precision = int(precision)
    cfg.base_model = int(base_model)
    return cfg


def stage_times():
    """{stage name: (total ms, launches)} of the recording started by edsnet_debug_stage_timing(1)."""
    h = lib()
    n = h.edsnet_debug_stage_count()
    ms = (C.c_double * n)()
    cnt = (C.c_int32 * n)()
    check(h.edsnet_debug_stage_times(ms, cnt, n))
    return {h.edsnet_debug_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
