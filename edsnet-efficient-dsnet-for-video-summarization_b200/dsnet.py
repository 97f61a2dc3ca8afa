"""Drop-in replacement for the reference's anchor-based scoring model.

Mirrors `DSNet` of the reference (src/anchor_based/dsnet.py:65-153): same constructor signature, same
parameter names / shapes / aliasing (so reference checkpoints load with strict=True and `model.apply(xavier_init)`
acts identically, src/anchor_based/train.py:19-24,51), same `forward(x) -> (pred_cls, pred_loc)` and
`predict(seq) -> (scores, left/right boxes)`.  The arithmetic is NOT torch: every call goes through the C ABI of
libedsnet_b200.so (include/edsnet_b200.h) to hand-written sm_100a kernels.  There is no CPU path: a CPU tensor,
a missing extension or an unsupported configuration raises.

Beyond the reference surface (whose loop is one video per call, evaluate.py:19-28) the class adds packed
multi-video entry points -- `forward_packed`, `proposals_packed` -- used by the throughput pipeline.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from . import _capi
from .plan import BatchPlan, DeviceBatch

NUM_FEATURE = 1024
NUM_HIDDEN = 128
NUM_HEAD = 8
DIM_HEAD = 64


class NystromAttention(nn.Module):
    """Parameter container with the layout of the reference's NystromAttention
    (src/transformer/nystroformer.py:32-65, built at src/modules/models.py:134-135).  It has no forward of
    its own: DSNet's fused kernels consume these parameters directly."""

    def __init__(self, dim: int, dim_head: int = 64, heads: int = 8, num_landmarks: int = 64,
                 pinv_iterations: int = 6, residual: bool = True, residual_conv_kernel: int = 33,
                 eps: float = 1e-8, dropout: float = 0.0):
        super().__init__()
        inner = heads * dim_head
        self.eps = eps
        self.num_landmarks = num_landmarks
        self.pinv_iterations = pinv_iterations
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.residual = residual
        if residual:
            k = residual_conv_kernel
            self.res_conv = nn.Conv2d(heads, heads, (k, 1), padding=(k // 2, 0), groups=heads, bias=False)

    def forward(self, x, mask=None, return_attn=False):  # pragma: no cover - guarded surface
        raise RuntimeError("edsnet_b200.NystromAttention only holds parameters; call DSNet.forward "
                           "(the attention block is fused into the scoring kernels)")


class AttentionExtractor(nn.Module):
    """Parameter container with the layout of the reference's full multi-head attention base
    (src/modules/models.py:29-74: bias-free Q / K / V / fc projections).  BASELINE.json config 4 only."""

    def __init__(self, num_head: int = 8, num_feature: int = 1024):
        super().__init__()
        self.num_head = num_head
        self.d_k = num_feature // num_head
        self.Q = nn.Linear(num_feature, num_feature, bias=False)
        self.K = nn.Linear(num_feature, num_feature, bias=False)
        self.V = nn.Linear(num_feature, num_feature, bias=False)
        self.fc = nn.Sequential(nn.Linear(num_feature, num_feature, bias=False), nn.Dropout(0.5))

    def forward(self, *inputs):  # pragma: no cover - guarded surface
        raise RuntimeError("edsnet_b200.AttentionExtractor only holds parameters; call DSNet.forward")


def layernorm_fold_operands(ln_w, ln_b, fc1_w, fc1_b, to_out_w, to_out_b):
    """Derived operands that fold LayerNorm(1024) into fc1 (dsnet.py:105-106; include/edsnet_b200.h, fc1_fold_*):

        fc1(LN(y)) = rstd * (z @ (W * gamma).T - mean(z) * rowsum(W * gamma)) + (W @ beta + b),   z = y - c

    for ANY per-row constant c (LayerNorm ignores it); mean / rstd are those of z.  The kernels use
    c = mean(x row) + mean(to_out bias), which is why the to_out bias arrives centred.  float64 sums, fp32 results."""
    fold_w = (fc1_w * ln_w[None, :]).contiguous()
    bc = (to_out_b.double() - to_out_b.double().mean()).float()
    return {
        "fc1_fold_w": fold_w,
        "fc1_fold_wgsum": fold_w.double().sum(1).float(),
        "fc1_fold_b": (fc1_w.double() @ ln_b.double() + fc1_b.double()).float(),
        "to_out_bc": bc,
        "to_out_bounds": (torch.stack([to_out_w.double().abs().sum(1).max(), bc.double().abs().max()]) * 1.001).float(),
    }


class DSNet(nn.Module):
    """`DSNet(base_model, num_feature, num_hidden, anchor_scales, num_head, fc_depth=5, orientation='paper',
    pooling_type='fft')` -- reference signature (dsnet.py:66-67).  Accelerated configuration only:
    base_model='nystromformer', pooling_type='roi', num_feature=1024, num_hidden=128, num_head=8, even scales.

    Extra keyword `precision`: arithmetic of the three big projections.
      'fp16x3' (default) tcgen05 tensor cores, fp16 hi/lo operand split, 3 MMA passes, fp32 accumulate: fp32-grade
      'fp16'             tcgen05 single pass, fp32 accumulate
      'fp32'             CUDA-core FFMA
    """

    def __init__(self, base_model, num_feature, num_hidden, anchor_scales, num_head, fc_depth=5,
                 orientation="paper", pooling_type="fft", precision: str = "fp16x3"):
        super().__init__()
        if type(anchor_scales) == int:                       # dsnet.py:69-70
            anchor_scales = [anchor_scales]
        anchor_scales = [int(s) for s in anchor_scales]
        if base_model not in _capi.BASE_MODELS:
            raise ValueError(f"edsnet_b200 accelerates base_model='nystromformer' (and 'attention' for the "
                             f"comparison config) only, got {base_model!r}")
        if pooling_type != "roi":
            raise ValueError(f"edsnet_b200 accelerates pooling_type='roi' only, got {pooling_type!r}")
        if (num_feature, num_hidden, num_head) != (NUM_FEATURE, NUM_HIDDEN, NUM_HEAD):
            raise ValueError("edsnet_b200 kernels are built for num_feature=1024, num_hidden=128, num_head=8")
        if precision not in _capi.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_capi.PRECISIONS)}")
        self.anchor_scales = anchor_scales
        self.num_scales = len(anchor_scales)
        self.base_model_type = base_model
        self.pooling_type = pooling_type
        self.fc_depth = int(fc_depth)
        self.precision = precision
        if base_model == "attention":
            self.base_model = AttentionExtractor(num_head, num_feature)
        else:
            self.base_model = NystromAttention(dim=num_feature, dim_head=DIM_HEAD, heads=num_head, num_landmarks=64,
                                               pinv_iterations=6, residual=True, residual_conv_kernel=33)
        self.layer_norm = nn.LayerNorm(num_feature)
        self.fc1 = nn.Linear(num_feature, num_hidden)
        self.fc_block = nn.Sequential(nn.Linear(num_hidden, num_hidden), nn.ReLU(), nn.Dropout(0.5),
                                      nn.LayerNorm(num_hidden))
        self.fc = nn.ModuleList([self.fc_block for _ in range(self.fc_depth)])    # ONE shared block, dsnet.py:96
        self.fc_cls = nn.Sequential(nn.Linear(num_hidden, 1))
        self.fc_loc = nn.Sequential(nn.Linear(num_hidden, 2))
        self._wcache = None
        self._wkey = None
        self._wlock = threading.Lock()
        self._workspace = None
        self._workspaces = {}
        self._drop_seed = None
        self._drop_offset = 0
        self._batch_cache = {}

    def _device_batch(self, batch, device) -> DeviceBatch:
        """Batch tables on the device for a list of video lengths (or a BatchPlan), cached per length tuple: repeated
        calls on the same videos (a training epoch, a CUDA-graph capture after its warm-up call) neither rebuild nor
        re-upload them -- and allocate no pinned staging memory, which is illegal under stream capture."""
        if isinstance(batch, DeviceBatch):
            return batch
        plan = batch if isinstance(batch, BatchPlan) else BatchPlan.build(batch)
        key = (str(device), tuple(int(t) for t in plan.lengths))
        hit = self._batch_cache.get(key)
        if hit is None:
            if len(self._batch_cache) >= 512:
                self._batch_cache.clear()
            hit = plan.to(device)
            self._batch_cache[key] = hit
        return hit

    def _next_dropout_stream(self):
        """(seed, offset) of the next train-mode forward's Philox dropout mask: the seed is torch's initial seed at the
        first use (torch.manual_seed makes runs repeatable), the offset counts the calls."""
        if self._drop_seed is None:
            self._drop_seed = int(torch.initial_seed()) & ((1 << 64) - 1)
        self._drop_offset += 1
        return self._drop_seed, self._drop_offset

    # caches and the lock are per-process state: a pickled / deep-copied model starts without them
    def __getstate__(self):
        st = self.__dict__.copy()
        st.update(_wcache=None, _wkey=None, _wlock=None, _workspace=None, _workspaces={}, _batch_cache={})
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._wlock = threading.Lock()

    # ------------------------------------------------------------------ weights / workspace
    def _named_weights(self):
        """edsnet_weights field -> parameter (fp32 tensors in the reference's state-dict layout)."""
        w = {"ln_w": self.layer_norm.weight, "ln_b": self.layer_norm.bias, "fc1_w": self.fc1.weight,
             "fc1_b": self.fc1.bias, "fcb_w": self.fc_block[0].weight, "fcb_b": self.fc_block[0].bias,
             "fcb_ln_w": self.fc_block[3].weight, "fcb_ln_b": self.fc_block[3].bias,
             "cls_w": self.fc_cls[0].weight, "cls_b": self.fc_cls[0].bias, "loc_w": self.fc_loc[0].weight,
             "loc_b": self.fc_loc[0].bias}
        bm = self.base_model
        if self.base_model_type == "attention":
            w.update(Q=bm.Q.weight, K=bm.K.weight, V=bm.V.weight, mha_fc_w=bm.fc[0].weight)
        else:
            w.update(to_qkv_w=bm.to_qkv.weight, to_out_w=bm.to_out[0].weight, to_out_b=bm.to_out[0].bias,
                     res_conv_w=bm.res_conv.weight)
        return w

    def _config(self) -> _capi.Config:
        # dsnet.py:113-115: an odd scale makes the reference's .view() raise RuntimeError; same surface here
        for s in self.anchor_scales:
            if s % 2:
                raise RuntimeError(f"odd anchor scale {s}: shape mismatch in view (reference dsnet.py:114 fails too)")
        return _capi.make_config(self.anchor_scales, self.fc_depth, _capi.PRECISIONS[self.precision],
                                 _capi.BASE_MODELS[self.base_model_type])

    def invalidate_weight_cache(self):
        """Forget the cached fp16 operand planes.  Needed after parameter updates that do not bump the tensors'
        version counters (CUDA-graph replays of an optimiser step)."""
        self._wkey = None

    def _weights(self, device, stream: int) -> _capi.Weights:
        with self._wlock:
            return self._weights_locked(device, stream)

    def _weights_locked(self, device, stream: int) -> _capi.Weights:
        named = self._named_weights()
        key = (self.precision, str(device)) + tuple((n, p.data_ptr(), p._version) for n, p in named.items())
        if self._wkey == key:
            # the operand planes were built on another stream: this one must not read them before they are complete
            built_on, ready = self._wcache[2], self._wcache[3]
            # (not under stream capture: an event recorded outside a capture cannot be waited on inside it, and
            # graphed_forward synchronises the device between its warm-up call and the capture anyway)
            if built_on != stream and not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(device).wait_event(ready)
            return self._wcache[0]
        keep = []
        w = _capi.Weights()
        tensors = {}
        for name, p in named.items():
            if p.device != device:
                raise RuntimeError(f"parameter {name} is on {p.device}, input is on {device}")
            if p.dtype != torch.float32:
                raise RuntimeError("edsnet_b200 parameters must be float32")
            tensors[name] = p.detach().contiguous()
        if self.base_model_type == "attention":
            tensors["mha_qkv_w"] = torch.cat([tensors.pop("Q"), tensors.pop("K"), tensors.pop("V")], dim=0).contiguous()
        for name, t in tensors.items():
            keep.append(t)
            setattr(w, name, t.data_ptr())
        if self.precision != "fp32":
            lib = _capi.lib()
            planes_of = (("mha_qkv_w16", "mha_qkv_w"), ("mha_fc_w16", "mha_fc_w")) \
                if self.base_model_type == "attention" else (("to_qkv_w16", "to_qkv_w"), ("to_out_w16", "to_out_w"))
            if self.base_model_type != "attention":
                derived = layernorm_fold_operands(tensors["ln_w"], tensors["ln_b"], tensors["fc1_w"], tensors["fc1_b"],
                                                  tensors["to_out_w"], tensors["to_out_b"])
                tensors["fc1_fold_w"] = derived.pop("fc1_fold_w")
                for field, t in derived.items():
                    t = t.contiguous()
                    keep.append(t)
                    setattr(w, field, t.data_ptr())
                keep.append(tensors["fc1_fold_w"])
                planes_of = planes_of + (("fc1_fold_w16", "fc1_fold_w"),)
            for field, src_name in planes_of + (("fc1_w16", "fc1_w"), ("fcb_w16", "fcb_w")):
                src = tensors[src_name]
                planes = torch.empty(lib.edsnet_split_f16_bytes(src.shape[0], src.shape[1]), dtype=torch.uint8,
                                     device=device)
                _capi.check(lib.edsnet_split_f16(src.data_ptr(), planes.data_ptr(), src.shape[0], src.shape[1],
                                                 stream))
                keep.append(planes)
                setattr(w, field, planes.data_ptr())
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(device))
        self._wcache = (w, keep, stream, ready)
        self._wkey = key
        return w

    def _get_workspace(self, cfg, total_rows: int, n_videos: int, device):
        """One workspace per (device, stream): forwards enqueued on different streams may overlap on the GPU."""
        need = _capi.lib().edsnet_workspace_bytes(cfg, total_rows, n_videos, None)
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        ws = self._workspaces.get(key)
        if ws is None or ws.numel() < need:
            self._workspaces.pop(key, None)
            ws = torch.empty(int(need), dtype=torch.uint8, device=device)
            self._workspaces[key] = ws
        self._workspace = ws                              # the most recently used one (tests read intermediates from it)
        return ws, need

    def launches_per_forward(self) -> int:
        """Kernel launches one forward enqueues (bench bookkeeping)."""
        return int(_capi.lib().edsnet_forward_launches(self._config()))

    @staticmethod
    def _check_input(x: torch.Tensor):
        if not isinstance(x, torch.Tensor):
            raise TypeError("expected a torch.Tensor")
        if not x.is_cuda:
            raise RuntimeError("edsnet_b200 has no CPU path: move the input (and the model) to a CUDA device")
        if x.dtype != torch.float32:
            raise RuntimeError("edsnet_b200 expects float32 features")

    # ------------------------------------------------------------------ packed multi-video API
    def forward_packed(self, x: torch.Tensor, batch: Union[DeviceBatch, BatchPlan, Sequence[int]]
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """x: [total_rows, 1024] float32 CUDA, rows of all videos concatenated.  Returns
        pred_cls [total_rows, S], pred_loc [total_rows, S, 2]; rows cu[v]..cu[v+1] belong to video v and equal
        the reference's `model(x_v[None])`."""
        self._check_input(x)
        if x.dim() != 2 or x.shape[1] != NUM_FEATURE:
            raise RuntimeError(f"expected [rows, {NUM_FEATURE}] features, got {tuple(x.shape)}")
        wants_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.training or wants_grad:
            # train(): Dropout(0.5) inside the shared fc block is active (dsnet.py:91-95); with gradients enabled the call
            # is differentiable with respect to the parameters.  Nystrom base: the training kernels (native_train.py:
            # edsnet_train_forward / edsnet_train_backward).  The comparison base ('attention') and gradients with respect
            # to the INPUT features are outside the hot path and go through the torch-op graph of autograd.py.
            if self.base_model_type == "nystromformer" and not x.requires_grad:
                from .native_train import scoring_with_native_grad
                return scoring_with_native_grad(self, x, batch)
            from .autograd import kernel_forward_with_grad, scoring_with_grad
            if self.training:
                return scoring_with_grad(self, x, batch)
            return kernel_forward_with_grad(self, x, batch)
        return self._forward_nograd(x, batch)

    def _forward_nograd(self, x, batch):
        if not isinstance(batch, DeviceBatch):
            batch = self._device_batch(batch, x.device)
        if batch.plan.total_rows != x.shape[0]:
            raise RuntimeError(f"batch plan covers {batch.plan.total_rows} rows, x has {x.shape[0]}")
        x = x.detach().contiguous()
        cfg = self._config()
        lib = _capi.lib()
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            w = self._weights(x.device, stream)
            ws, need = self._get_workspace(cfg, x.shape[0], batch.plan.n_videos, x.device)
            S = self.num_scales
            pred_cls = torch.empty((x.shape[0], S), dtype=torch.float32, device=x.device)
            pred_loc = torch.empty((x.shape[0], S, 2), dtype=torch.float32, device=x.device)
            _capi.check(lib.edsnet_forward(cfg, w, batch.struct, x.data_ptr(), pred_cls.data_ptr(),
                                           pred_loc.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return pred_cls, pred_loc

    def decode_packed(self, pred_loc: torch.Tensor, batch: DeviceBatch) -> Tuple[torch.Tensor, torch.Tensor]:
        """Offsets -> (float32 left/right boxes, clipped+rounded int32 boxes), both [total_rows*S, 2]
        (dsnet.py:146-153 + evaluate.py:26)."""
        cfg = self._config()
        n = batch.plan.total_rows * self.num_scales
        dev = pred_loc.device
        boxes_f = torch.empty((n, 2), dtype=torch.float32, device=dev)
        boxes_i = torch.empty((n, 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().edsnet_decode_boxes(cfg, batch.struct, pred_loc.contiguous().data_ptr(),
                                                        boxes_f.data_ptr(), boxes_i.data_ptr(), stream))
        return boxes_f, boxes_i

    def nms_packed(self, pred_cls: torch.Tensor, pred_loc: torch.Tensor, batch: DeviceBatch,
                   nms_thresh: float = 0.5):
        """decode + clip/round + greedy temporal NMS for every video of the batch, on device.
        Returns dict of device tensors: keep_count [V], keep_idx / keep_scores [total_rows*S],
        keep_boxes [total_rows*S, 2]; video v's kept proposals (descending score) start at cu_rows[v]*S."""
        cfg = self._config()
        plan = batch.plan
        dev = pred_cls.device
        S = self.num_scales
        n = plan.total_rows * S
        out = {
            "boxes_i32": torch.empty((n, 2), dtype=torch.int32, device=dev),
            "keep_count": torch.empty((plan.n_videos,), dtype=torch.int32, device=dev),
            "keep_idx": torch.empty((n,), dtype=torch.int32, device=dev),
            "keep_scores": torch.empty((n,), dtype=torch.float32, device=dev),
            "keep_boxes": torch.empty((n, 2), dtype=torch.int32, device=dev),
        }
        off_ptr, scratch_ptr, keep = None, None, None
        if plan.max_rows * S > 4096:
            off, total = plan.nms_scratch(S)
            off_t = torch.from_numpy(off).to(dev)
            scratch = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            keep = (off_t, scratch)
            off_ptr, scratch_ptr = off_t.data_ptr(), scratch.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().edsnet_decode_nms(
                cfg, batch.struct, pred_cls.contiguous().data_ptr(), pred_loc.contiguous().data_ptr(),
                float(nms_thresh), None, out["boxes_i32"].data_ptr(), out["keep_count"].data_ptr(),
                out["keep_idx"].data_ptr(), out["keep_scores"].data_ptr(), out["keep_boxes"].data_ptr(),
                off_ptr, scratch_ptr, stream))
        out["_scratch"] = keep
        return out

    def proposals_packed(self, x: torch.Tensor, batch: Union[DeviceBatch, BatchPlan, Sequence[int]],
                         nms_thresh: float = 0.5) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Features -> per video (keep_scores, keep_boxes) exactly as evaluate.py:24-28 produces them."""
        if not isinstance(batch, DeviceBatch):
            plan = batch if isinstance(batch, BatchPlan) else BatchPlan.build(batch)
            batch = plan.to(x.device)
        with torch.no_grad():
            cls, loc = self._forward_nograd(x, batch)
            r = self.nms_packed(cls, loc, batch, nms_thresh)
        counts = r["keep_count"].cpu().numpy()
        ks = r["keep_scores"].cpu().numpy()
        kb = r["keep_boxes"].cpu().numpy()
        _capi.raise_on_tc_timeout()
        S = self.num_scales
        out = []
        for v in range(batch.plan.n_videos):
            o = int(batch.plan.cu_rows[v]) * S
            c = int(counts[v])
            out.append((ks[o:o + c].copy(), kb[o:o + c].copy()))
        return out

    # ------------------------------------------------------------------ CUDA-graph replay for repeated shapes
    def graphed_forward(self, lengths: Sequence[int], device=None):
        """Capture `forward_packed` for one fixed list of video lengths into a CUDA graph and return
        `run(x) -> (pred_cls, pred_loc)`.  A single TVSum-sized video is 11 launches of a few microseconds each, so the
        eager call is bound by launch latency; a replay submits them with one driver call.  `run` copies x into the
        graph's static input and returns views of its static outputs (overwritten by the next call).  The graph holds
        the weights' operand planes of the moment: capture again after a parameter update.  eval() / no_grad only."""
        if self.training:
            raise RuntimeError("graphed_forward replays the inference kernels: call model.eval() first")
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("graphed_forward needs the model on a CUDA device (there is no CPU fallback)")
        plan = BatchPlan.build([int(t) for t in lengths])
        batch = plan.to(dev)
        x_static = torch.zeros((plan.total_rows, NUM_FEATURE), dtype=torch.float32, device=dev)
        with torch.no_grad():
            self._forward_nograd(x_static, batch)                      # warm-up: weight planes, workspace, smem opt-ins
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    out = self._forward_nograd(x_static, batch)
            torch.cuda.current_stream(dev).wait_stream(side)
        keep = (batch, self._workspace, self._wcache)                  # everything the captured launches point at

        def run(x: torch.Tensor):
            self._check_input(x)
            x_static.copy_(x.reshape(plan.total_rows, NUM_FEATURE))
            graph.replay()
            return out

        run.graph, run.keep = graph, keep
        return run

    # ------------------------------------------------------------------ reference surface
    def forward(self, x: torch.Tensor):
        """x: (1, T, 1024) float32 on a CUDA device -> pred_cls (T, S), pred_loc (T, S, 2)   (dsnet.py:100-115)."""
        self._check_input(x)
        if x.dim() != 3:
            raise ValueError(f"not enough values to unpack: expected a (1, T, {NUM_FEATURE}) tensor")
        if x.shape[0] != 1:
            # the reference's cat(dim=0)/view(seq_len, num_scales) (dsnet.py:113-115) only works for batch 1
            raise RuntimeError(f"shape '[{x.shape[1]}, {self.num_scales}]' is invalid for a batch of {x.shape[0]} "
                               "videos: DSNet.forward scores one video per call; use forward_packed for many")
        if x.shape[1] < 1:
            raise RuntimeError("empty sequence")
        return self.forward_packed(x[0], [x.shape[1]])

    def predict(self, seq: torch.Tensor):
        """(scores float32 (T*S,), boxes float32 (T*S, 2) left/right), NumPy, as dsnet.py:140-153."""
        pred_cls, pred_loc = self(seq)
        with torch.no_grad():
            batch = BatchPlan.build([seq.shape[1]]).to(seq.device)
            boxes_f, _ = self.decode_packed(pred_loc.detach(), batch)
        out = pred_cls.detach().cpu().numpy().reshape(-1), boxes_f.cpu().numpy().reshape(-1, 2)
        _capi.raise_on_tc_timeout()
        return out

    def proposals(self, seq: torch.Tensor, nms_thresh: float = 0.5):
        """evaluate.py:24-28 in one call: predict -> clip/round -> nms.  Returns (keep_scores, keep_boxes)."""
        self._check_input(seq)
        return self.proposals_packed(seq[0], [seq.shape[1]], nms_thresh)[0]
