"""Host side of the native training step (BASELINE.json config 3): ctypes calls into the train-mode forward, the loss
gradient, the backward and the Adam kernels of libedsnet_b200.so (include/edsnet_b200.h, "training step").

What it mirrors in the reference (paths relative to /root/reference/src):
  train_forward / _NativeScoring   `pred_cls, pred_loc = model(seq)` in train() mode, anchor_based/train.py:116 with
                                   anchor_based/dsnet.py:100-115 (Dropout(0.5) of the shared fc block active, :91-95)
  loss_and_grad                    calc_cls_loss / calc_loc_loss, anchor_based/losses.py:5-57, combined as train.py:119-123
  train_backward                   `loss.backward()`, train.py:126
  NativeDataParallelStep           the loop body train.py:110-128 for k videos per rank, made data parallel with ONE flat
                                   NCCL all-reduce of the gradient (the reference has no collective), then Adam (:53-55, :127)

`_NativeScoring` is a torch.autograd.Function, so the reference's own training loop works unchanged on the drop-in model:
its torch losses produce d loss / d pred_cls, d loss / d pred_loc, and `loss.backward()` lands in the backward kernels.
No torch op computes any part of the model here; torch provides device memory, streams and the autograd plumbing.
"""
from __future__ import annotations

import collections
import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from .plan import BatchPlan, DeviceBatch

_SEED_MASK = (1 << 64) - 1


def _fp32_weights(model, device) -> Tuple[_capi.Weights, list]:
    """edsnet_weights with the fp32 pointers only (the training forward builds the operand planes of the step itself)."""
    w = _capi.Weights()
    keep = []
    for name, p in model._named_weights().items():
        if p.device != device:
            raise RuntimeError(f"parameter {name} is on {p.device}, input is on {device}")
        if p.dtype != torch.float32:
            raise RuntimeError("edsnet_b200 parameters must be float32")
        t = p.detach()
        if not t.is_contiguous():
            raise RuntimeError(f"parameter {name} must be contiguous for the training kernels")
        keep.append(t)
        setattr(w, name, t.data_ptr())
    return w, keep


def _grads_struct(tensors: dict) -> _capi.Grads:
    g = _capi.Grads()
    for name in _capi.GRAD_FIELDS:
        t = tensors[name]
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"gradient buffer {name} must be contiguous float32")
        setattr(g, name, t.data_ptr())
    return g


class TrainContext:
    """Everything one forward leaves for its backward: the workspace with the saved activations, the batch tables, the
    inputs and outputs the backward reads."""
    __slots__ = ("cfg", "batch", "x", "pred_cls", "pred_loc", "workspace", "dropout", "weights", "keep")


def _as_batch(batch, device) -> DeviceBatch:
    if isinstance(batch, DeviceBatch):
        return batch
    plan = batch if isinstance(batch, BatchPlan) else BatchPlan.build(batch)
    return plan.to(device)


def train_forward(model, x: torch.Tensor, batch, dropout: bool, seed: int, offset: int,
                  offset_dev: Optional[torch.Tensor] = None) -> TrainContext:
    """x: packed [rows, 1024] float32 CUDA.  Runs edsnet_train_forward; returns the context (ctx.pred_cls, ctx.pred_loc).
    offset_dev: optional int64 device scalar added to `offset` at run time (fresh dropout masks under graph replay)."""
    model._check_input(x)
    if model.base_model_type != "nystromformer":
        raise RuntimeError("the native training kernels cover base_model='nystromformer' (the hot path)")
    batch = _as_batch(batch, x.device)
    if batch.plan.total_rows != x.shape[0]:
        raise RuntimeError(f"batch plan covers {batch.plan.total_rows} rows, x has {x.shape[0]}")
    cfg = _capi.make_config(model.anchor_scales, model.fc_depth, _capi.PREC_FP16X3, 0)
    model._config()                                   # same validation surface as the inference path (odd scales raise)
    lib = _capi.lib()
    x = x.detach().contiguous()
    ctx = TrainContext()
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        w, keep = _fp32_weights(model, x.device)
        need = lib.edsnet_train_workspace_bytes(cfg, x.shape[0], batch.plan.n_videos, None)
        ws = torch.empty(int(need), dtype=torch.uint8, device=x.device)
        S = model.num_scales
        pred_cls = torch.empty((x.shape[0], S), dtype=torch.float32, device=x.device)
        pred_loc = torch.empty((x.shape[0], S, 2), dtype=torch.float32, device=x.device)
        _capi.check(lib.edsnet_train_forward(cfg, w, batch.struct, x.data_ptr(), 1 if dropout else 0,
                                             seed & _SEED_MASK, offset & _SEED_MASK,
                                             offset_dev.data_ptr() if offset_dev is not None else None,
                                             pred_cls.data_ptr(), pred_loc.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    ctx.cfg, ctx.batch, ctx.x, ctx.pred_cls, ctx.pred_loc = cfg, batch, x, pred_cls, pred_loc
    ctx.workspace, ctx.dropout, ctx.weights, ctx.keep = ws, bool(dropout), w, keep
    return ctx


def loss_and_grad(ctx: TrainContext, cls_label: torch.Tensor, loc_label: torch.Tensor, lambda_reg: float = 1.0,
                  scale: float = 1.0):
    """cls_label int32 [rows, S], loc_label float32 [rows, S, 2] on the device.  Returns (loss [V, 3] = total / cls / loc per
    video, d_logit [rows, S], d_loc [rows, S, 2]) -- the gradients already multiplied by `scale`."""
    dev = ctx.x.device
    if cls_label.dtype != torch.int32 or loc_label.dtype != torch.float32:
        raise RuntimeError("labels: cls_label int32, loc_label float32")
    if tuple(cls_label.shape) != tuple(ctx.pred_cls.shape) or tuple(loc_label.shape) != tuple(ctx.pred_loc.shape):
        raise RuntimeError("label shapes do not match the predictions")
    cls_label, loc_label = cls_label.contiguous(), loc_label.contiguous()
    V = ctx.batch.plan.n_videos
    loss = torch.empty((V, 3), dtype=torch.float32, device=dev)
    d_logit = torch.empty_like(ctx.pred_cls)
    d_loc = torch.empty_like(ctx.pred_loc)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().edsnet_loss_grad(ctx.cfg, ctx.batch.struct, ctx.pred_cls.data_ptr(), ctx.pred_loc.data_ptr(),
                                                 cls_label.data_ptr(), loc_label.data_ptr(), float(lambda_reg), float(scale),
                                                 d_logit.data_ptr(), d_loc.data_ptr(), loss.data_ptr(), stream))
    return loss, d_logit, d_loc


def train_backward(ctx: TrainContext, d_cls: torch.Tensor, d_loc: torch.Tensor, grads: dict, logit_grad: bool,
                   side_stream: Optional[torch.cuda.Stream] = None) -> None:
    """grads: {edsnet_weights field name: ZERO-filled float32 tensor of the parameter's shape}; filled on return.
    side_stream: second stream for the independent part of the backward (forked / joined inside the call)."""
    dev = ctx.x.device
    d_cls, d_loc = d_cls.contiguous(), d_loc.contiguous()
    if d_cls.dtype != torch.float32 or d_loc.dtype != torch.float32:
        raise RuntimeError("output gradients must be float32")
    g = _grads_struct(grads)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().edsnet_train_backward(ctx.cfg, ctx.weights, ctx.batch.struct, ctx.x.data_ptr(),
                                                      ctx.pred_cls.data_ptr(), d_cls.data_ptr(), d_loc.data_ptr(),
                                                      1 if logit_grad else 0, 1 if ctx.dropout else 0, g,
                                                      ctx.workspace.data_ptr(), ctx.workspace.numel(), stream,
                                                      side_stream.cuda_stream if side_stream is not None else None))
        # (the call joins the side stream back into `stream` before it returns: stream order on `stream` alone keeps the
        # workspace and the gradient buffers safe)


class _NativeScoring(torch.autograd.Function):
    """model(seq) with gradients: forward = edsnet_train_forward, backward = edsnet_train_backward."""

    @staticmethod
    def forward(ctx, model, batch, dropout, seed, offset, x, *params):
        tctx = train_forward(model, x, batch, dropout, seed, offset)
        pred_cls, pred_loc = tctx.pred_cls, tctx.pred_loc
        tctx.pred_cls = tctx.pred_loc = None           # outputs go through save_for_backward (no reference cycle)
        ctx.save_for_backward(pred_cls)
        ctx.tctx = tctx
        ctx.shapes = [(p.shape, p.requires_grad) for p in params]
        return pred_cls, pred_loc

    @staticmethod
    def backward(ctx, g_cls, g_loc):
        tctx = ctx.tctx
        (tctx.pred_cls,) = ctx.saved_tensors
        dev = tctx.x.device
        grads = {name: torch.zeros(shape, dtype=torch.float32, device=dev)
                 for name, (shape, _) in zip(_capi.GRAD_FIELDS, ctx.shapes)}
        g_cls = torch.zeros_like(tctx.pred_cls) if g_cls is None else g_cls
        if g_loc is None:
            g_loc = torch.zeros(tuple(tctx.pred_cls.shape) + (2,), dtype=torch.float32, device=dev)
        train_backward(tctx, g_cls, g_loc, grads, logit_grad=False)
        out = tuple(grads[name] if req else None for name, (_, req) in zip(_capi.GRAD_FIELDS, ctx.shapes))
        return (None, None, None, None, None, None) + out


class _NativeEvalScoring(torch.autograd.Function):
    """eval()-mode call with gradients enabled.  The VALUES come from the inference kernels (bit-identical to the same
    call under torch.no_grad(), whatever `precision` the model runs in); the backward recomputes the training forward
    without dropout -- the same function -- to get the activations edsnet_train_backward needs."""

    @staticmethod
    def forward(ctx, model, batch, x, *params):
        ctx.model, ctx.batch = model, batch
        ctx.save_for_backward(x)
        ctx.shapes = [(p.shape, p.requires_grad) for p in params]
        with torch.no_grad():
            return model._forward_nograd(x, batch)

    @staticmethod
    def backward(ctx, g_cls, g_loc):
        (x,) = ctx.saved_tensors
        tctx = train_forward(ctx.model, x, ctx.batch, False, 0, 0)
        dev = x.device
        grads = {name: torch.zeros(shape, dtype=torch.float32, device=dev)
                 for name, (shape, _) in zip(_capi.GRAD_FIELDS, ctx.shapes)}
        g_cls = torch.zeros_like(tctx.pred_cls) if g_cls is None else g_cls
        g_loc = torch.zeros_like(tctx.pred_loc) if g_loc is None else g_loc
        train_backward(tctx, g_cls, g_loc, grads, logit_grad=False)
        out = tuple(grads[name] if req else None for name, (_, req) in zip(_capi.GRAD_FIELDS, ctx.shapes))
        return (None, None, None) + out


def scoring_with_native_grad(model, x: torch.Tensor, batch):
    """The differentiable call behind DSNet.forward / forward_packed: Dropout follows model.training."""
    if x.requires_grad:
        raise RuntimeError("edsnet_b200 trains the model parameters; the input features get no gradient "
                           "(the reference's features are data: anchor_based/train.py:113)")
    named = model._named_weights()
    params = [named[k] for k in _capi.GRAD_FIELDS]
    batch = _as_batch(batch, x.device) if isinstance(batch, DeviceBatch) else model._device_batch(batch, x.device)
    if not model.training:
        return _NativeEvalScoring.apply(model, batch, x, *params)
    seed, offset = model._next_dropout_stream()
    p_drop = float(model.fc_block[2].p)
    if p_drop not in (0.0, 0.5):
        raise RuntimeError(f"the training kernels implement the reference's Dropout(0.5) (or p = 0), got p = {p_drop}")
    dropout = bool(model.training) and p_drop > 0.0
    return _NativeScoring.apply(model, batch, dropout, seed, offset, x, *params)


class NativeDataParallelStep:
    """One optimiser step over the videos of every rank, nothing but kernels of libedsnet_b200.so and ONE collective:
    train-mode forward, loss gradient, backward into a flat gradient buffer, all-reduce (sum; NCCL over NVLink on GPUs),
    Adam on the flat parameter buffer with the 1 / world factor folded in.  loss = mean over the step's videos of
    cls_loss + lambda_reg * loc_loss (anchor_based/train.py:119-123); lr 5e-5, weight decay 1e-5 as train.py:53-55.

    The parameters become views of one flat buffer (so do their .grad), in the order of edsnet_grads.

    use_graphs: the forward + loss + backward launch sequence (about 45 kernels of a few microseconds each for one
    TVSum-sized video: launch bound from the host) is captured into a CUDA graph per tuple of video lengths the second
    time that tuple is seen and replayed afterwards; features and labels are copied into the graph's static buffers, the
    dropout mask still changes every step (the Philox offset lives in device memory).  At most `max_graphs` graphs are
    kept (least recently used first out); a dataset of N videos taken k at a time in a fixed order needs N / k."""

    def __init__(self, model, lr: float = 5e-5, weight_decay: float = 1e-5, lambda_reg: float = 1.0, world_size: int = 1,
                 group=None, betas=(0.9, 0.999), eps: float = 1e-8, dropout: bool = True, seed: Optional[int] = None,
                 use_graphs: bool = True, max_graphs: int = 64):
        named = model._named_weights()
        params = [named[k] for k in _capi.GRAD_FIELDS]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("NativeDataParallelStep needs the model on a CUDA device (there is no CPU fallback)")
        self.model, self.device = model, dev
        self.lr, self.weight_decay, self.lambda_reg = float(lr), float(weight_decay), float(lambda_reg)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.world_size, self.group = int(world_size), group
        self.dropout = bool(dropout)
        self.seed = int(torch.initial_seed() if seed is None else seed) & _SEED_MASK
        self.n_params = int(sum(p.numel() for p in params))
        self.flat_param = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.grad_views = {}
        o = 0
        for name, p in zip(_capi.GRAD_FIELDS, params):
            n = p.numel()
            p.data = self.flat_param[o:o + n].view_as(p)
            self.grad_views[name] = self.flat_grad[o:o + n].view_as(p)
            p.grad = self.grad_views[name]
            o += n
        self.step_count = 0
        self.skip_allreduce = False
        self.use_graphs, self.max_graphs = bool(use_graphs), int(max_graphs)
        self._loss = None
        self._graphs = collections.OrderedDict()
        self._seen = set()
        self._offset_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._side = torch.cuda.Stream(dev)
        self._bwd_side = torch.cuda.Stream(dev)        # independent part of the backward (pinv chain re-run, dW products)
        self.graph_replays = 0
        fwd, bwd = C.c_int32(0), C.c_int32(0)
        _capi.check(_capi.lib().edsnet_train_launches(model._config(), C.byref(fwd), C.byref(bwd)))
        self.launches_per_step = int(fwd.value) + int(bwd.value) + 3          # + gradient memset, loss gradient, Adam

    # ---- labels: host NumPy (anchor_labels / LabelCache) or tensors -> one int32 and one float32 array per step ----
    @staticmethod
    def _host_labels(cls_labels, loc_labels):
        def to_np(a):
            return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        cl = np.ascontiguousarray(np.concatenate([to_np(c) for c in cls_labels]), dtype=np.int32)
        ll = np.ascontiguousarray(np.concatenate([to_np(l) for l in loc_labels]), dtype=np.float32)
        return cl, ll

    def _forward_backward(self, x, batch, cl, ll, k, offset, offset_dev):
        """zero gradient, train-mode forward, loss gradient (scaled by 1 / local videos), backward -> flat_grad."""
        self.flat_grad.zero_()
        ctx = train_forward(self.model, x, batch, self.dropout, self.seed, offset, offset_dev)
        loss, d_logit, d_loc = loss_and_grad(ctx, cl, ll, self.lambda_reg, 1.0 / k)
        train_backward(ctx, d_logit, d_loc, self.grad_views, logit_grad=True, side_stream=self._bwd_side)
        return ctx, loss, d_logit, d_loc

    def backward_only(self, seqs: Sequence[torch.Tensor], cls_labels, loc_labels) -> torch.Tensor:
        """Forward + loss + backward of this rank's videos into flat_grad (already divided by the local video count);
        returns the per-video losses [k, 3] (device).  No collective, no update, no graph."""
        lengths = tuple(int(s.shape[0]) for s in seqs)
        x = seqs[0] if len(seqs) == 1 else torch.cat(list(seqs))
        batch = self.model._device_batch(lengths, self.device)
        cl, ll = self._host_labels(cls_labels, loc_labels)
        cl, ll = torch.from_numpy(cl).to(self.device), torch.from_numpy(ll).to(self.device)
        _, loss, _, _ = self._forward_backward(x, batch, cl.reshape(x.shape[0], -1), ll.reshape(x.shape[0], -1, 2),
                                               len(seqs), self.step_count, None)
        return loss

    def _capture(self, lengths, k):
        dev, S = self.device, self.model.num_scales
        R = int(sum(lengths))
        e = {"x": torch.zeros((R, 1024), dtype=torch.float32, device=dev),
             "cl": torch.zeros((R, S), dtype=torch.int32, device=dev),
             "ll": torch.zeros((R, S, 2), dtype=torch.float32, device=dev),
             "cl_pin": torch.zeros((R, S), dtype=torch.int32).pin_memory(),
             "ll_pin": torch.zeros((R, S, 2), dtype=torch.float32).pin_memory(),
             "copied": torch.cuda.Event()}
        batch = self.model._device_batch(lengths, dev)
        graph = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream(dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            with torch.cuda.graph(graph, stream=self._side):
                e["keep"] = self._forward_backward(e["x"], batch, e["cl"], e["ll"], k, 0, self._offset_dev)
        cur.wait_stream(self._side)
        e["graph"], e["batch"], e["loss"] = graph, batch, e["keep"][1]
        e["copied"].record(cur)
        return e

    def step(self, seqs: Sequence[torch.Tensor], cls_labels, loc_labels) -> torch.Tensor:
        """Returns the per-video losses [k, 3] of this rank (device tensor, overwritten by a later step on the same
        length tuple when graphs are in use)."""
        self.model.train()
        lengths = tuple(int(s.shape[0]) for s in seqs)
        entry = None
        if self.use_graphs:
            entry = self._graphs.get(lengths)
            if entry is None and lengths in self._seen:
                entry = self._capture(lengths, len(seqs))          # second sighting: everything lazy has run once
                self._graphs[lengths] = entry
                if len(self._graphs) > self.max_graphs:
                    self._graphs.popitem(last=False)
            self._seen.add(lengths)
        if entry is not None:
            self._graphs.move_to_end(lengths)
            cl, ll = self._host_labels(cls_labels, loc_labels)
            entry["copied"].synchronize()                          # the previous upload out of the pinned buffers is done
            entry["cl_pin"].numpy().reshape(-1)[:] = cl.reshape(-1)
            entry["ll_pin"].numpy().reshape(-1)[:] = ll.reshape(-1)
            entry["cl"].copy_(entry["cl_pin"], non_blocking=True)
            entry["ll"].copy_(entry["ll_pin"], non_blocking=True)
            entry["copied"].record(torch.cuda.current_stream(self.device))
            if len(seqs) == 1:
                entry["x"].copy_(seqs[0])
            else:
                torch.cat(list(seqs), out=entry["x"])
            self._offset_dev.fill_(self.step_count)
            entry["graph"].replay()
            self.graph_replays += 1
            self._loss = entry["loss"]
        else:
            self._loss = self.backward_only(seqs, cls_labels, loc_labels)
        if self.world_size > 1 and not self.skip_allreduce:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.step_count += 1
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _capi.check(_capi.lib().edsnet_adam_step(
                self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.n_params, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                1.0 / self.world_size, stream))
        self.model.invalidate_weight_cache()          # the kernels wrote the weights behind torch's version counters
        return self._loss

    def last_loss(self) -> float:
        """Mean over the last step's local videos of cls + lambda * loc (synchronises)."""
        if self._loss is None:
            return float("nan")
        v = float(self._loss[:, 0].mean().item())
        _capi.raise_on_tc_timeout()
        return v
