"""Torch-op graph of the scoring model, kept for the two cases that are NOT on the hot path.

Training of the Nystrom model runs on the native kernels (native_train.py,
csrc/train.cuh).  This graph remains for (i) the comparison base ('attention', BASELINE config 4) in train() mode and
(ii) gradients with respect to the INPUT features.  It runs on the GPU, is numerically the same function (same
parameter tensors, same operation order as src/transformer/nystroformer.py:67-150 and
src/anchor_based/dsnet.py:100-115) and supports the train-mode Dropout(0.5) of the shared fc block.  It is never used
when `torch.no_grad()` + `eval()` hold (evaluate.py:15-17).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _attention(bm, x):
    """x: (T, 1024) -> (T, 1024); landmark attention with front zero padding (nystroformer.py:72-75)."""
    T = x.shape[0]
    h, m, d = bm.heads, bm.num_landmarks, 64
    pad = (m - T % m) % m
    xp = F.pad(x, (0, 0, pad, 0)) if pad else x
    n = T + pad
    seg = n // m
    q, k, v = (t.reshape(n, h, d).permute(1, 0, 2) for t in bm.to_qkv(xp).chunk(3, dim=-1))
    q = q * bm.scale
    ql = q.reshape(h, m, seg, d).sum(dim=2) / seg
    kl = k.reshape(h, m, seg, d).sum(dim=2) / seg
    a1 = torch.softmax(q @ kl.transpose(1, 2), dim=-1)
    a2 = torch.softmax(ql @ kl.transpose(1, 2), dim=-1)
    a3 = torch.softmax(ql @ k.transpose(1, 2), dim=-1)
    mag = a2.abs()
    z = a2.transpose(-1, -2) / (mag.sum(dim=-1).max() * mag.sum(dim=-2).max())     # nystroformer.py:16-19
    eye = torch.eye(m, device=x.device, dtype=x.dtype)
    for _ in range(bm.pinv_iterations):
        az = a2 @ z
        z = 0.25 * z @ (13 * eye - az @ (15 * eye - az @ (7 * eye - az)))
    out = (a1 @ z) @ (a3 @ v)
    out = out + bm.res_conv(v.unsqueeze(0))[0]
    out = out.permute(1, 0, 2).reshape(n, h * d)
    return bm.to_out(out)[pad:]


def _mha(bm, x):
    """Full multi-head attention base (src/modules/models.py:46-65); both Dropout(0.5) follow the module's mode."""
    T, Fd = x.shape
    h, dk = bm.num_head, bm.d_k
    q, k, v = (m(x).reshape(T, h, dk).permute(1, 0, 2) for m in (bm.Q, bm.K, bm.V))
    attn = torch.softmax((q @ k.transpose(1, 2)) / dk ** 0.5, dim=-1)
    attn = F.dropout(attn, 0.5, bm.training)
    y = (attn @ v).permute(1, 0, 2).reshape(T, Fd)
    return bm.fc(y)


def _score_one(model, x):
    out = (_mha(model.base_model, x) if model.base_model_type == "attention" else _attention(model.base_model, x)) + x
    out = model.fc1(model.layer_norm(out))
    for fc in model.fc:
        out = fc(out)
    ut = out.t().unsqueeze(0)
    pooled = torch.stack([F.avg_pool1d(ut, s, stride=1, padding=s // 2)[0].t()[:-1] for s in model.anchor_scales], 1)
    T = x.shape[0]
    return (model.fc_cls(pooled).sigmoid().view(T, model.num_scales),
            model.fc_loc(pooled).view(T, model.num_scales, 2))


def scoring_with_grad(model, x, batch):
    """x: packed [rows, 1024]; batch: lengths / BatchPlan / DeviceBatch.  Returns autograd-tracked outputs."""
    from .plan import BatchPlan, DeviceBatch
    if isinstance(batch, DeviceBatch):
        lengths = batch.plan.lengths
    elif isinstance(batch, BatchPlan):
        lengths = batch.lengths
    else:
        lengths = batch
    model._config()                      # same validation surface as the kernel path (odd scales raise)
    cls, loc, o = [], [], 0
    for t in lengths:
        t = int(t)
        c, l = _score_one(model, x[o:o + t])
        cls.append(c)
        loc.append(l)
        o += t
    return torch.cat(cls), torch.cat(loc)


class _KernelForward(torch.autograd.Function):
    """eval()-mode call with gradients enabled: values from the CUDA kernels, gradients by recomputing the torch-op
    graph above in backward (no dropout in eval mode, so both describe the same function)."""

    @staticmethod
    def forward(ctx, model, batch, x, *params):
        ctx.model, ctx.batch = model, batch
        ctx.save_for_backward(x)
        with torch.no_grad():
            cls, loc = model._forward_nograd(x, batch)
        return cls, loc

    @staticmethod
    def backward(ctx, g_cls, g_loc):
        (x,) = ctx.saved_tensors
        model = ctx.model
        params = [p for p in model.parameters()]
        with torch.enable_grad():
            xin = x.detach().requires_grad_(ctx.needs_input_grad[2])
            cls, loc = scoring_with_grad(model, xin, ctx.batch)
            wanted = ([xin] if ctx.needs_input_grad[2] else []) + [p for p in params if p.requires_grad]
            grads = torch.autograd.grad([cls, loc], wanted, [g_cls, g_loc], allow_unused=True)
        grads = list(grads)
        gx = grads.pop(0) if ctx.needs_input_grad[2] else None
        out = []
        for p in params:
            out.append(grads.pop(0) if p.requires_grad else None)
        return (None, None, gx, *out)


def kernel_forward_with_grad(model, x, batch):
    return _KernelForward.apply(model, batch, x, *list(model.parameters()))
