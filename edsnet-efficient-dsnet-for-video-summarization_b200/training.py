"""Training-side pieces of the anchor-based path (BASELINE.json config 3): anchor label assignment, the cls / loc
losses and a data-parallel optimiser step with ONE flat gradient all-reduce (NCCL over NVLink on GPUs).

What each function restates (reference files relative to /root/reference/src):
  mask_to_segments     helpers/bbox_helper.py:34-46 (seq2bbox) + :8-18 (lr2cw)
  anchor_labels        anchor_based/anchor_helper.py:22-50 (get_pos_label), :53-71 (get_neg_label), :96-112
                       (bbox2offset) and the sampling recipe of anchor_based/train.py:91-108
  cls_loss / loc_loss  anchor_based/losses.py:32-57 / :5-29
  DataParallelStep     the per-video loop body of anchor_based/train.py:110-128, batched over the videos of a rank and
                       made data parallel: the reference has no collective at all; averaging the
                       gradient over W ranks turns its batch-1 SGD into batch-W, which is what config 3 asks for.
  NativeDataParallelStep  (native_train.py) the same step on the training kernels of libedsnet_b200.so: train-mode forward,
                       loss gradient, backward, ONE flat NCCL all-reduce, Adam -- no torch op computes any part of it.
DataParallelStep drives the model through torch autograd (model(x) in train() mode lands in the same kernels via
native_train._NativeScoring; the losses and Adam are torch ops there): it is the reference-shaped loop,
NativeDataParallelStep is the fast path bench.py --config c3 measures (and replays as CUDA graphs).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


from .native_train import NativeDataParallelStep      # noqa: E402,F401  (re-exported: training.NativeDataParallelStep)


# ----------------------------------------------------------------------------------------------- labels (host, NumPy)
def mask_to_segments(mask: np.ndarray) -> np.ndarray:
    """Binary keyshot mask -> [center, width] float32 boxes of its runs of ones."""
    m = np.concatenate([[0], np.asarray(mask, dtype=bool).astype(np.int8), [0]])
    d = np.diff(m)
    start, end = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
    lr = np.stack([start, end], axis=1).astype(np.int32)
    return np.stack([(lr[:, 0] + lr[:, 1]) / 2, lr[:, 1] - lr[:, 0]], axis=1).astype(np.float32)


def _hull_iou(anchors_cw: np.ndarray, target_cw: np.ndarray) -> np.ndarray:
    """Overlap of every anchor with one target: intersection / hull, in the float32 the reference computes in
    (cw2lr casts to float32, helpers/bbox_helper.py:21-31,49-75)."""
    a = anchors_cw.astype(np.float32)
    t = target_cw.astype(np.float32)
    al, ar = a[:, 0] - a[:, 1] / 2, a[:, 0] + a[:, 1] / 2
    tl, tr = t[0] - t[1] / 2, t[0] + t[1] / 2
    inter = np.minimum(ar, tr) - np.maximum(al, tl)
    inter[inter < 0] = 0
    hull = np.maximum(ar, tr) - np.minimum(al, tl)
    hull[hull <= 0] = 1e-6
    return inter / hull


def positive_labels(seq_len: int, scales: Sequence[int], targets_cw: np.ndarray, iou_thresh: float
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """cls_label (T, S) int32 in {0, 1}, loc_label (T, S, 2) float64; a later target overrides an earlier one."""
    S = len(scales)
    anchors = np.empty((seq_len, S, 2), dtype=np.int32)
    anchors[:, :, 0] = np.arange(seq_len)[:, None]
    anchors[:, :, 1] = np.asarray(scales, dtype=np.int32)[None, :]
    flat = anchors.reshape(-1, 2)
    loc = np.zeros((seq_len * S, 2))
    cls = np.zeros(seq_len * S, dtype=np.int32)
    for tgt in np.asarray(targets_cw):
        pos = np.nonzero(_hull_iou(flat, tgt) > iou_thresh)[0]
        cls[pos] = 1
        aw = flat[pos, 1]
        loc[pos, 0] = (tgt[0] - flat[pos, 0]) / aw
        loc[pos, 1] = np.log(tgt[1] / aw)
    return cls.reshape(seq_len, S), loc.reshape(seq_len, S, 2)


def sample_negatives(cls_label: np.ndarray, num_neg: int, rng: np.random.Generator) -> np.ndarray:
    """Mark `num_neg` random non-positive anchors with -1 (the reference uses the global np.random state)."""
    shape = cls_label.shape
    out = cls_label.copy().reshape(-1)
    out[out < 0] = 0
    idx = np.nonzero(out == 0)[0]
    rng.shuffle(idx)
    out[idx[:num_neg]] = -1
    return out.reshape(shape)


def anchor_labels(target_mask: np.ndarray, scales: Sequence[int], rng: np.random.Generator,
                  pos_iou_thresh: float = 0.6, neg_iou_thresh: float = 0.0, incomplete_iou_thresh: float = 0.3,
                  neg_sample_ratio: float = 2.0, incomplete_sample_ratio: float = 1.0
                  ) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """Label recipe of anchor_based/train.py:86-108 for one video; None when the mask is empty (the loop skips it)."""
    target_mask = np.asarray(target_mask, dtype=bool)
    if not target_mask.any():
        return None
    T = target_mask.size
    tg = mask_to_segments(target_mask)
    cls, loc = positive_labels(T, scales, tg, pos_iou_thresh)
    num_pos = int(cls.sum())
    neg, _ = positive_labels(T, scales, tg, neg_iou_thresh)
    neg = sample_negatives(neg, int(neg_sample_ratio * num_pos), rng)
    inc, _ = positive_labels(T, scales, tg, incomplete_iou_thresh)
    inc[neg != 1] = 1
    inc = sample_negatives(inc, int(incomplete_sample_ratio * num_pos), rng)
    cls[neg == -1] = -1
    cls[inc == -1] = -1
    return cls, loc


class LabelCache:
    """The deterministic part of `anchor_labels`, computed once per video (precompute / cache).

    The reference redoes three IoU sweeps over all T x S anchors for every video of every epoch
    (anchor_based/train.py:86-108) although they only depend on the ground-truth mask; what changes from step to step is
    the random choice of negatives.  `labels(key, rng)` consumes the random stream exactly like `anchor_labels` does
    (two shuffles, same index arrays), so both return identical labels for the same generator state."""

    def __init__(self, scales: Sequence[int], pos_iou_thresh: float = 0.6, neg_iou_thresh: float = 0.0,
                 incomplete_iou_thresh: float = 0.3, neg_sample_ratio: float = 2.0,
                 incomplete_sample_ratio: float = 1.0):
        self.scales = [int(s) for s in scales]
        self.thresholds = (pos_iou_thresh, neg_iou_thresh, incomplete_iou_thresh)
        self.ratios = (neg_sample_ratio, incomplete_sample_ratio)
        self._store = {}

    def add(self, key, target_mask: np.ndarray) -> bool:
        """False (nothing stored) for an empty mask: the training loop skips such videos."""
        target_mask = np.asarray(target_mask, dtype=bool)
        if not target_mask.any():
            self._store[key] = None
            return False
        T = target_mask.size
        tg = mask_to_segments(target_mask)
        cls, loc = positive_labels(T, self.scales, tg, self.thresholds[0])
        neg, _ = positive_labels(T, self.scales, tg, self.thresholds[1])
        inc, _ = positive_labels(T, self.scales, tg, self.thresholds[2])
        self._store[key] = (cls, loc, neg, inc, int(cls.sum()))
        return True

    def __contains__(self, key) -> bool:
        return key in self._store

    def labels(self, key, rng: np.random.Generator) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        entry = self._store[key]
        if entry is None:
            return None
        cls, loc, neg, inc, num_pos = entry
        neg = sample_negatives(neg, int(self.ratios[0] * num_pos), rng)
        inc = inc.copy()
        inc[neg != 1] = 1
        inc = sample_negatives(inc, int(self.ratios[1] * num_pos), rng)
        out = cls.copy()
        out[neg == -1] = -1
        out[inc == -1] = -1
        return out, loc


# ----------------------------------------------------------------------------------------------- losses (torch)
def cls_loss(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """0.5 * (mean over positives of -log p + mean over negatives of -log(1 - p)); label in {1, -1, 0 = ignored}."""
    pred, label = pred.reshape(-1), label.reshape(-1)
    pos, neg = pred[label == 1], pred[label == -1]
    return 0.5 * (-(pos.log()).mean() - ((1 - neg).log()).mean())


def loc_loss(pred_loc: torch.Tensor, loc_label: torch.Tensor, cls_label: torch.Tensor, use_smooth: bool = True
             ) -> torch.Tensor:
    """Smooth-L1 (or L1) over the two offsets of the positive anchors only."""
    pos = cls_label == 1
    p, t = pred_loc[pos], loc_label[pos]
    return F.smooth_l1_loss(p, t) if use_smooth else (p - t).abs().mean()


# ----------------------------------------------------------------------------------------------- data parallel step
def allreduce_gradients(params: Sequence[torch.Tensor], world_size: int, group=None) -> int:
    """Average the gradients of `params` over the ranks with ONE collective on one flat fp32 bucket (2.25 M values =
    9 MB for this model: latency bound, so a single bucket).  Parameters without a gradient contribute zeros.
    The shared fc block appears once (its gradient is already the sum over its fc_depth uses).  Returns the number of
    reduced elements."""
    import torch.distributed as dist
    seen, uniq = set(), []
    for p in params:
        if p.requires_grad and id(p) not in seen:
            seen.add(id(p))
            uniq.append(p)
    if not uniq:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in uniq])
    if world_size > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= world_size
    o = 0
    for p in uniq:
        n = p.numel()
        g = flat[o:o + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        o += n
    return o


class DataParallelStep:
    """One optimiser step over the videos of every rank: loss = mean over the local videos of
    cls_loss + lambda_reg * loc_loss (anchor_based/train.py:119-123), backward, flat all-reduce (mean over ranks),
    identical Adam update on every rank (lr 5e-5, weight decay 1e-5: anchor_based/train.py:53-55)."""

    def __init__(self, model, lr: float = 5e-5, weight_decay: float = 1e-5, lambda_reg: float = 1.0,
                 world_size: int = 1, group=None):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay)
        self.lambda_reg = lambda_reg
        self.world_size = world_size
        self.group = group

    def loss(self, seqs: List[torch.Tensor], cls_labels: List[torch.Tensor], loc_labels: List[torch.Tensor]):
        lengths = [int(s.shape[0]) for s in seqs]
        pred_cls, pred_loc = self.model.forward_packed(torch.cat(seqs), lengths)
        total, o = 0.0, 0
        for t, cl, ll in zip(lengths, cls_labels, loc_labels):
            total = total + cls_loss(pred_cls[o:o + t], cl) + self.lambda_reg * loc_loss(pred_loc[o:o + t], ll, cl)
            o += t
        return total / len(seqs)

    def step(self, seqs, cls_labels, loc_labels) -> float:
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss(seqs, cls_labels, loc_labels)
        loss.backward()
        allreduce_gradients(self.params, self.world_size, self.group)
        self.optimizer.step()
        return float(loss.detach())
