"""CPU oracle for the EDSNet anchor-based scoring hot path.

TEST INFRASTRUCTURE ONLY.  This module is a from-scratch CPU restatement of
the reference algorithm (torch-CPU ops for the floating-point tensor part,
NumPy for decode / NMS / summary).  It is the checker for the CUDA path and
the ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``.  Nothing in
the product package (``edsnet_b200``) imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` may.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference/src`` (with import stubs for absent
third-party modules), runs it on seeded inputs and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here
against those vectors and against the known-answer literals in the reference's
own unit tests.  The one unpinned piece is the knapsack tie-break (ortools is
absent; see ``knapsack_dp``).

All ``file:line`` citations are relative to ``/root/reference/src``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

HEADS = 8          # modules/models.py:135 (num_head comes from the ctor, 8 on every call site)
DIM_HEAD = 64      # modules/models.py:135
LANDMARKS = 64     # modules/models.py:135
PINV_ITERS = 6     # modules/models.py:135
CONV_TAPS = 33     # modules/models.py:135


# --------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------

# Reference state-dict layout (anchor_based/dsnet.py:66-98 + transformer/nystroformer.py:52-65).
PARAM_SHAPES = {
    "base_model.to_qkv.weight": lambda F, H, nh: (3 * nh * DIM_HEAD, F),
    "base_model.to_out.0.weight": lambda F, H, nh: (F, nh * DIM_HEAD),
    "base_model.to_out.0.bias": lambda F, H, nh: (F,),
    "base_model.res_conv.weight": lambda F, H, nh: (nh, 1, CONV_TAPS, 1),
    "layer_norm.weight": lambda F, H, nh: (F,),
    "layer_norm.bias": lambda F, H, nh: (F,),
    "fc1.weight": lambda F, H, nh: (H, F),
    "fc1.bias": lambda F, H, nh: (H,),
    "fc_block.0.weight": lambda F, H, nh: (H, H),
    "fc_block.0.bias": lambda F, H, nh: (H,),
    "fc_block.3.weight": lambda F, H, nh: (H,),
    "fc_block.3.bias": lambda F, H, nh: (H,),
    "fc_cls.0.weight": lambda F, H, nh: (1, H),
    "fc_cls.0.bias": lambda F, H, nh: (1,),
    "fc_loc.0.weight": lambda F, H, nh: (2, H),
    "fc_loc.0.bias": lambda F, H, nh: (2,),
}


def synth_features(T: int, seed: int, num_feature: int = 1024) -> torch.Tensor:
    """Synthetic stand-in for L2-normalised GoogLeNet pool5 rows
    (helpers/video_helper.py:72): non-negative, unit L2 norm per frame."""
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(T, num_feature, generator=g))
    x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return x.contiguous()


def synth_params(seed: int, init: str = "default", num_feature: int = 1024,
                 num_hidden: int = 128, num_head: int = HEADS) -> Dict[str, torch.Tensor]:
    """Random weights with the reference state-dict names.

    ``default`` mimics the scale of torch's module defaults (U(-1/sqrt(fan_in), .));
    ``xavier`` mimics anchor_based/train.py:19-24 (xavier-uniform gain sqrt(2),
    bias 0.1, LayerNorm untouched).  These are *synthetic* weights: parity runs
    feed the same tensors to the reference / oracle / CUDA path.
    """
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shp in PARAM_SHAPES.items():
        shape = shp(num_feature, num_hidden, num_head)
        is_ln = name.startswith("layer_norm") or name.startswith("fc_block.3")
        if is_ln:
            if init == "default":
                t = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            else:  # perturbed affine so LN weight/bias are actually exercised
                t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith("weight") \
                    else 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            if init == "xavier":
                t = torch.full(shape, 0.1)
            else:
                fan_in = {"base_model.to_out.0.bias": num_head * DIM_HEAD, "fc1.bias": num_feature}.get(name, num_hidden)
                b = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * b
        else:
            if name == "base_model.res_conv.weight":     # (heads, 1, 33, 1), groups = heads
                fan_in, fan_out = CONV_TAPS, num_head * CONV_TAPS
            else:
                fan_out, fan_in = shape[0], shape[1]
            if init == "xavier":
                b = math.sqrt(2.0) * math.sqrt(6.0 / (fan_in + fan_out))
            else:
                b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        out[name] = t.to(torch.float32).contiguous()
    return out


# --------------------------------------------------------------------------
# Nystrom landmark attention  (transformer/nystroformer.py)
# --------------------------------------------------------------------------

def pinv_iterative(a: torch.Tensor, iters: int = PINV_ITERS) -> torch.Tensor:
    """Iterative Moore-Penrose pseudo-inverse of a stack of square matrices.

    Follows transformer/nystroformer.py:13-28: the start value is a^T divided
    by ONE scalar -- (largest absolute row sum) x (largest absolute column sum),
    both maxima taken over the whole stack (all heads) -- then ``iters`` rounds
    of the cubic Newton-Schulz style update
    z <- z/4 * (13 I - az (15 I - az (7 I - az))).
    """
    mag = a.abs()
    scale = mag.sum(dim=-1).max() * mag.sum(dim=-2).max()
    z = a.transpose(-1, -2) / scale
    eye = torch.eye(a.shape[-1], dtype=a.dtype)
    for _ in range(iters):
        az = a @ z
        z = 0.25 * z @ (13 * eye - az @ (15 * eye - az @ (7 * eye - az)))
    return z


def nystrom_attention(x: torch.Tensor, p: Dict[str, torch.Tensor], heads: int = HEADS,
                      stages: dict | None = None) -> torch.Tensor:
    """One video through the landmark attention block. x: (T, F) -> (T, F).

    Follows transformer/nystroformer.py:67-150 with mask=None, dropout 0:
    front zero-padding to a multiple of 64 rows (:72-75; the pad rows are real
    zero tokens), bias-free qkv projection (:82), q scaled by 1/8 (:91),
    landmarks = means of n/64 consecutive rows (:95-111), three softmax
    kernels (:115-130), pseudo-inverse of the landmark kernel (:131), value
    aggregation (:133), depth-wise 33-tap value convolution residual
    (:137-138), head merge + output projection, last T rows kept (:142-144).
    """
    T, F = x.shape
    d = DIM_HEAD
    m = LANDMARKS
    pad = (m - T % m) % m
    n = T + pad
    seg = n // m
    xp = torch.cat([x.new_zeros(pad, F), x], dim=0) if pad else x
    qkv = xp @ p["base_model.to_qkv.weight"].t()                      # (n, 3*h*d)
    q, k, v = (t.reshape(n, heads, d).permute(1, 0, 2) for t in qkv.chunk(3, dim=-1))   # (h, n, d)
    q = q * (d ** -0.5)
    ql = q.reshape(heads, m, seg, d).sum(dim=2) / seg                  # (h, m, d)
    kl = k.reshape(heads, m, seg, d).sum(dim=2) / seg
    a1 = torch.softmax(q @ kl.transpose(1, 2), dim=-1)                 # (h, n, m)
    a2 = torch.softmax(ql @ kl.transpose(1, 2), dim=-1)                # (h, m, m)
    a3 = torch.softmax(ql @ k.transpose(1, 2), dim=-1)                 # (h, m, n)
    z = pinv_iterative(a2)
    o = (a1 @ z) @ (a3 @ v)                                            # (h, n, d)
    # depth-wise FIR along time: same 33 taps for all 64 channels of a head
    w = p["base_model.res_conv.weight"].reshape(heads, 1, CONV_TAPS, 1)
    o = o + torch.nn.functional.conv2d(v.unsqueeze(0), w, padding=(CONV_TAPS // 2, 0), groups=heads)[0]
    merged = o.permute(1, 0, 2).reshape(n, heads * d)
    y = merged @ p["base_model.to_out.0.weight"].t() + p["base_model.to_out.0.bias"]
    if stages is not None:
        stages.update(qkv=qkv, q_land=ql, k_land=kl, attn2=a2, pinv=z, a3v=a3 @ v, merged=merged, attn_out=y[pad:])
    return y[pad:]


# --------------------------------------------------------------------------
# Full multi-head attention base (config 4 only)  (modules/models.py:12-74)
# --------------------------------------------------------------------------

MHA_PARAM_SHAPES = {
    "base_model.Q.weight": (1024, 1024), "base_model.K.weight": (1024, 1024), "base_model.V.weight": (1024, 1024),
    "base_model.fc.0.weight": (1024, 1024),
}


def synth_params_mha(seed: int, init: str = "default") -> Dict[str, torch.Tensor]:
    """Weights of DSNet(base_model='attention'): the shared tail of ``synth_params`` plus the four bias-free
    1024 x 1024 projections of modules/models.py:33-44."""
    p = {k: v for k, v in synth_params(seed, init).items() if not k.startswith("base_model.")}
    g = torch.Generator().manual_seed(seed + 7919)
    for name, (fo, fi) in MHA_PARAM_SHAPES.items():
        b = math.sqrt(2.0) * math.sqrt(6.0 / (fi + fo)) if init == "xavier" else 1.0 / math.sqrt(fi)
        p[name] = ((torch.rand((fo, fi), generator=g) * 2 - 1) * b).float().contiguous()
    return p


def mha_attention(x: torch.Tensor, p: Dict[str, torch.Tensor], heads: int = HEADS) -> torch.Tensor:
    """x: (T, F) -> (T, F).  modules/models.py:46-65 in eval mode (both Dropout(0.5) are the identity):
    Q/K/V = x W^T, heads split along the feature axis (d_k = F / heads), softmax(Q K^T / sqrt(d_k)) V,
    heads merged, bias-free output projection."""
    T, Fdim = x.shape
    dk = Fdim // heads
    q, k, v = (x @ p[f"base_model.{n}.weight"].t() for n in ("Q", "K", "V"))
    q, k, v = (t.reshape(T, heads, dk).permute(1, 0, 2) for t in (q, k, v))          # (h, T, dk)
    attn = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(dk), dim=-1)
    y = (attn @ v).permute(1, 0, 2).reshape(T, Fdim)
    return y @ p["base_model.fc.0.weight"].t()


# --------------------------------------------------------------------------
# DSNet forward (anchor_based/dsnet.py:100-115)
# --------------------------------------------------------------------------

def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def roi_pool(u: torch.Tensor, scales: Sequence[int]) -> torch.Tensor:
    """(T, H) -> (T, S, H).  anchor_based/dsnet.py:78-80,111-113: AvgPool1d(s, 1, s//2)
    with zero padding counted in the divisor, last of the T+1 outputs dropped, i.e.
    pooled[t] = (1/s) * sum_{j=t-s/2}^{t+s/2-1} u[j] with u[j] = 0 outside [0, T)."""
    T, H = u.shape
    outs = []
    csum = torch.cat([u.new_zeros(1, H, dtype=torch.float64), u.double().cumsum(0)], 0)
    for s in scales:
        if s % 2:
            raise RuntimeError("odd anchor scale: the reference's .view() fails for odd scales")
        lo = (torch.arange(T) - s // 2).clamp(0, T)
        hi = (torch.arange(T) + s // 2).clamp(0, T)
        outs.append(((csum[hi] - csum[lo]) / s).to(u.dtype))
    return torch.stack(outs, dim=1)


def roi_pool_direct(u: torch.Tensor, scales: Sequence[int]) -> torch.Tensor:
    """Same quantity as ``roi_pool`` via torch's own avg_pool1d (the op the
    reference calls), used to pin ``roi_pool``."""
    outs = []
    ut = u.t().unsqueeze(0)
    for s in scales:
        o = torch.nn.functional.avg_pool1d(ut, s, stride=1, padding=s // 2)[0].t()[:-1]
        outs.append(o)
    return torch.stack(outs, dim=1)


def dsnet_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], scales: Sequence[int],
                  fc_depth: int = 5, heads: int = HEADS, stages: dict | None = None, base: str = "nystromformer",
                  keep: torch.Tensor | None = None, dropout_gen: torch.Generator | None = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x: (T, F) float -> pred_cls (T, S), pred_loc (T, S, 2).  anchor_based/dsnet.py:100-115.
    Eval mode by default (Dropout is the identity).  Train mode (dsnet.py:91-95: Dropout(0.5) after the ReLU of every
    application of the shared block): pass ``keep`` = bool (fc_depth, T, H) mask of the kept activations (what the CUDA
    path draws with Philox, so both sides can use the SAME mask) or ``dropout_gen`` to draw one with torch."""
    if base == "attention":
        y = mha_attention(x, p, heads) + x
    else:
        y = nystrom_attention(x, p, heads, stages) + x
    u = layer_norm(y, p["layer_norm.weight"], p["layer_norm.bias"])
    u = u @ p["fc1.weight"].t() + p["fc1.bias"]
    for i in range(fc_depth):          # ONE shared block applied fc_depth times (dsnet.py:91-96)
        u = torch.relu(u @ p["fc_block.0.weight"].t() + p["fc_block.0.bias"])
        if keep is not None:
            u = u * keep[i].to(u.dtype) * 2.0
        elif dropout_gen is not None:
            u = u * (torch.rand(u.shape, generator=dropout_gen) < 0.5).to(u.dtype) * 2.0
        u = layer_norm(u, p["fc_block.3.weight"], p["fc_block.3.bias"])
    pooled = roi_pool_direct(u, scales)
    cls = torch.sigmoid(pooled @ p["fc_cls.0.weight"].t() + p["fc_cls.0.bias"]).reshape(x.shape[0], len(scales))
    loc = (pooled @ p["fc_loc.0.weight"].t() + p["fc_loc.0.bias"]).reshape(x.shape[0], len(scales), 2)
    if stages is not None:
        stages.update(hidden=u, pooled=pooled)
    return cls, loc


# --------------------------------------------------------------------------
# predict / decode / NMS  (NumPy semantics)
# --------------------------------------------------------------------------

def anchor_grid(T: int, scales: Sequence[int]) -> np.ndarray:
    """(T, S, 2) int32 [position, scale]; anchor_based/anchor_helper.py:8-19."""
    g = np.empty((T, len(scales), 2), dtype=np.int32)
    g[:, :, 0] = np.arange(T, dtype=np.int32)[:, None]
    g[:, :, 1] = np.asarray(scales, dtype=np.int32)[None, :]
    return g


def exp_f32(x: np.ndarray, mode: str = "numpy") -> np.ndarray:
    """float32 exp.  ``numpy``: whatever ``np.exp`` does on a float32 array on THIS machine -- what the reference
    executes (anchor_helper.py:88), but NumPy's SIMD float32 exp is not correctly rounded (<= 2.5 ulp; it differs
    from the correctly rounded value for ~39 % of N(0, 0.7) arguments on AVX-512) and is dispatch-dependent.
    ``cr``: correctly rounded (float64 exp, one rounding) -- machine-independent; this is what the CUDA decode
    kernel computes and what GPU parity tests compare bit for bit."""
    x = np.asarray(x, dtype=np.float32)
    if mode == "numpy":
        return np.exp(x)
    if mode == "cr":
        return np.exp(x.astype(np.float64)).astype(np.float32)
    raise ValueError(mode)


def decode_boxes(pred_loc: np.ndarray, T: int, scales: Sequence[int], exp_mode: str = "numpy") -> np.ndarray:
    """Offsets -> float32 left/right boxes, (T*S, 2).

    anchor_based/anchor_helper.py:74-93 then helpers/bbox_helper.py:21-31:
    centre = oc * aw + ac and width = exp(ow) * aw are float32 x int32 products,
    which NumPy evaluates in float64; the float32 exp comes first.  The CW pair
    is then cast to float32 and left/right = c -/+ w/2 in float32.
    """
    off = np.asarray(pred_loc, dtype=np.float32).reshape(-1, 2)
    anc = anchor_grid(T, scales).reshape(-1, 2)
    aw = anc[:, 1].astype(np.float64)
    centre = off[:, 0].astype(np.float64) * aw + anc[:, 0].astype(np.float64)
    width = exp_f32(off[:, 1], exp_mode).astype(np.float64) * aw
    c32 = centre.astype(np.float32)
    w32 = width.astype(np.float32)
    half = w32 / np.float32(2)
    return np.stack([c32 - half, c32 + half], axis=1)


def clip_round(boxes: np.ndarray, T: int) -> np.ndarray:
    """evaluate.py:26 -- clip to [0, T], round half to even, int32."""
    return np.rint(np.clip(boxes, 0, T)).astype(np.int32)


def score_order(scores: np.ndarray) -> np.ndarray:
    """Descending-score visiting order.  helpers/bbox_helper.py:95 uses
    ``argsort()[::-1]`` with NumPy's default (unstable) sort, so the order of
    exactly-equal scores is implementation-defined in the reference.  The oracle
    and the CUDA kernel fix it: stable ascending sort, reversed => among equal
    scores the HIGHER original index is visited first."""
    return np.argsort(scores, kind="stable")[::-1]


def nms_1d(scores: np.ndarray, boxes: np.ndarray, thresh: float
           ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Greedy temporal NMS; helpers/bbox_helper.py:80-118 with iou_lr :49-70.

    Returns (keep_scores, keep_boxes, keep_idx) in visiting order; keep_idx
    indexes the ORIGINAL (unfiltered) arrays.  Overlap measure is
    intersection / hull length, survivors need overlap < thresh (strict).
    """
    scores = np.asarray(scores)
    boxes = np.asarray(boxes)
    valid = np.nonzero(boxes[:, 0] < boxes[:, 1])[0]
    order = valid[score_order(scores[valid])]
    lo = boxes[order, 0].astype(np.float64)
    hi = boxes[order, 1].astype(np.float64)
    alive = np.ones(len(order), dtype=bool)
    keep: List[int] = []
    for i in range(len(order)):
        if not alive[i]:
            continue
        keep.append(i)
        inter = np.minimum(hi, hi[i]) - np.maximum(lo, lo[i])
        inter[inter < 0] = 0
        hull = np.maximum(hi, hi[i]) - np.minimum(lo, lo[i])
        hull[hull <= 0] = 1e-6
        alive &= (inter / hull) < thresh
    kidx = order[np.asarray(keep, dtype=np.int64)] if keep else np.zeros(0, dtype=np.int64)
    return scores[kidx], boxes[kidx], kidx


def predict(x: torch.Tensor, p: Dict[str, torch.Tensor], scales: Sequence[int], fc_depth: int = 5
            ) -> Tuple[np.ndarray, np.ndarray]:
    """anchor_based/dsnet.py:140-153: flat scores (T*S,) and float32 LR boxes (T*S, 2)."""
    with torch.no_grad():
        cls, loc = dsnet_forward(x, p, scales, fc_depth)
    return cls.numpy().reshape(-1), decode_boxes(loc.numpy(), x.shape[0], scales)


def proposals(x: torch.Tensor, p: Dict[str, torch.Tensor], scales: Sequence[int], fc_depth: int = 5,
              nms_thresh: float = 0.5):
    """evaluate.py:24-28: predict -> clip/round -> NMS."""
    s, b = predict(x, p, scales, fc_depth)
    return nms_1d(s, clip_round(b, x.shape[0]), nms_thresh)


# --------------------------------------------------------------------------
# summary (parity criterion, stays on host; helpers/vsumm_helper.py)
# --------------------------------------------------------------------------

def knapsack_dp(values: Sequence[int], weights: Sequence[int], capacity: int) -> List[int]:
    """Exact 0/1 knapsack by DP over capacity.

    The reference calls ortools' KNAPSACK_MULTIDIMENSION_BRANCH_AND_BOUND_SOLVER
    (helpers/vsumm_helper.py:26-45; ortools unpinned in requirements.txt:3,
    9.10.4067 in the notebook log training_weight:36).  ortools is absent here;
    this returns AN optimal item set (max total value, weight <= capacity).  The
    optimum value is unique; the choice among equal-value optima is PARITY
    UNPINNED.  Tie rule here: scan items last to first, take an item only when
    leaving it out would lose value.
    """
    values = [int(v) for v in values]
    weights = [int(w) for w in weights]
    capacity = int(capacity)
    n = len(values)
    if n == 0 or capacity <= 0:
        return [i for i in range(n) if weights[i] <= 0 < values[i]] if capacity >= 0 else []
    g = math.gcd(capacity, *[w for w in weights if w > 0]) if any(w > 0 for w in weights) else 1
    cap = capacity // g
    wts = [w // g for w in weights]
    best = np.zeros((n + 1, cap + 1), dtype=np.int64)
    for i in range(n):
        w, v = wts[i], values[i]
        best[i + 1] = best[i]
        if w <= cap:
            cand = best[i, : cap + 1 - w] + v
            np.maximum(best[i + 1, w:], cand, out=best[i + 1, w:])
    chosen, c = [], cap
    for i in range(n - 1, -1, -1):
        if best[i + 1, c] != best[i, c]:
            chosen.append(i)
            c -= wts[i]
    return sorted(chosen)


def keyshot_summary(pred: np.ndarray, cps: np.ndarray, n_frames: int, nfps: np.ndarray,
                    picks: np.ndarray, proportion: float = 0.15) -> np.ndarray:
    """helpers/vsumm_helper.py:53-98: upsample scores by picks, per-shot score =
    int(1000 * float32 mean), knapsack under int(n_frames * proportion) frames."""
    picks = np.asarray(picks, dtype=np.int32)
    n_frames = int(n_frames)
    frame_scores = np.zeros(n_frames, dtype=np.float32)
    bounds = list(picks) + [n_frames]
    for i in range(len(picks)):
        frame_scores[bounds[i]:bounds[i + 1]] = pred[i]
    seg = np.zeros(len(cps), dtype=np.int32)
    for j, (first, last) in enumerate(cps):
        seg[j] = int(1000 * frame_scores[first:last + 1].mean())
    picked = knapsack_dp(seg, nfps, int(n_frames * proportion))
    summ = np.zeros(n_frames, dtype=bool)
    for j in picked:
        summ[cps[j][0]:cps[j][1] + 1] = True
    return summ


def bbox_summary(T: int, keep_scores: np.ndarray, keep_boxes: np.ndarray, cps, n_frames, nfps, picks):
    """helpers/vsumm_helper.py:101-116: per-position score = max over kept boxes covering it."""
    score = np.zeros(T, dtype=np.float32)
    for s, (lo, hi) in zip(keep_scores, keep_boxes):
        score[lo:hi] = np.maximum(score[lo:hi], s)
    return keyshot_summary(score, cps, n_frames, nfps, picks)


# --------------------------------------------------------------------------
# evaluation metrics (evaluate.py:31-37; helpers/vsumm_helper.py:8-23, 48-50, 119-172)
# --------------------------------------------------------------------------

def f1_score(pred: np.ndarray, test: np.ndarray) -> float:
    """helpers/vsumm_helper.py:8-23: integer counts, float64 precision / recall / harmonic mean."""
    pred = np.asarray(pred, dtype=bool)
    test = np.asarray(test, dtype=bool)
    overlap = (pred & test).sum()
    if overlap == 0:
        return 0.0
    precision = overlap / pred.sum()
    recall = overlap / test.sum()
    return float(2 * precision * recall / (precision + recall))


def summ_f1score(pred_summ: np.ndarray, test_summ: np.ndarray, eval_metric: str = "avg") -> float:
    """helpers/vsumm_helper.py:142-172: the prediction is cut / zero-padded to the users' frame count, then the
    mean ('avg', TVSum) or the maximum ('max', SumMe) of the per-user F1 scores."""
    pred_summ = np.asarray(pred_summ, dtype=bool)
    test_summ = np.asarray(test_summ, dtype=bool)
    n = test_summ.shape[1]
    if pred_summ.size > n:
        pred_summ = pred_summ[:n]
    elif pred_summ.size < n:
        pred_summ = np.pad(pred_summ, (0, n - pred_summ.size))
    f1s = [f1_score(u, pred_summ) for u in test_summ]
    if eval_metric == "avg":
        return float(np.mean(f1s))
    if eval_metric == "max":
        return float(np.max(f1s))
    raise ValueError(f"Invalid eval metric {eval_metric}")


def downsample_summ(summ: np.ndarray) -> np.ndarray:
    """helpers/vsumm_helper.py:48-50."""
    return summ[::15]


def summ_diversity(pred_summ: np.ndarray, features: np.ndarray) -> float:
    """helpers/vsumm_helper.py:119-139: mean over ordered pairs i != j of the selected rows of f_i . f_j."""
    assert len(pred_summ) == len(features)
    pos = np.asarray(features)[np.asarray(pred_summ, dtype=bool)]
    if len(pos) < 2:
        return 0.0
    div = 0.0
    for f in pos:
        div += (f * pos).sum() - (f * f).sum()
    return float(div / (len(pos) * (len(pos) - 1)))


# --------------------------------------------------------------------------
# kernel temporal segmentation (the shot boundaries infer.py feeds the path with):
# kts/cpd_nonlin.py:4-92, kts/cpd_auto.py:6-33, helpers/video_helper.py:109-126
# --------------------------------------------------------------------------

def kts_scatters(K: np.ndarray) -> np.ndarray:
    """kts/cpd_nonlin.py:4-27.  J[i, j] = within-segment scatter of frames i..j (0 below the diagonal).  The 2-D
    prefix sums run in K's own dtype (float32 for a float32 kernel matrix, sequential along each axis), everything
    else in float64, in the reference's operation order."""
    n = K.shape[0]
    k1 = np.cumsum(np.concatenate([[0.0], np.diag(K).astype(np.float64)]))
    k2 = np.zeros((n + 1, n + 1))
    k2[1:, 1:] = np.cumsum(np.cumsum(K, axis=0), axis=1)
    d2 = np.diag(k2)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    length = (jj - ii + 1).astype(np.float32) + (jj == ii - 1).astype(np.float32)
    cross = ((d2[1:][None, :] + d2[:-1][:, None]) - k2[1:, :-1].T) - k2[:-1, 1:]
    J = (k1[1:][None, :] - k1[:-1][:, None]) - cross / length
    J[jj < ii] = 0
    return J


def kts_dp(K: np.ndarray, m: int, lmin: int = 1, lmax: int = 100000):
    """kts/cpd_nonlin.py:30-92 with back-tracking: change points of the best segmentation into m+1 pieces and the
    objective for 0..m change points.  One vectorised row update per k (the reference loops over l in Python)."""
    m = int(m)
    n = K.shape[0]
    assert (m + 1) * lmin <= n <= (m + 1) * lmax and 1 <= lmin <= lmax
    J = kts_scatters(K)
    BIG = 1e101
    I = np.full((m + 1, n + 1), BIG)
    hi = min(lmax, n + 1)
    I[0, lmin:hi] = J[0, lmin - 1:hi - 1]
    prev = np.zeros((m + 1, n + 1), dtype=np.int64)
    for k in range(1, m + 1):
        for l in range((k + 1) * lmin, n + 1):
            t0, t1 = max(k * lmin, l - lmax), l - lmin + 1
            c = J[t0:t1, l - 1] + I[k - 1, t0:t1]
            a = int(np.argmin(c))
            I[k, l] = c[a]
            prev[k, l] = a + t0
    cps = np.zeros(m, dtype=np.int64)
    cur = n
    for k in range(m, 0, -1):
        cps[k - 1] = prev[k, cur]
        cur = cps[k - 1]
    scores = I[:, n].copy()
    scores[scores > 1e99] = np.inf
    return cps, scores


def kts_auto(K: np.ndarray, ncp: int, vmax: float, desc_rate: int = 1, lmin: int = 1, lmax: int = 100000):
    """kts/cpd_auto.py:6-33: number of change points chosen by the penalised objective, then the segmentation."""
    m = int(ncp)
    _, scores = kts_dp(K, m, lmin, lmax)
    N = K.shape[0]
    N2 = N * desc_rate
    pen = np.zeros(m + 1)
    q = np.arange(1, m + 1)
    pen[1:] = (vmax * q / (2.0 * N2)) * (np.log(float(N2) / q) + 1)
    costs = scores / float(N) + pen
    m_best = int(np.argmin(costs))
    return kts_dp(K, m_best, lmin, lmax)


def kts_shots(n_frames: int, features: np.ndarray, sample_rate: int = 15):
    """helpers/video_helper.py:109-126: (change_points [n_seg, 2] inclusive, frames per segment, picks)."""
    T = len(features)
    picks = np.arange(0, T) * sample_rate
    K = np.matmul(features, features.T)
    cps, _ = kts_auto(K, T - 1, 1)
    cps = cps * sample_rate
    cps = np.hstack((0, cps, n_frames))
    begin, end = cps[:-1], cps[1:]
    return np.vstack((begin, end - 1)).T, end - begin, picks


# --------------------------------------------------------------------------
# error metrics used by every parity test
# --------------------------------------------------------------------------

def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rel_max(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
