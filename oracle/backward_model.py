"""Hand-derived backward of the anchor-based scoring model, staged exactly like the CUDA training kernels.

TEST INFRASTRUCTURE ONLY (same rules as dsnet_oracle.py: imported by tests/, never by the product package).

Two things live here:
  * ``dsnet_forward_saved`` / ``dsnet_backward_staged``: the forward with every activation the kernels keep, and the
    backward written as the SAME sequence of matrix products, softmax / LayerNorm back-substitutions and reductions the
    kernels in csrc/train.cuh execute (one function per kernel), in torch-CPU ops.  ``tests/test_backward_model.py``
    pins it against torch.autograd of the oracle forward in float64 (<= 1e-10), so a GPU stage test can compare any
    intermediate buffer (d logits, d u0, d y, d merged, d qkv, ...) with a value that is known to be right.
  * ``reference_losses``: anchor_based/losses.py:5-57 restated (mean over positives / negatives, smooth-L1 over the
    positive anchors' two offsets) with the closed-form gradient with respect to the LOGITS that csrc/train.cuh uses.

All ``file:line`` citations are relative to /root/reference/src.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import dsnet_oracle as orc

H, D, M = orc.HEADS, orc.DIM_HEAD, orc.LANDMARKS


# --------------------------------------------------------------------------------------------------------------------
# dropout mask: Philox4x32-10, counter = (row, layer, offset_lo, offset_hi), key = (seed_lo, seed_hi); bit c of the 128
# output bits decides column c (keep-probability 1/2).  Restated from csrc/train.cuh: philox4x32_10 / dropout_words.
# --------------------------------------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: Tuple[int, int]) -> np.ndarray:
    """ctr: uint32 [..., 4] -> uint32 [..., 4]."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0] & 0xFFFFFFFF), np.uint64(key[1] & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PH_M0) * c[0]
        p1 = np.uint64(_PH_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(_PH_W0)) & mask
        k1 = (k1 + np.uint64(_PH_W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def dropout_mask(rows: int, depth: int, seed: int, offset: int) -> np.ndarray:
    """bool [depth, rows, 128]: True = kept.  Row index = row of the packed batch."""
    ctr = np.zeros((depth, rows, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(rows, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.arange(depth, dtype=np.uint32)[:, None]
    ctr[..., 2] = np.uint32(offset & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((offset >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))              # [depth, rows, 4]
    bits = (w[..., :, None] >> np.arange(32, dtype=np.uint32)[None, None, None, :]) & np.uint32(1)
    return bits.reshape(depth, rows, 128).astype(bool)


# --------------------------------------------------------------------------------------------------------------------
# losses (anchor_based/losses.py) and their gradient with respect to the logits
# --------------------------------------------------------------------------------------------------------------------
def reference_losses(pred_cls: torch.Tensor, pred_loc: torch.Tensor, cls_label: torch.Tensor, loc_label: torch.Tensor,
                     lambda_reg: float = 1.0):
    """(loss, cls_loss, loc_loss) of one video exactly as anchor_based/train.py:119-123 combines losses.py:5-57."""
    pred, lab = pred_cls.reshape(-1), cls_label.reshape(-1)
    pos, neg = lab == 1, lab == -1
    cls = 0.5 * (-(pred[pos].log()).mean() - ((1 - pred[neg]).log()).mean())
    p, t = pred_loc[cls_label == 1], loc_label[cls_label == 1]
    loc = torch.nn.functional.smooth_l1_loss(p, t)
    return cls + lambda_reg * loc, cls, loc


def loss_grad_logits(pred_cls, pred_loc, cls_label, loc_label, lambda_reg=1.0, scale=1.0):
    """Closed form used by loss_grad_kernel: d loss / d logit (the value BEFORE the sigmoid) and d loss / d pred_loc.
       positives:  -0.5 (1 - p) / n_pos        negatives:  0.5 p / n_neg        (no division by p or 1 - p)
       offsets of positives: lambda clamp(p - t, -1, 1) / (2 n_pos)"""
    lab = cls_label
    pos, neg = lab == 1, lab == -1
    n_pos, n_neg = pos.sum(), neg.sum()
    dlogit = torch.zeros_like(pred_cls)
    dlogit[pos] = -0.5 * (1 - pred_cls[pos]) / n_pos
    dlogit[neg] = 0.5 * pred_cls[neg] / n_neg
    dloc = torch.zeros_like(pred_loc)
    d = pred_loc[pos] - loc_label[pos]
    dloc[pos] = lambda_reg * d.clamp(-1, 1) / (2 * n_pos)
    return dlogit * scale, dloc * scale


# --------------------------------------------------------------------------------------------------------------------
# forward with saved activations (one video)
# --------------------------------------------------------------------------------------------------------------------
def _ln_stats(x):
    mu = x.mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    return mu, rstd


def dsnet_forward_saved(x: torch.Tensor, p: Dict[str, torch.Tensor], scales: Sequence[int], fc_depth: int,
                        keep: Optional[torch.Tensor] = None) -> dict:
    """Forward of one video keeping what the training kernels keep.  keep: bool [fc_depth, T, 128] dropout mask (None =
    eval mode).  Returns a dict with pred_cls / pred_loc / logits and the saved tensors."""
    T, F = x.shape
    pad = (M - T % M) % M
    n = T + pad
    seg = n // M
    s = {"T": T, "pad": pad, "n": n, "seg": seg, "x": x}
    qkv = x @ p["base_model.to_qkv.weight"].t()                                  # real rows only (pad rows project to 0)
    q, k, v = (t.reshape(T, H, D).permute(1, 0, 2) for t in qkv.chunk(3, -1))    # (h, T, d)
    q = q * 0.125
    zer = x.new_zeros(H, pad, D)
    qp, kp = torch.cat([zer, q], 1), torch.cat([zer, k], 1)
    ql = qp.reshape(H, M, seg, D).sum(2) / seg
    kl = kp.reshape(H, M, seg, D).sum(2) / seg
    a1 = torch.softmax(q @ kl.transpose(1, 2), -1)                               # (h, T, m): real rows
    a2 = torch.softmax(ql @ kl.transpose(1, 2), -1)
    s3 = ql @ kp.transpose(1, 2)                                                 # pad keys take part (logit 0, value 0)
    m3 = s3.max(-1, keepdim=True).values
    l3 = torch.exp(s3 - m3).sum(-1, keepdim=True)
    a3 = torch.exp(s3[:, :, pad:] - m3) / l3                                     # (h, m, T): real keys
    b = a3 @ v                                                                   # a3v (h, m, d)
    col = a2.sum(-2)                                                             # |a2| = a2
    rmax, cmax = a2.sum(-1).max(), col.max()
    c = rmax * cmax
    zs = [a2.transpose(1, 2) / c]
    eye = torch.eye(M, dtype=x.dtype)
    for _ in range(orc.PINV_ITERS):
        z = zs[-1]
        pz = a2 @ z
        zs.append(0.25 * z @ (13 * eye - pz @ (15 * eye - pz @ (7 * eye - pz))))
    z = zs[-1]
    w = z @ b                                                                    # wmat
    o = a1 @ w
    taps = p["base_model.res_conv.weight"].reshape(H, orc.CONV_TAPS)
    vpad = torch.nn.functional.pad(v, (0, 0, 16, 16))
    conv = sum(taps[:, t, None, None] * vpad[:, t:t + T] for t in range(orc.CONV_TAPS))
    merged = (o + conv).permute(1, 0, 2).reshape(T, H * D)
    y = merged @ p["base_model.to_out.0.weight"].t() + p["base_model.to_out.0.bias"] + x
    mu, rstd = _ln_stats(y)
    yn = (y - mu) * rstd * p["layer_norm.weight"] + p["layer_norm.bias"]
    u = yn @ p["fc1.weight"].t() + p["fc1.bias"]
    uin, hs = [], []
    for l in range(fc_depth):
        uin.append(u)
        h = torch.relu(u @ p["fc_block.0.weight"].t() + p["fc_block.0.bias"])
        if keep is not None:
            h = h * keep[l].to(h.dtype) * 2.0
        hs.append(h)
        hm, hr = _ln_stats(h)
        u = (h - hm) * hr * p["fc_block.3.weight"] + p["fc_block.3.bias"]
    heads = torch.stack([u @ p["fc_cls.0.weight"][0], u @ p["fc_loc.0.weight"][0], u @ p["fc_loc.0.weight"][1]], 1)
    S = len(scales)
    csum = torch.cat([heads.new_zeros(1, 3), heads.cumsum(0)], 0)
    pooled = []
    for sc in scales:
        lo = (torch.arange(T) - sc // 2).clamp(0, T)
        hi = (torch.arange(T) + sc // 2).clamp(0, T)
        pooled.append((csum[hi] - csum[lo]) / sc)
    pooled = torch.stack(pooled, 1)                                              # (T, S, 3)
    logits = pooled[..., 0] + p["fc_cls.0.bias"]
    loc = pooled[..., 1:] + p["fc_loc.0.bias"]
    s.update(qkv=torch.cat([q.permute(1, 0, 2).reshape(T, H * D), qkv[:, H * D:]], 1),   # q part pre-scaled, as the kernels keep it
             q=q, k=k, v=v, ql=ql, kl=kl, a1=a1, a2=a2, a3=a3, m3=m3, l3=l3, b=b, c=c, rmax=rmax, cmax=cmax, col=col, zs=zs,
             w=w, merged=merged, y=y, yn=yn, u0=uin[0] if uin else u, uin=uin, hs=hs, uD=u, heads=heads, logits=logits,
             pred_cls=torch.sigmoid(logits), pred_loc=loc, scales=list(scales), S=S)
    return s


# --------------------------------------------------------------------------------------------------------------------
# staged backward (one function per kernel of csrc/train.cuh)
# --------------------------------------------------------------------------------------------------------------------
def roi_heads_bwd(s: dict, dlogit: torch.Tensor, dloc: torch.Tensor):
    """roi_heads_bwd_kernel: g[t, c] = sum_s (1/s) sum_{i = t - s/2 + 1}^{t + s/2} dpre[i, s, c] (i inside the video)."""
    T = s["T"]
    dpre = torch.cat([dlogit[..., None], dloc], -1)                              # (T, S, 3)
    g = dpre.new_zeros(T, 3)
    for si, sc in enumerate(s["scales"]):
        cs = torch.cat([dpre.new_zeros(1, 3), dpre[:, si].cumsum(0)], 0)
        lo = (torch.arange(T) - sc // 2 + 1).clamp(0, T)
        hi = (torch.arange(T) + sc // 2 + 1).clamp(0, T)
        g = g + (cs[hi] - cs[lo]) / sc
    return g, dpre.sum((0, 1))                                                   # g, (d b_cls, d b_loc0, d b_loc1)


def fc_stack_bwd(s: dict, p: dict, g: torch.Tensor, train: bool):
    """fc_stack_bwd_kernel: heads, then the D shared blocks in reverse.  Returns du0, the per-layer da (for the dW GEMM),
    the per-layer block inputs and the small gradients."""
    wc, wl = p["fc_cls.0.weight"], p["fc_loc.0.weight"]
    wh = torch.stack([wc[0], wl[0], wl[1]], 0)                                   # (3, 128)
    uD = s["uD"]
    d_wh = g.t() @ uD                                                            # (3, 128)
    do = g @ wh                                                                  # (T, 128)
    gam = p["fc_block.3.weight"]
    W = p["fc_block.0.weight"]
    dgam, dbet, db = torch.zeros_like(gam), torch.zeros_like(gam), torch.zeros_like(gam)
    das = [None] * len(s["hs"])
    for l in range(len(s["hs"]) - 1, -1, -1):
        h = s["hs"][l]
        mu, rstd = _ln_stats(h)
        hh = (h - mu) * rstd
        dgam = dgam + (do * hh).sum(0)
        dbet = dbet + do.sum(0)
        dhh = do * gam
        dh = rstd * (dhh - dhh.mean(-1, keepdim=True) - hh * (dhh * hh).mean(-1, keepdim=True))
        da = dh * (2.0 if train else 1.0) * (h > 0).to(h.dtype)
        db = db + da.sum(0)
        das[l] = da
        do = da @ W
    dW = sum(da.t() @ u for da, u in zip(das, s["uin"])) if das else torch.zeros_like(W)
    return {"du0": do, "das": das, "d_cls_w": d_wh[:1], "d_loc_w": d_wh[1:], "d_fcb_w": dW, "d_fcb_b": db,
            "d_fcb_ln_w": dgam, "d_fcb_ln_b": dbet, "d_fc1_b": do.sum(0)}


def ln1024_bwd(s: dict, p: dict, dyn: torch.Tensor):
    """ln1024_bwd_kernel: dy, d gamma, d beta, and the column sums of dy (= d to_out.bias)."""
    y = s["y"]
    mu, rstd = _ln_stats(y)
    yh = (y - mu) * rstd
    dyh = dyn * p["layer_norm.weight"]
    dy = rstd * (dyh - dyh.mean(-1, keepdim=True) - yh * (dyh * yh).mean(-1, keepdim=True))
    return dy, (dyn * yh).sum(0), dyn.sum(0), dy.sum(0)


def attn_bwd_rows(s: dict, p: dict, dout: torch.Tensor):
    """attn_bwd_rows_kernel (per 64-row tile and head; here all rows at once).  dout: (h, T, d).
    Returns dq_part (h,T,d), dv_conv (h,T,d), dW (h,m,d), dkl_part (h,m,d), dtaps (h,33)."""
    a1, w, q, kl, v = s["a1"], s["w"], s["q"], s["kl"], s["v"]
    T = s["T"]
    dp = dout @ w.transpose(1, 2)                                                # (h, T, m)
    ds = a1 * (dp - (dp * a1).sum(-1, keepdim=True))
    dq = ds @ kl
    dW = a1.transpose(1, 2) @ dout
    dkl = ds.transpose(1, 2) @ q
    taps = p["base_model.res_conv.weight"].reshape(H, orc.CONV_TAPS)
    dpad = torch.nn.functional.pad(dout, (0, 0, 16, 16))
    vpad = torch.nn.functional.pad(v, (0, 0, 16, 16))
    # out[r] += w[t] v[r + t - 16]  =>  dv[r'] = sum_t w[t] dout[r' - t + 16],  dw[t] = sum_r dout[r] . v[r + t - 16]
    dv = sum(taps[:, t, None, None] * dpad[:, 32 - t:32 - t + T] for t in range(orc.CONV_TAPS))
    dt = torch.stack([(dout * vpad[:, t:t + T]).sum((1, 2)) for t in range(orc.CONV_TAPS)], 1)
    return dq, dv, dW, dkl, dt


def pinv_bwd(s: dict, dW: torch.Tensor):
    """pinv_bwd_kernel per (video, head): W = Z B and the six Newton-Schulz steps in reverse.
    Returns dB (h,m,d), dA2 without the start-scale term (h,m,m), dc per head (h,)."""
    a, b, zs, c = s["a2"], s["b"], s["zs"], s["c"]
    eye = torch.eye(M, dtype=a.dtype)
    z = zs[-1]
    dz = dW @ b.transpose(1, 2)
    dB = z.transpose(1, 2) @ dW
    dA = torch.zeros_like(a)
    for k in range(orc.PINV_ITERS - 1, -1, -1):
        zk = zs[k]
        pz = a @ zk
        t1 = 7 * eye - pz
        t2 = 15 * eye - pz @ t1
        t3 = 13 * eye - pz @ t2
        dt3 = 0.25 * zk.transpose(1, 2) @ dz
        dz_new = 0.25 * dz @ t3.transpose(1, 2)
        dp = -dt3 @ t2.transpose(1, 2)
        dt2 = -pz.transpose(1, 2) @ dt3
        dp = dp - dt2 @ t1.transpose(1, 2)
        dt1 = -pz.transpose(1, 2) @ dt2
        dp = dp - dt1
        dA = dA + dp @ zk.transpose(1, 2)
        dz = dz_new + a.transpose(1, 2) @ dp
    dA = dA + dz.transpose(1, 2) / c
    dc = -(dz * a.transpose(1, 2)).sum((1, 2)) / (c * c)
    return dB, dA, dc


def attn2_bwd(s: dict, dA_part: torch.Tensor, dc: torch.Tensor):
    """attn2_bwd_kernel: start-scale term (c = max row sum x max column sum over ALL heads, nystroformer.py:16-19; the row
    sums of a softmax are constant, so only the column maximum carries a gradient: column j* of head h*), then the
    softmax back-substitution.  Returns dql (h,m,d), dkl (h,m,d) contributions."""
    a2, ql, kl = s["a2"], s["ql"], s["kl"]
    dA = dA_part.clone()
    col = s["col"]                                                               # (h, m)
    flat = int(torch.argmax(col.reshape(-1)))
    hs, js = flat // M, flat % M
    # d c / d a2[hs, i, js] = rmax for every i; the row-sum part is annihilated by the softmax back-substitution, but it
    # is kept so that this stage equals autograd term by term: + cmax on row i* of head h'*
    rs = a2.sum(-1)
    flat_r = int(torch.argmax(rs.reshape(-1)))
    dA[hs, :, js] += dc.sum() * s["rmax"]
    dA[flat_r // M, flat_r % M, :] += dc.sum() * s["cmax"]
    ds = a2 * (dA - (dA * a2).sum(-1, keepdim=True))
    return ds @ kl, ds.transpose(1, 2) @ ql


def attn_bwd_keys(s: dict, dB: torch.Tensor):
    """attn_bwd_keys_kernel (per 64-key tile and head): softmax over keys, delta_j = <dB_j, B_j>.
    Returns dk_part (h,T,d), dv_agg (h,T,d), dql contribution (h,m,d)."""
    a3, v, k, ql, b = s["a3"], s["v"], s["k"], s["ql"], s["b"]
    da3 = dB @ v.transpose(1, 2)                                                 # (h, m, T)
    delta = (dB * b).sum(-1, keepdim=True)
    ds = a3 * (da3 - delta)
    return ds.transpose(1, 2) @ ql, a3.transpose(1, 2) @ dB, ds @ k


def dqkv_finish(s: dict, dq_part, dk_part, dv, dql, dkl):
    """dqkv_finish_kernel: landmark means back to their rows, the 1/8 of q, head merge.  Returns (T, 1536)."""
    T, pad, seg = s["T"], s["pad"], s["seg"]
    j = (torch.arange(T) + pad) // seg
    dq = 0.125 * (dq_part + dql[:, j] / seg)
    dk = dk_part + dkl[:, j] / seg
    return torch.cat([t.permute(1, 0, 2).reshape(T, H * D) for t in (dq, dk, dv)], 1)


def dsnet_backward_staged(s: dict, p: dict, dlogit: torch.Tensor, dloc: torch.Tensor, train: bool) -> dict:
    """All parameter gradients of one video from d loss / d logits and d loss / d pred_loc, stage by stage."""
    out = {}
    g, dbh = roi_heads_bwd(s, dlogit, dloc)
    fc = fc_stack_bwd(s, p, g, train)
    du0 = fc["du0"]
    out["fc1.weight"] = du0.t() @ s["yn"]
    dyn = du0 @ p["fc1.weight"]
    dy, dg, dbt, dbo = ln1024_bwd(s, p, dyn)
    out["base_model.to_out.0.weight"] = dy.t() @ s["merged"]
    dmerged = dy @ p["base_model.to_out.0.weight"]
    T = s["T"]
    dout = dmerged.reshape(T, H, D).permute(1, 0, 2)
    dq_part, dv_conv, dW, dkl_a, dtaps = attn_bwd_rows(s, p, dout)
    dB, dA_part, dc = pinv_bwd(s, dW)
    dql_2, dkl_2 = attn2_bwd(s, dA_part, dc)
    dk_part, dv_agg, dql_3 = attn_bwd_keys(s, dB)
    dqkv = dqkv_finish(s, dq_part, dk_part, dv_conv + dv_agg, dql_2 + dql_3, dkl_a + dkl_2)
    out["base_model.to_qkv.weight"] = dqkv.t() @ s["x"]
    out.update({"base_model.to_out.0.bias": dbo, "base_model.res_conv.weight": dtaps.reshape(H, 1, orc.CONV_TAPS, 1),
                "layer_norm.weight": dg, "layer_norm.bias": dbt, "fc1.bias": fc["d_fc1_b"],
                "fc_block.0.weight": fc["d_fcb_w"], "fc_block.0.bias": fc["d_fcb_b"], "fc_block.3.weight": fc["d_fcb_ln_w"],
                "fc_block.3.bias": fc["d_fcb_ln_b"], "fc_cls.0.weight": fc["d_cls_w"], "fc_cls.0.bias": dbh[:1],
                "fc_loc.0.weight": fc["d_loc_w"], "fc_loc.0.bias": dbh[1:]})
    out["_stages"] = {"g": g, "du0": du0, "dyn": dyn, "dy": dy, "dmerged": dmerged, "dW": dW, "dB": dB, "dA_part": dA_part,
                      "dc": dc, "dqkv": dqkv, "das": fc["das"]}
    return out
