"""Native training step on the GPU (BASELINE.json config 3) against the staged backward model (oracle/backward_model.py,
pinned to torch.autograd and to the REAL reference's gradients by tests/test_backward_model.py).  `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import backward_model as bm
from oracle import dsnet_oracle as orc
from tests.test_backward_model import check_against_grad_digest, grad_case
from tests.util import make_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# edsnet_grads field -> state-dict name
FIELD2NAME = {"to_qkv_w": "base_model.to_qkv.weight", "to_out_w": "base_model.to_out.0.weight",
              "to_out_b": "base_model.to_out.0.bias", "res_conv_w": "base_model.res_conv.weight",
              "ln_w": "layer_norm.weight", "ln_b": "layer_norm.bias", "fc1_w": "fc1.weight", "fc1_b": "fc1.bias",
              "fcb_w": "fc_block.0.weight", "fcb_b": "fc_block.0.bias", "fcb_ln_w": "fc_block.3.weight",
              "fcb_ln_b": "fc_block.3.bias", "cls_w": "fc_cls.0.weight", "cls_b": "fc_cls.0.bias",
              "loc_w": "fc_loc.0.weight", "loc_b": "fc_loc.0.bias"}
GRAD_TOL = 1e-4          # gradients within 1e-4 of the reference's autograd on all 16 tensors


def _capi():
    from edsnet_b200 import _capi
    return _capi, _capi.lib()


def _no_tc_timeout():
    capi, lib = _capi()
    torch.cuda.synchronize()
    assert lib.edsnet_debug_tc_status(1) == 0, "a tcgen05 pipeline wait timed out"


def _labels(T, S, seed):
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(T, S, generator=g)
    cls = torch.zeros(T, S, dtype=torch.int64)
    cls[r < 0.12] = 1
    cls[(r >= 0.12) & (r < 0.4)] = -1
    loc = torch.randn(T, S, 2, generator=g) * 1.2
    return cls, loc


def _ws_view(ctx, layout, field, shape, dtype=torch.float32):
    off = getattr(layout, field)
    n = int(np.prod(shape))
    itemsize = torch.tensor([], dtype=dtype).element_size()
    return ctx.workspace[off:off + n * itemsize].view(dtype).reshape(shape)


def _layout(ctx):
    capi, lib = _capi()
    L = capi.TrainLayout()
    lib.edsnet_train_workspace_bytes(ctx.cfg, ctx.batch.plan.total_rows, ctx.batch.plan.n_videos, C.byref(L))
    return L


def _zero_grads(model):
    from edsnet_b200 import _capi
    named = model._named_weights()
    return {k: torch.zeros_like(named[k]) for k in _capi.GRAD_FIELDS}


def _oracle_mean_grads(xs, p, scales, depth, labels, keeps=None):
    """float64 staged-model gradients of mean_v loss_v and the per-video losses."""
    p64 = {k: v.double() for k, v in p.items()}
    tot = {k: torch.zeros_like(v) for k, v in p64.items()}
    losses, saved = [], []
    for i, (x, (cl, ll)) in enumerate(zip(xs, labels)):
        with torch.no_grad():
            s = bm.dsnet_forward_saved(x.double(), p64, scales, depth, None if keeps is None else keeps[i])
            loss, cls, loc = bm.reference_losses(s["pred_cls"], s["pred_loc"], cl, ll.double())
            dlogit, dloc = bm.loss_grad_logits(s["pred_cls"], s["pred_loc"], cl, ll.double(), scale=1.0 / len(xs))
            g = bm.dsnet_backward_staged(s, p64, dlogit, dloc, train=keeps is not None)
        for k in tot:
            tot[k] += g[k]
        losses.append((float(loss), float(cls), float(loc)))
        saved.append((s, g["_stages"], dlogit, dloc))
    return tot, losses, saved


# ------------------------------------------------------------------------------------------------ building blocks
@pytest.mark.parametrize("rows,cols", [(333, 1536), (64, 128), (1, 1024), (2500, 128), (800, 512)])
def test_split_t_planes(rows, cols):
    """Transposed operand planes: hi + lo reproduces the scaled transpose to 2^-22, padding is zero, scales are powers of
    two that put every output row's maximum into [2^14, 2^15)."""
    capi, lib = _capi()
    g = torch.Generator().manual_seed(rows + cols)
    src = (torch.randn(rows, cols, generator=g) * torch.logspace(-6, 2, cols)[None, :]).to(DEV)
    kp = (rows + 63) // 64 * 64
    dst = torch.zeros(lib.edsnet_split_f16_bytes(cols, kp) + 4 * cols, dtype=torch.uint8, device=DEV)
    capi.check(lib.edsnet_split_f16_t(src.data_ptr(), rows, cols, dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    hi = dst[:cols * kp * 2].view(torch.float16).reshape(cols, kp).double().cpu()
    lo = dst[cols * kp * 2:cols * kp * 4].view(torch.float16).reshape(cols, kp).double().cpu()
    inv = dst[cols * kp * 4:cols * kp * 4 + cols * 4].view(torch.float32).double().cpu()
    want = src.t().double().cpu()
    rec = (hi + lo) * inv[:, None]
    assert torch.equal(rec[:, rows:], torch.zeros(cols, kp - rows, dtype=torch.float64))
    err = (rec[:, :rows] - want).abs().amax(1) / want.abs().amax(1)
    assert float(err.max()) < 2.0 ** -21
    mx = (hi.abs().amax(1))
    assert bool(((mx >= 2.0 ** 14 - 8) & (mx <= 2.0 ** 15)).all())
    assert bool((torch.log2(inv) == torch.log2(inv).round()).all())


@pytest.mark.parametrize("M,N,K", [(1536, 1024, 333), (128, 128, 2500), (128, 1024, 90), (1024, 512, 800)])
def test_dw_product_on_tensor_cores(M, N, K):
    """dW = dY^T X as the backward runs it: both operands through split_f16_t, three split-fp16 tcgen05 passes."""
    capi, lib = _capi()
    g = torch.Generator().manual_seed(M + N + K)
    dY = (torch.randn(K, M, generator=g) * 1e-3).to(DEV)
    X = (torch.randn(K, N, generator=g) * 0.05).to(DEV)
    kp = (K + 63) // 64 * 64
    st = torch.cuda.current_stream().cuda_stream
    A16 = torch.empty(lib.edsnet_split_f16_bytes(M, kp) + 4 * M, dtype=torch.uint8, device=DEV)
    B16 = torch.empty(lib.edsnet_split_f16_bytes(N, kp) + 4 * N, dtype=torch.uint8, device=DEV)
    capi.check(lib.edsnet_split_f16_t(dY.data_ptr(), K, M, A16.data_ptr(), st))
    capi.check(lib.edsnet_split_f16_t(X.data_ptr(), K, N, B16.data_ptr(), st))
    Cd = torch.full((M, N), float("nan"), device=DEV)
    capi.check(lib.edsnet_gemm(capi.PREC_FP16X3, 0, None, A16.data_ptr(), None, B16.data_ptr(), Cd.data_ptr(), M, N, kp,
                               None, None, 0, st))
    _no_tc_timeout()
    ref = dY.double().t() @ X.double()
    err = float((Cd.double() - ref).norm() / ref.norm())
    assert err < 2e-6, err


def test_dropout_mask_equals_the_philox_restatement():
    capi, lib = _capi()
    rows, depth, seed, offset = 301, 5, 0x1234_5678_9ABC_DEF0, (1 << 40) + 17
    out = torch.empty((depth, rows, 128), dtype=torch.uint8, device=DEV)
    capi.check(lib.edsnet_dropout_mask(seed, offset, rows, depth, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    want = bm.dropout_mask(rows, depth, seed, offset)
    assert np.array_equal(out.cpu().numpy().astype(bool), want)


# ------------------------------------------------------------------------------------------------ forward
@pytest.mark.parametrize("dropout", [False, True])
def test_train_forward_and_saved_activations(dropout):
    from edsnet_b200 import native_train as nt
    T, scales, depth = 203, [4, 8], 4
    p = orc.synth_params(61, "xavier")
    x = orc.synth_features(T, 62)
    model = make_model(p, scales, depth, "fp16x3", DEV)
    seed, offset = 987654321, 3
    ctx = nt.train_forward(model, x.to(DEV), [T], dropout, seed, offset)
    _no_tc_timeout()
    keep = torch.from_numpy(bm.dropout_mask(T, depth, seed, offset)) if dropout else None
    with torch.no_grad():
        s = bm.dsnet_forward_saved(x.double(), {k: v.double() for k, v in p.items()}, scales, depth, keep)
    assert orc.rel_l2(ctx.pred_cls.cpu().numpy(), s["pred_cls"].numpy()) < 1e-5
    assert orc.rel_l2(ctx.pred_loc.cpu().numpy(), s["pred_loc"].numpy()) < 1e-5
    L = _layout(ctx)
    got = {"merged": _ws_view(ctx, L, "merged", (T, 512)), "y": _ws_view(ctx, L, "y", (T, 1024)),
           "yn": _ws_view(ctx, L, "yn", (T, 1024)), "u0": _ws_view(ctx, L, "uin", (T, 128)),
           "hs": _ws_view(ctx, L, "hs", (depth, T, 128)), "uD": _ws_view(ctx, L, "u_last", (T, 128))}
    want = {"merged": s["merged"], "y": s["y"], "yn": s["yn"], "u0": s["u0"], "hs": torch.stack(s["hs"]), "uD": s["uD"]}
    for k in got:
        e = orc.rel_l2(got[k].cpu().numpy(), want[k].numpy())
        assert e < 2e-5, (k, e)
    if dropout:
        # the mask is visible in the saved rows: a dropped position is exactly zero
        hs = got["hs"].cpu().numpy()
        assert np.all(hs[~keep.numpy()] == 0.0)
        # eval-mode inference differs, train-mode without the mask differs
        with torch.no_grad():
            c_eval, _ = model(x[None].to(DEV))
        assert not torch.allclose(c_eval, ctx.pred_cls, atol=1e-4)


# ------------------------------------------------------------------------------------------------ backward, stage by stage
def test_loss_gradient_kernel():
    from edsnet_b200 import native_train as nt
    T, scales, depth = 150, [4, 8, 16, 32], 2
    p = orc.synth_params(5, "xavier")
    x = orc.synth_features(T, 6)
    model = make_model(p, scales, depth, "fp16x3", DEV)
    ctx = nt.train_forward(model, x.to(DEV), [T], False, 0, 0)
    cl, ll = _labels(T, len(scales), 9)
    loss, d_logit, d_loc = nt.loss_and_grad(ctx, cl.to(torch.int32).to(DEV), ll.to(DEV), lambda_reg=0.7, scale=0.25)
    pc, pl = ctx.pred_cls.cpu().double(), ctx.pred_loc.cpu().double()
    want_loss = bm.reference_losses(pc, pl, cl, ll.double(), 0.7)
    want_dl, want_dc = bm.loss_grad_logits(pc, pl, cl, ll.double(), 0.7, 0.25)
    got = loss.cpu().numpy()[0]
    assert np.allclose(got, [float(v) for v in want_loss], rtol=2e-6, atol=1e-7)
    assert orc.rel_l2(d_logit.cpu().numpy(), want_dl.numpy()) < 1e-6
    assert orc.rel_l2(d_loc.cpu().numpy(), want_dc.numpy()) < 1e-6


@pytest.mark.parametrize("T,scales,depth,init", [(203, [4, 8], 4, "xavier"), (64, [12], 5, "default"),
                                                  (450, [4, 8, 16, 32], 5, "xavier"), (37, [4, 32], 1, "xavier")])
def test_backward_stages_and_gradients(T, scales, depth, init):
    """Every intermediate of the backward against the staged model, then all 16 parameter gradients (dropout off)."""
    from edsnet_b200 import native_train as nt
    p = orc.synth_params(71, init)
    x = orc.synth_features(T, 72)
    S = len(scales)
    cl, ll = _labels(T, S, 73)
    model = make_model(p, scales, depth, "fp16x3", DEV)
    ctx = nt.train_forward(model, x.to(DEV), [T], False, 0, 0)
    loss, d_logit, d_loc = nt.loss_and_grad(ctx, cl.to(torch.int32).to(DEV), ll.to(DEV))
    grads = _zero_grads(model)
    nt.train_backward(ctx, d_logit, d_loc, grads, logit_grad=True)
    _no_tc_timeout()
    want, losses, saved = _oracle_mean_grads([x], p, scales, depth, [(cl, ll)])
    s, st, _, _ = saved[0]
    assert abs(float(loss[0, 0]) - losses[0][0]) < 5e-6 * max(1.0, abs(losses[0][0]))
    L = _layout(ctx)
    hm = lambda t: t.permute(1, 0, 2).reshape(T, 512)                      # (h, T, d) -> head-merged rows
    stage = {
        "g": (_ws_view(ctx, L, "g", (T, 4))[:, :3], st["g"]),
        "du0": (_ws_view(ctx, L, "du0", (T, 128)), st["du0"]),
        "das": (_ws_view(ctx, L, "das", (depth, T, 128)), torch.stack(st["das"])),
        "dyn": (_ws_view(ctx, L, "dyn", (T, 1024)), st["dyn"]),
        "dy": (_ws_view(ctx, L, "dy", (T, 1024)), st["dy"]),
        "dmerged": (_ws_view(ctx, L, "dmerged", (T, 512)), st["dmerged"]),
        "dW": (_ws_view(ctx, L, "dw_att", (8, 64, 64)), st["dW"]),
        "dB": (_ws_view(ctx, L, "db_att", (8, 64, 64)), st["dB"]),
        "dA2": (_ws_view(ctx, L, "da2", (8, 64, 64)), st["dA_part"]),
        "dc": (_ws_view(ctx, L, "dc_part", (8,)), st["dc"]),
        "dqkv": (_ws_view(ctx, L, "dqkv", (T, 1536)), st["dqkv"]),
    }
    report = {}
    for k, (got, exp) in stage.items():
        report[k] = orc.rel_l2(got.cpu().numpy(), exp.numpy())
    print("stages", {k: f"{v:.1e}" for k, v in report.items()})
    # dA2 / dc are intermediates of the fp32 pseudo-inverse chain whose error lives almost entirely in directions the
    # softmax back-substitution annihilates (row constants) or that cancel in the sum (dc): what they feed -- dqkv and the
    # parameter gradients below -- is held to the full bar; they themselves only to "not broken"
    loose = {"dA2": 5e-3, "dc": float("inf")}
    for k, e in report.items():
        assert e < loose.get(k, GRAD_TOL), (k, e, report)
    errs = {FIELD2NAME[f]: orc.rel_l2(grads[f].cpu().numpy(), want[FIELD2NAME[f]].numpy()) for f in grads}
    print("grads", {k: f"{v:.1e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e < GRAD_TOL, (k, e, errs)


@pytest.mark.parametrize("name", ["g_T320_s12", "g_T150_s4_8_16_32", "g_T77_s4_8_default"])
def test_gradients_match_the_reference_run(name):
    """All 16 gradient tensors against what the REAL reference's loss.backward() produced (reference goldens)."""
    from edsnet_b200 import native_train as nt
    x, p, scales, depth, cl, ll, g = grad_case(name)
    model = make_model(p, scales, depth, "fp16x3", DEV)
    T = x.shape[0]
    ctx = nt.train_forward(model, x.to(DEV), [T], False, 0, 0)
    loss, d_logit, d_loc = nt.loss_and_grad(ctx, cl.to(torch.int32).to(DEV), ll.to(DEV))
    grads = _zero_grads(model)
    nt.train_backward(ctx, d_logit, d_loc, grads, logit_grad=True)
    _no_tc_timeout()
    got = loss.cpu().numpy()[0]
    assert abs(got[0] - float(g["loss"])) < 5e-6 and abs(got[1] - float(g["cls_loss"])) < 5e-6
    worst = check_against_grad_digest(g, {FIELD2NAME[f]: grads[f].cpu().numpy() for f in grads}, GRAD_TOL)
    print(name, "worst gradient error vs the reference run:", worst)


def test_packed_videos_and_dropout_gradients():
    """Three videos of different lengths in one packed call, Dropout ON: gradient of the mean loss equals the staged model
    with the SAME Philox mask (rows of the mask = packed rows)."""
    from edsnet_b200 import native_train as nt
    lengths, scales, depth = [70, 200, 129], [4, 8], 3
    p = orc.synth_params(81, "xavier")
    xs = [orc.synth_features(t, 82 + i) for i, t in enumerate(lengths)]
    labels = [_labels(t, 2, 90 + i) for i, t in enumerate(lengths)]
    model = make_model(p, scales, depth, "fp16x3", DEV)
    seed, offset = 424242, 11
    ctx = nt.train_forward(model, torch.cat(xs).to(DEV), lengths, True, seed, offset)
    cl = torch.cat([c for c, _ in labels]).to(torch.int32).to(DEV)
    ll = torch.cat([l for _, l in labels]).to(DEV)
    loss, d_logit, d_loc = nt.loss_and_grad(ctx, cl, ll, scale=1.0 / len(lengths))
    grads = _zero_grads(model)
    nt.train_backward(ctx, d_logit, d_loc, grads, logit_grad=True)
    _no_tc_timeout()
    mask = torch.from_numpy(bm.dropout_mask(sum(lengths), depth, seed, offset))
    cu = np.concatenate([[0], np.cumsum(lengths)])
    keeps = [mask[:, cu[i]:cu[i + 1]] for i in range(len(lengths))]
    want, losses, _ = _oracle_mean_grads(xs, p, scales, depth, labels, keeps)
    assert np.allclose(loss.cpu().numpy()[:, 0], [l[0] for l in losses], rtol=1e-5, atol=1e-6)
    errs = {FIELD2NAME[f]: orc.rel_l2(grads[f].cpu().numpy(), want[FIELD2NAME[f]].numpy()) for f in grads}
    print("grads", {k: f"{v:.1e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e < GRAD_TOL, (k, e, errs)


# ------------------------------------------------------------------------------------------------ drop-in surface
def test_reference_training_loop_body_on_the_drop_in_model():
    """anchor_based/train.py:113-127 verbatim on edsnet_b200.DSNet: model(seq) in train() mode, the reference-shaped torch
    losses, loss.backward(), torch.optim.Adam.  The gradients land in the backward kernels and equal the staged model
    with the mask the call drew."""
    from edsnet_b200 import training as tr
    T, scales, depth = 180, [4, 8, 16, 32], 5
    p = orc.synth_params(91, "xavier")
    x = orc.synth_features(T, 92)
    cl, ll = _labels(T, 4, 93)
    model = make_model(p, scales, depth, "fp16x3", DEV).train()
    opt = torch.optim.Adam([q for q in model.parameters() if q.requires_grad], lr=5e-5, weight_decay=1e-5)
    pred_cls, pred_loc = model(x[None].to(DEV))
    assert pred_cls.requires_grad and pred_cls.shape == (T, 4) and pred_loc.shape == (T, 4, 2)
    loss = tr.cls_loss(pred_cls, cl.to(DEV).float()) + tr.loc_loss(pred_loc, ll.to(DEV), cl.to(DEV).float())
    opt.zero_grad()
    loss.backward()
    _no_tc_timeout()
    keep = torch.from_numpy(bm.dropout_mask(T, depth, model._drop_seed, model._drop_offset))
    want, losses, _ = _oracle_mean_grads([x], p, scales, depth, [(cl, ll)], [keep])
    assert abs(float(loss) - losses[0][0]) < 1e-5
    named = dict(model.named_parameters())
    for k in p:
        e = orc.rel_l2(named[k].grad.cpu().numpy(), want[k].numpy())
        assert e < GRAD_TOL, (k, e)
    before = {k: v.detach().clone() for k, v in named.items()}
    opt.step()
    assert all(not torch.equal(before[k], named[k].detach()) for k in p)
    # eval() + gradients: no dropout, still the native backward
    model.eval()
    c2, l2 = model(x[None].to(DEV))
    assert c2.requires_grad
    with torch.no_grad():
        c3, _ = model(x[None].to(DEV))
    assert orc.rel_l2(c2.detach().cpu().numpy(), c3.cpu().numpy()) < 1e-5


def test_native_step_follows_torch_adam():
    """NativeDataParallelStep (Dropout off) against torch.optim.Adam on the float64 staged gradients: same losses and the
    same parameter updates over four optimiser steps of two videos each."""
    from edsnet_b200 import training as tr
    scales, depth = [4, 8], 3
    p = orc.synth_params(101, "xavier")
    lengths = [120, 77, 200, 64]
    xs = [orc.synth_features(t, 110 + i) for i, t in enumerate(lengths)]
    labels = [_labels(t, 2, 120 + i) for i, t in enumerate(lengths)]
    model = make_model(p, scales, depth, "fp16x3", DEV)
    stepper = tr.NativeDataParallelStep(model, dropout=False)
    ref = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    opt = torch.optim.Adam(list(ref.values()), lr=5e-5, weight_decay=1e-5)
    for it in range(4):
        sel = [(2 * it) % 4, (2 * it + 1) % 4]
        stepper.step([xs[i].to(DEV) for i in sel], [labels[i][0].numpy() for i in sel], [labels[i][1].numpy() for i in sel])
        got_loss = stepper.last_loss()
        grads, losses, _ = _oracle_mean_grads([xs[i] for i in sel], {k: v.detach() for k, v in ref.items()}, scales, depth,
                                              [labels[i] for i in sel])
        opt.zero_grad()
        for k, v in ref.items():
            v.grad = grads[k]
        opt.step()
        assert abs(got_loss - float(np.mean([l[0] for l in losses]))) < 2e-5, (it, got_loss, losses)
    _no_tc_timeout()
    named = dict(model.named_parameters())
    for k in p:
        d_ref = (ref[k].detach() - p[k].double()).numpy()
        d_got = (named[k].detach().cpu().double() - p[k].double()).numpy()
        e = float(np.linalg.norm(d_got - d_ref) / max(np.linalg.norm(d_ref), 1e-30))
        assert e < 2e-2, (k, e)          # Adam divides by sqrt(v): a 1e-5 gradient difference on a near-zero entry flips a step
        assert np.abs(d_got).max() <= 4 * 5e-5 * 1.01 + 1e-9                # |update| <= lr per step
    # the flat parameter views are what the model scores with afterwards
    model.eval()
    with torch.no_grad():
        c, _ = model(xs[0][None].to(DEV))
        rc, _ = orc.dsnet_forward(xs[0].double(), {k: v.detach() for k, v in ref.items()}, scales, depth)
    assert orc.rel_l2(c.cpu().numpy(), rc.numpy()) < 1e-4
