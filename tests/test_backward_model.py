"""CPU: the staged backward model (oracle/backward_model.py, one function per training kernel) against torch.autograd
of the oracle forward in float64, and the loss / dropout-mask restatements the GPU tests rely on."""
import numpy as np
import pytest
import torch

from oracle import backward_model as bm
from oracle import dsnet_oracle as orc


def _case(T, scales, depth, init, seed):
    p = {k: v.double() for k, v in orc.synth_params(seed, init).items()}
    x = orc.synth_features(T, seed + 1).double()
    g = torch.Generator().manual_seed(seed)
    S = len(scales)
    cls_label = torch.zeros(T, S, dtype=torch.int64)
    r = torch.rand(T, S, generator=g)
    cls_label[r < 0.15] = 1
    cls_label[(r > 0.15) & (r < 0.5)] = -1
    loc_label = torch.randn(T, S, 2, generator=g).double() * 1.5        # some |d| > 1: both smooth-L1 branches
    return p, x, cls_label, loc_label


@pytest.mark.parametrize("T,scales,depth,init", [(90, [4, 8], 3, "xavier"), (64, [12], 5, "default"),
                                                  (150, [4, 8, 16, 32], 2, "xavier")])
def test_staged_backward_equals_autograd(T, scales, depth, init):
    p, x, cl, ll = _case(T, scales, depth, init, 7)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rc, rl = orc.dsnet_forward(x, pr, scales, depth)
    loss, _, _ = bm.reference_losses(rc, rl, cl, ll)
    loss.backward()
    with torch.no_grad():
        s = bm.dsnet_forward_saved(x, p, scales, depth)
        assert orc.rel_l2(s["pred_cls"].numpy(), rc.detach().numpy()) < 1e-12
        assert orc.rel_l2(s["pred_loc"].numpy(), rl.detach().numpy()) < 1e-12
        dlogit, dloc = bm.loss_grad_logits(s["pred_cls"], s["pred_loc"], cl, ll)
        grads = bm.dsnet_backward_staged(s, p, dlogit, dloc, train=False)
    for k in p:
        e = orc.rel_l2(grads[k].numpy(), pr[k].grad.numpy())
        assert e < 1e-9, (k, e)


def test_staged_backward_with_dropout_mask():
    T, scales, depth = 70, [4, 8], 4
    p, x, cl, ll = _case(T, scales, depth, "xavier", 11)
    keep = torch.from_numpy(bm.dropout_mask(T, depth, seed=1234, offset=5))
    assert 0.4 < keep.float().mean() < 0.6
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    s = bm.dsnet_forward_saved(x, pr, scales, depth, keep)
    loss, _, _ = bm.reference_losses(s["pred_cls"], s["pred_loc"], cl, ll)
    loss.backward()
    with torch.no_grad():
        s0 = bm.dsnet_forward_saved(x, p, scales, depth, keep)
        dlogit, dloc = bm.loss_grad_logits(s0["pred_cls"], s0["pred_loc"], cl, ll)
        grads = bm.dsnet_backward_staged(s0, p, dlogit, dloc, train=True)
    for k in p:
        e = orc.rel_l2(grads[k].numpy(), pr[k].grad.numpy())
        assert e < 1e-9, (k, e)
    # the oracle's own train-mode forward takes the same mask
    with torch.no_grad():
        c, l = orc.dsnet_forward(x, p, scales, depth, keep=keep)
    assert torch.allclose(c, s0["pred_cls"]) and torch.allclose(l, s0["pred_loc"])


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors: zero, all-ones and the pi-digits case)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = bm.philox4x32_10(np.asarray([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in got) == want


def test_loss_gradient_closed_form():
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(50, 4, generator=g) * 3).double().requires_grad_(True)
    loc = torch.randn(50, 4, 2, generator=g).double().requires_grad_(True)
    lab = torch.randint(-1, 2, (50, 4), generator=g)
    tgt = torch.randn(50, 4, 2, generator=g).double() * 2
    loss, _, _ = bm.reference_losses(torch.sigmoid(logits), loc, lab, tgt, lambda_reg=0.7)
    loss.backward()
    dl, dc = bm.loss_grad_logits(torch.sigmoid(logits.detach()), loc.detach(), lab, tgt, lambda_reg=0.7)
    assert torch.allclose(dl, logits.grad, atol=1e-12) and torch.allclose(dc, loc.grad, atol=1e-12)


# ------------------------------------------------------------------------------------------------ reference goldens
GRAD = None


def _grad_golden():
    global GRAD
    if GRAD is None:
        from tests.util import load_npz
        GRAD = load_npz("grad_golden.npz")
    return GRAD


def grad_case(name):
    """(x, p, scales, depth, cls_label int64 (T,S), loc_label float32 (T,S,2), golden record) of one reference case."""
    G = _grad_golden()
    g = {k.split("/", 1)[1]: G[k] for k in G.files if k.startswith(name + "/")}
    x = orc.synth_features(int(g["T"]), int(g["x_seed"]))
    p = orc.synth_params(int(g["w_seed"]), str(g["init"]))
    return x, p, [int(s) for s in g["scales"]], int(g["fc_depth"]), torch.from_numpy(g["cls_label"].astype(np.int64)), \
        torch.from_numpy(g["loc_label"]), g


def check_against_grad_digest(g, grads, tol):
    """grads: {state-dict name: array}.  Compares with what tests/golden/make_grad_golden.py kept of the REFERENCE's
    gradient: small tensors in full, the weight matrices by norm / random projection / strided sample."""
    worst = 0.0
    for k in orc.PARAM_SHAPES:
        mine = np.asarray(grads[k], dtype=np.float64)
        if f"grad/{k}/full" in g:
            e = orc.rel_l2(mine, g[f"grad/{k}/full"])
        else:
            flat = mine.reshape(-1)
            idx = np.linspace(0, flat.size - 1, 4096).astype(np.int64)
            proj = np.random.default_rng(12345).standard_normal(flat.size)
            e = max(orc.rel_l2(flat[idx], g[f"grad/{k}/sample"]),
                    abs(np.linalg.norm(flat) - float(g[f"grad/{k}/norm"])) / float(g[f"grad/{k}/norm"]),
                    abs(flat @ proj - float(g[f"grad/{k}/proj"])) / float(g[f"grad/{k}/norm"]))
        assert e < tol, (k, e)
        worst = max(worst, e)
    return worst


@pytest.mark.parametrize("name", ["g_T320_s12", "g_T150_s4_8_16_32", "g_T77_s4_8_default"])
def test_oracle_gradients_match_the_reference_run(name):
    """The oracle's autograd (fp32, as the reference computes) and the staged backward model (fp64) against the gradients
    the REAL reference produced with its own losses (tests/golden/make_grad_golden.py)."""
    x, p, scales, depth, cl, ll, g = grad_case(name)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rc, rl = orc.dsnet_forward(x, pr, scales, depth)
    loss, cls, loc = bm.reference_losses(rc, rl, cl, ll)
    assert abs(float(loss) - float(g["loss"])) < 2e-6 and abs(float(cls) - float(g["cls_loss"])) < 2e-6
    loss.backward()
    check_against_grad_digest(g, {k: v.grad.numpy() for k, v in pr.items()}, 2e-5)
    with torch.no_grad():
        p64 = {k: v.double() for k, v in p.items()}
        s = bm.dsnet_forward_saved(x.double(), p64, scales, depth)
        dlogit, dloc = bm.loss_grad_logits(s["pred_cls"], s["pred_loc"], cl, ll.double())
        grads = bm.dsnet_backward_staged(s, p64, dlogit, dloc, train=False)
    check_against_grad_digest(g, {k: grads[k].numpy() for k in p}, 2e-5)
