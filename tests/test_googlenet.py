"""Frame-feature extractor (second half): the oracle against the real torchvision run (CPU), the device
pipeline against the oracle and the same goldens (GPU, through the C ABI)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import googlenet_oracle as gno
from oracle.dsnet_oracle import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "googlenet.npz")


def _golden():
    g = np.load(GOLD)
    x = gno.synth_frames(g["feats"].shape[0], int(g["x_seed"]))
    assert hashlib.sha256(x.numpy().tobytes()).hexdigest() == str(g["x_sha"])
    return g["feats"], x, gno.synth_googlenet_params(int(g["w_seed"]))


def test_oracle_matches_the_torchvision_run():
    """tests/golden/make_golden_googlenet.py ran torchvision's googlenet the way video_helper.py:36-40,61-73 wraps it."""
    feats, x, p = _golden()
    with torch.no_grad():
        got = gno.pool5_features(x, p).numpy()
    assert got.shape == (3, 1024)
    assert rel_l2(got, feats) < 2e-6
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


def test_parameter_table_matches_torchvision_shapes():
    shapes = gno.conv_shapes()
    assert len(shapes) == 57
    assert sum(co * ci * k * k for co, ci, k in shapes.values()) == 5_585_344     # conv weights of torchvision googlenet (no aux heads)
    # channel bookkeeping of the module chain
    cin = 192
    for name, c_in, c1, c3r, c3, c5r, c5, pp in gno.INCEPTIONS:
        assert c_in == cin, name
        cin = c1 + c3 + c5 + pp
    assert cin == 1024


def test_batchnorm_fold_equals_conv_bn():
    from edsnet_b200.features import fold_batchnorm
    p = gno.synth_googlenet_params(5)
    name = "inception3a.branch2.1"
    x = torch.randn(2, 96, 9, 9, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = F.batch_norm(F.conv2d(x, p[f"{name}.conv.weight"], padding=1), p[f"{name}.bn.running_mean"],
                           p[f"{name}.bn.running_var"], p[f"{name}.bn.weight"], p[f"{name}.bn.bias"], training=False,
                           eps=gno.BN_EPS)
        w, b = fold_batchnorm(p, name)
        got = F.conv2d(x, w.permute(0, 3, 1, 2), b, padding=1)
    assert rel_l2(got.numpy(), ref.numpy()) < 1e-6


def test_cpu_frames_are_refused():
    from edsnet_b200 import GoogLeNetPool5
    with pytest.raises(RuntimeError):
        GoogLeNetPool5(gno.synth_googlenet_params(1), "cpu")


# ------------------------------------------------------------------------------------------------- GPU
def _act(pieces, n, h, w, relu, nchw=False):
    from edsnet_b200.features import _Act
    return _Act(pieces, n, h, w, relu, nchw)


@pytest.mark.gpu
@pytest.mark.parametrize("k,stride,pad,h,w", [(3, 1, 1, 9, 7), (7, 2, 3, 21, 18), (1, 1, 0, 5, 5), (3, 2, 0, 11, 11)])
def test_im2col_planes_vs_unfold(k, stride, pad, h, w):
    from edsnet_b200 import _capi
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(k * 100 + h)
    n, ca, cb = 2, 19, 40                                   # two sources, odd channel counts, column offsets
    a = torch.randn(n * h * w, 24, generator=g).to(dev)
    b = torch.randn(n * h * w, 64, generator=g).to(dev)
    act = _act([(a, 24, 3, ca), (b, 64, 8, cb)], n, h, w, relu=True)
    cc = ca + cb
    kpad = (k * k * cc + 63) // 64 * 64
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    m = n * oh * ow
    lib = _capi.lib()
    planes = torch.zeros(int(lib.edsnet_split_f16_bytes(m, kpad)), dtype=torch.uint8, device=dev)
    ci = act.struct()
    _capi.check(lib.edsnet_cnn_im2col(C.byref(ci), n, h, w, k, k, stride, pad, kpad, planes.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    hi = planes[:m * kpad * 2].view(torch.float16).view(m, kpad).float()
    lo = planes[m * kpad * 2:m * kpad * 4].view(torch.float16).view(m, kpad).float()
    inv = planes[m * kpad * 4:].view(torch.float32)
    got = ((hi + lo) * inv[:, None]).cpu()
    x = torch.relu(torch.cat([a[:, 3:3 + ca], b[:, 8:8 + cb]], dim=1)).cpu().view(n, h, w, cc).permute(0, 3, 1, 2)
    ref = F.unfold(x, k, padding=pad, stride=stride)                       # (n, cc * k * k, L), channel slowest
    ref = ref.view(n, cc, k * k, oh * ow).permute(0, 3, 2, 1).reshape(m, k * k * cc)
    assert torch.equal(got[:, k * k * cc:], torch.zeros(m, kpad - k * k * cc))
    assert rel_l2(got[:, :k * k * cc].numpy(), ref.numpy()) < 2e-7         # 22 bits
    assert float(hi.abs().max()) < 32768.0


@pytest.mark.gpu
@pytest.mark.parametrize("k,stride,pad,h", [(3, 2, 0, 112), (3, 1, 1, 14), (2, 2, 0, 7), (3, 2, 0, 28)])
def test_maxpool_ceil_mode_vs_torch(k, stride, pad, h):
    from edsnet_b200 import _capi
    dev = torch.device("cuda", 0)
    n, c = 2, 33
    a = torch.randn(n * h * h, 48, generator=torch.Generator().manual_seed(h)).to(dev)
    act = _act([(a, 48, 5, c)], n, h, h, relu=True)
    ref = F.max_pool2d(torch.relu(a[:, 5:5 + c]).view(n, h, h, c).permute(0, 3, 1, 2), k, stride=stride, padding=pad,
                       ceil_mode=True)
    oh = ref.shape[2]
    out = torch.empty(n * oh * oh, c, device=dev)
    ci = act.struct()
    _capi.check(_capi.lib().edsnet_cnn_maxpool(C.byref(ci), n, h, h, k, stride, pad, out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out.view(n, oh, oh, c).permute(0, 3, 1, 2), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp16x3", 2e-5), ("fp16x2", 2e-3)])
def test_pool5_features_match_oracle_and_golden(precision, tol):
    from edsnet_b200 import GoogLeNetPool5, _capi
    feats, x, p = _golden()
    dev = torch.device("cuda", 0)
    net = GoogLeNetPool5(p, dev, precision=precision)
    got = net(x.to(dev))
    torch.cuda.synchronize()
    assert _capi.lib().edsnet_debug_tc_status(0) == 0
    got = got.cpu().numpy()
    assert got.shape == (3, 1024) and np.isfinite(got).all()
    err = rel_l2(got, feats)
    print(f"pool5 [{precision}] rel-L2 vs the torchvision run: {err:.2e}, {net.launches} launches")
    assert err < tol
    # frame by frame == batched (rows are independent), and a different batch size
    one = net(x[1:2].to(dev)).cpu().numpy()
    assert np.array_equal(one[0], got[1])


@pytest.mark.gpu
def test_pool5_other_resolution():
    """The extractor is resolution agnostic like the torchvision module (ceil_mode pools, adaptive average)."""
    from edsnet_b200 import GoogLeNetPool5
    p = gno.synth_googlenet_params(9)
    x = F.interpolate(gno.synth_frames(2, 3), size=(160, 200), mode="bilinear", align_corners=False).contiguous()
    with torch.no_grad():
        ref = gno.pool5_features(x, p).numpy()
    dev = torch.device("cuda", 0)
    got = GoogLeNetPool5(p, dev)(x.to(dev)).cpu().numpy()
    assert rel_l2(got, ref) < 2e-5


@pytest.mark.gpu
def test_frames_to_summary_chain_vs_oracle():
    """VideoPreprocessor.run + infer.py:26-36 from the sampled frames on: pool5 features -> segmentation -> scores -> NMS
    -> keyshot summary on the device; checked against the oracle's features (tolerance) and, from the device's own
    features on, against the oracle's host chain (exact: change points, shots, summary)."""
    from edsnet_b200 import GoogLeNetPool5, summarize_frames
    from oracle import dsnet_oracle as orc
    from tests.util import make_model
    dev = torch.device("cuda", 0)
    p = gno.synth_googlenet_params(11)
    lengths, sample_rate = [40, 25], 15
    n_frames = [40 * 15 - 7, 25 * 15 - 3]
    frames = [F.interpolate(gno.synth_frames(t, 100 + t), size=(96, 128), mode="bilinear", align_corners=False).contiguous()
              for t in lengths]
    scales = [4, 8]
    model = make_model(orc.synth_params(9, "xavier"), scales, 5, "fp16x3", dev)
    res = summarize_frames(model, GoogLeNetPool5(p, dev), [f.to(dev) for f in frames], n_frames, 0.5, sample_rate,
                           frame_batch=16)
    for f, t, nf, r in zip(frames, lengths, n_frames, res):
        with torch.no_grad():
            want_feat = gno.pool5_features(f, p).numpy()
        feat = r["features"].cpu().numpy()
        assert rel_l2(feat, want_feat) < 2e-5
        cps, nfps, picks = orc.kts_shots(nf, feat, sample_rate)
        assert np.array_equal(r["change_points"], cps) and np.array_equal(r["nfps"], nfps) and np.array_equal(r["picks"], picks)
        assert r["summary"].shape == (nf,) and r["summary"].sum() <= int(nf * 0.15)
