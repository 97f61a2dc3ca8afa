"""Parity of the CUDA path (through the C ABI) against the oracle and the reference-generated golden vectors.
Every test here needs a B200: `pytest -m gpu`."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import dsnet_oracle as orc
from tests.util import TOL, golden_case, load_npz, make_model

pytestmark = pytest.mark.gpu

FWD = load_npz("forward_golden.npz")
DN = load_npz("decode_nms_golden.npz")
CASES = list(FWD["forward_cases"])
DEV = "cuda:0"


_PARITY_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_r02.jsonl")


def _record_parity(test, case, rec):
    """Append one line per case to gpurun_out/parity_r02.jsonl (copied to profiles/ and committed): how close the GPU
    path is to the REFERENCE RUN where exact identity cannot be demanded unconditionally."""
    import json
    try:
        os.makedirs(os.path.dirname(_PARITY_LOG), exist_ok=True)
        with open(_PARITY_LOG, "a") as f:
            f.write(json.dumps({"test": test, "case": case, **rec}) + "\n")
    except OSError:
        pass


def _lib():
    from edsnet_b200 import _capi
    return _capi, _capi.lib()


def _no_tc_timeout():
    capi, lib = _lib()
    torch.cuda.synchronize()
    assert lib.edsnet_debug_tc_status(1) == 0, "a tcgen05 pipeline wait timed out"


# ------------------------------------------------------------------------------------------------ GEMM stage
@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16", "fp16x2"])
@pytest.mark.parametrize("M,N,K,epi", [(320, 1536, 1024, 1), (320, 1024, 512, 3), (320, 128, 1024, 2),
                                        (37, 1536, 1024, 1), (1000, 1024, 512, 3), (129, 128, 1024, 2),
                                        (4096, 1536, 1024, 0)])
def test_gemm_stage(precision, M, N, K, epi):
    _gemm_check(precision, M, N, K, epi)


def _gemm_check(precision, M, N, K, epi, tol=None):
    capi, lib = _lib()
    g = torch.Generator().manual_seed(M * 7 + N + epi)
    A = torch.randn(M, K, generator=g) * 0.05
    B = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    bias = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g) * 0.1
    ref = A.double() @ B.double().t()
    if epi == 1:
        ref[:, :512] *= 0.125
    if epi >= 2:
        ref += bias.double()
    if epi == 3:
        ref += res.double()
    Ad, Bd, bd, rd = (t.to(DEV).contiguous() for t in (A, B, bias, res))
    Cd = torch.full((M, N), float("nan"), device=DEV)
    A16 = B16 = None
    st = torch.cuda.current_stream().cuda_stream
    if precision != "fp32":
        A16 = torch.empty(lib.edsnet_split_f16_bytes(M, K), dtype=torch.uint8, device=DEV)
        B16 = torch.empty(lib.edsnet_split_f16_bytes(N, K), dtype=torch.uint8, device=DEV)
        capi.check(lib.edsnet_split_f16(Ad.data_ptr(), A16.data_ptr(), M, K, st))
        capi.check(lib.edsnet_split_f16(Bd.data_ptr(), B16.data_ptr(), N, K, st))
    capi.check(lib.edsnet_gemm(capi.PRECISIONS[precision], epi, Ad.data_ptr(),
                               A16.data_ptr() if A16 is not None else None, Bd.data_ptr(),
                               B16.data_ptr() if B16 is not None else None, Cd.data_ptr(), M, N, K,
                               bd.data_ptr(), rd.data_ptr(), 512, st))
    _no_tc_timeout()
    out = Cd.cpu().double()
    assert torch.isfinite(out).all()
    err = float((out - ref).norm() / ref.norm())
    # operand rounding: fp32/fp16x3 ~ 2^-22..2^-24 per product, fp16 ~ 2^-11 on both operands, fp16x2 ~ 2^-11 on A only
    assert err < (tol or {"fp32": 1e-6, "fp16x3": 1e-6, "fp16": 1e-3, "fp16x2": 5e-4}[precision]), err
    if precision == "fp16x2":
        # two passes must be EXACTLY the product of the rounded A with the 22-bit B (up to fp32 accumulation): compare
        # against that product to make sure the cross term A_hi.B_lo really is in the result
        s = torch.ldexp(torch.ones(M, dtype=torch.float64), (14 - torch.floor(torch.log2(A.abs().amax(1).double()))).int())
        A_hi = (A.double() * s[:, None]).half().double() / s[:, None]
        ref2 = A_hi @ B.double().t()
        if epi == 1:
            ref2[:, :512] *= 0.125
        if epi >= 2:
            ref2 += bias.double()
        if epi == 3:
            ref2 += res.double()
        err2 = float((out - ref2).norm() / ref2.norm())
        assert err2 < 2e-6, err2


@pytest.mark.parametrize("variant", [3, 4])
def test_gemm_kernel_variants(variant):
    """The double-buffered two-accumulator variant (3) and the CTA-pair cta_group::2 kernel (4) behind the profiling
    knob compute the same products as the default kernel (ragged M, all fp32 epilogues)."""
    capi, lib = _lib()
    capi.check(lib.edsnet_debug_set_tc_variant(variant))
    try:
        for M, N, K, epi in [(1000, 1536, 1024, 1), (333, 1024, 512, 3), (4096, 128, 1024, 2), (77, 1536, 1024, 0)]:
            # variant 3 keeps all 64 hi.hi steps of a K = 1024 product in ONE truncating accumulator: ~1.3e-6
            _gemm_check("fp16x3", M, N, K, epi, tol=2.5e-6 if variant == 3 else None)
    finally:
        capi.check(lib.edsnet_debug_set_tc_variant(0))


def test_two_pass_pair_kernel():
    """The two-pass CTA-pair kernel (cta_group::2: B_hi in the leader, B_lo in its peer; EDSNET_TC_VARIANT=5) computes the
    same products as the one-CTA two-pass kernel (ragged M, odd numbers of pair tiles, all fp32 epilogues)."""
    capi, lib = _lib()
    capi.check(lib.edsnet_debug_set_tc_variant(5))
    try:
        for M, N, K, epi in [(1000, 1536, 1024, 1), (333, 1024, 512, 3), (4096, 128, 1024, 2), (77, 1536, 1024, 0),
                             (257, 128, 128, 0)]:
            _gemm_check("fp16x2", M, N, K, epi)
    finally:
        capi.check(lib.edsnet_debug_set_tc_variant(0))


@pytest.mark.parametrize("variant", [7, 8])
def test_to_out_read_out_variants_agree(variant):
    """to_out in the two-pass mode: the 16-column slab epilogue (7) and the eight-warp wide read-out (8) behind the
    profiling knob give the default (sixteen warps in pairs) kernel's scores up to the order of the LayerNorm row sums."""
    capi, lib = _lib()
    g, x, p = golden_case(FWD, CASES[0])
    scales = [int(s) for s in g["scales"]]
    model = make_model(p, scales, int(g["fc_depth"]), "fp16x2", DEV)
    xd = x[None].to(DEV)
    with torch.no_grad():
        cls0, loc0 = model(xd)
        capi.check(lib.edsnet_debug_set_tc_variant(variant))
        try:
            cls1, loc1 = model(xd)
        finally:
            capi.check(lib.edsnet_debug_set_tc_variant(0))
    assert lib.edsnet_debug_tc_status(1) == 0
    assert orc.rel_l2(cls1.cpu().numpy(), cls0.cpu().numpy()) < 2e-6
    assert orc.rel_l2(loc1.cpu().numpy(), loc0.cpu().numpy()) < 2e-6
    assert orc.rel_l2(loc1.cpu().numpy(), g["pred_loc"]) < TOL["fp16x2"]


# ------------------------------------------------------------------------------------------------ forward
@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16", "fp16x2"])
@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_golden(name, precision):
    g, x, p = golden_case(FWD, name)
    scales = [int(s) for s in g["scales"]]
    model = make_model(p, scales, int(g["fc_depth"]), precision, DEV)
    with torch.no_grad():
        cls, loc = model(x[None].to(DEV))
    _no_tc_timeout()
    assert cls.shape == g["pred_cls"].shape and loc.shape == g["pred_loc"].shape
    e_cls = orc.rel_l2(cls.cpu().numpy(), g["pred_cls"])
    e_loc = orc.rel_l2(loc.cpu().numpy(), g["pred_loc"])
    print(f"{name} {precision}: rel-l2 cls {e_cls:.2e} loc {e_loc:.2e}")
    assert e_cls < TOL[precision] and e_loc < TOL[precision], (e_cls, e_loc)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_forward_stages_match_oracle(precision):
    """Intermediates of the forward, read back from the workspace, against the oracle's stages."""
    capi, lib = _lib()
    g, x, p = golden_case(FWD, "T450_s4_8_16_32_d7")
    scales, depth = [int(s) for s in g["scales"]], int(g["fc_depth"])
    T = x.shape[0]
    stages = {}
    with torch.no_grad():
        orc.dsnet_forward(x, p, scales, depth, stages=stages)
    model = make_model(p, scales, depth, precision, DEV)
    with torch.no_grad():
        model(x[None].to(DEV))
    _no_tc_timeout()
    L = capi.WorkspaceLayout()
    lib.edsnet_workspace_bytes(model._config(), T, 1, C.byref(L))
    ws = model._workspace

    def view(off, shape):
        n = int(np.prod(shape))
        return ws[off:off + 4 * n].view(torch.float32).reshape(shape).cpu()

    pad = (64 - T % 64) % 64
    # qkv region is reused for LayerNorm output (yn), so only the per-head matrices and later stages are checked
    for name, key, shape in (("q_land", "q_land", (8, 64, 64)), ("k_land", "k_land", (8, 64, 64)),
                             ("attn2", "attn2", (8, 64, 64)), ("a3v", "a3v", (8, 64, 64)),
                             ("zmat", "pinv", (8, 64, 64))):
        got = view(getattr(L, name), shape)
        err = orc.rel_l2(got.numpy(), stages[key].numpy())
        print(name, err)
        # the Newton-Schulz chain amplifies rounding differences of an ill-conditioned attn2 (cond ~1e3..1e4)
        assert err < (2e-3 if name == "zmat" else 2e-5), (name, err)
    if precision == "fp32":
        # (tcgen05 precisions: merged leaves the value-convolution kernel as the to_out operand planes in a region
        # the LayerNorm planes overwrite later in the same forward; the stages below depend on it)
        merged = view(L.merged, (T, 512))
        assert orc.rel_l2(merged.numpy(), stages["merged"][pad:].numpy()) < 2e-5
    if precision == "fp32":
        u1 = view(L.u1, (T, 128))
        assert orc.rel_l2(u1.numpy(), stages["hidden"].numpy()) < 2e-5
    else:
        # tcgen05 precisions: the fc stack emits the three head projections per row instead of the hidden rows
        p32 = {k: v.float() for k, v in p.items()}
        hw = torch.cat([p32["fc_cls.0.weight"], p32["fc_loc.0.weight"]], 0)             # (3, 128)
        want = stages["hidden"] @ hw.t()
        got = view(L.u1, (T, 4))[:, :3]
        assert orc.rel_l2(got.numpy(), want.numpy()) < 2e-5
        # LayerNorm(1024) is folded into to_out's and fc1's epilogues: y never exists in fp32; fc1's output does
        with torch.no_grad():
            y = orc.nystrom_attention(x, p, orc.HEADS, None) + x
            u0 = orc.layer_norm(y, p["layer_norm.weight"], p["layer_norm.bias"]) @ p["fc1.weight"].t() + p["fc1.bias"]
        got0 = view(L.u0, (T, 128))
        err0 = orc.rel_l2(got0.numpy(), u0.numpy())
        print("fc1(LN(y))", err0)
        assert err0 < 2e-5
        # and the row statistics the fold uses: sum over the 16 column slots = (sum z, sum z^2), z = y - mean(x row)
        zs = view(L.zstat, (T, 16, 2)).double().sum(1)
        z = (y - x.mean(1, keepdim=True) - p["base_model.to_out.0.bias"].mean()).double()
        assert orc.rel_l2(zs[:, 0].numpy(), z.sum(1).numpy()) < 1e-4      # a cancelling sum: absolute accuracy
        assert orc.rel_l2(zs[:, 1].numpy(), (z * z).sum(1).numpy()) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_packed_batch_equals_per_video(precision):
    """Videos in one packed launch are independent (per-video pinv scale, no cross-video coupling)."""
    p = orc.synth_params(31, "xavier")
    scales, depth = [4, 8, 16, 32], 5
    lengths = [1, 37, 64, 65, 128, 200, 333, 800, 100, 129]
    xs = [orc.synth_features(t, 1000 + i) for i, t in enumerate(lengths)]
    model = make_model(p, scales, depth, precision, DEV)
    xp = torch.cat(xs).to(DEV)
    with torch.no_grad():
        cls_p, loc_p = model.forward_packed(xp, lengths)
        o = 0
        for x, t in zip(xs, lengths):
            c1, l1 = model(x[None].to(DEV))
            assert torch.equal(c1, cls_p[o:o + t]), "packed rows differ from the single-video call"
            assert torch.equal(l1, loc_p[o:o + t])
            with torch.no_grad():
                rc, rl = orc.dsnet_forward(x, p, scales, depth)
            assert orc.rel_l2(c1.cpu().numpy(), rc.numpy()) < TOL[precision]
            assert orc.rel_l2(l1.cpu().numpy(), rl.numpy()) < TOL[precision]
            o += t
    _no_tc_timeout()


def test_linearity_of_pool_and_heads_at_full_size():
    """Size-independent property at C5 size (T=16384, 4 scales): pred_loc is linear in the hidden sequence, so
    roi_pool_heads(a) + roi_pool_heads(b) - bias == roi_pool_heads(a + b)."""
    capi, lib = _lib()
    T, scales = 16384, [4, 8, 16, 32]
    p = orc.synth_params(5, "xavier")
    model = make_model(p, scales, 5, "fp32", DEV)
    batch = __import__("edsnet_b200").BatchPlan.build([T]).to(DEV)
    g = torch.Generator().manual_seed(3)
    a = torch.randn(T, 128, generator=g).to(DEV)
    b = torch.randn(T, 128, generator=g).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    w = model._weights(torch.device(DEV), st)
    outs = []
    for u in (a, b, a + b):
        cls = torch.empty(T, 4, device=DEV)
        loc = torch.empty(T, 4, 2, device=DEV)
        capi.check(lib.edsnet_roi_pool_heads(model._config(), w, batch.struct, u.data_ptr(), cls.data_ptr(),
                                             loc.data_ptr(), st))
        outs.append(loc)
    bias = p["fc_loc.0.bias"].to(DEV)
    lhs = outs[0] + outs[1] - bias
    assert float((lhs - outs[2]).abs().max()) < 1e-4
    # and against the oracle's pooling on the same data
    pooled = orc.roi_pool_direct((a + b).cpu(), scales)
    ref = pooled @ p["fc_loc.0.weight"].t() + p["fc_loc.0.bias"]
    assert orc.rel_l2(outs[2].cpu().numpy(), ref.numpy()) < 1e-5


def test_long_video_forward_c5_shape():
    """C5 shape (T=16384, scales [4,8,16,32]) against the oracle (a few seconds of CPU)."""
    T, scales, depth = 16384, [4, 8, 16, 32], 5
    x = orc.synth_features(T, 77)
    p = orc.synth_params(78, "xavier")
    with torch.no_grad():
        rc, rl = orc.dsnet_forward(x, p, scales, depth)
    for precision in ("fp32", "fp16x3"):
        model = make_model(p, scales, depth, precision, DEV)
        with torch.no_grad():
            cls, loc = model(x[None].to(DEV))
        _no_tc_timeout()
        e = (orc.rel_l2(cls.cpu().numpy(), rc.numpy()), orc.rel_l2(loc.cpu().numpy(), rl.numpy()))
        print("C5", precision, e)
        assert max(e) < 5e-5, e       # fp32 reference itself sits ~1e-5 from fp64 at this length


def test_few_long_videos_use_key_and_row_range_splits():
    """A batch of few videos makes a3v stream its keys in several ranges per (video, head) (partials merged flash-decoding
    style) and attn_out take row ranges: ragged lengths, a range count that does not divide the tile count, a video
    shorter than one range.  Same rows as the one-video-per-call path, and the oracle."""
    from edsnet_b200 import BatchPlan
    scales = [4, 8]
    p = orc.synth_params(33, "xavier")
    lengths = [5000, 130, 64, 1999]
    xs = [orc.synth_features(t, 700 + i) for i, t in enumerate(lengths)]
    model = make_model(p, scales, 5, "fp16x3", DEV)
    with torch.no_grad():
        cls, loc = model.forward_packed(torch.cat(xs).to(DEV), BatchPlan.build(lengths).to(DEV))
    _no_tc_timeout()
    o = 0
    for x, t in zip(xs, lengths):
        with torch.no_grad():
            rc, rl = orc.dsnet_forward(x, p, scales, 5)
        e = max(orc.rel_l2(cls[o:o + t].cpu().numpy(), rc.numpy()), orc.rel_l2(loc[o:o + t].cpu().numpy(), rl.numpy()))
        assert e < 3e-5, (t, e)          # fp32 reference itself is ~1e-5 from float64 at T = 5000
        o += t


# ------------------------------------------------------------------------------------------------ decode + NMS
@pytest.mark.parametrize("name", list(DN["cases"]))
def test_decode_nms_bit_exact_vs_reference_golden(name):
    from edsnet_b200 import BatchPlan
    T = int(DN[f"{name}/T"])
    scales = [int(s) for s in DN[f"{name}/scales"]]
    S = len(scales)
    p = orc.synth_params(1, "default")
    model = make_model(p, scales, 5, "fp32", DEV)
    batch = BatchPlan.build([T]).to(DEV)
    loc = torch.from_numpy(DN[f"{name}/loc"]).to(DEV).reshape(T, S, 2)
    scores = torch.from_numpy(DN[f"{name}/scores"]).to(DEV).reshape(T, S)
    bf, bi = model.decode_packed(loc, batch)
    bf, bi = bf.cpu().numpy(), bi.cpu().numpy()
    # bit-exact against the oracle with the correctly rounded exp (machine independent) ...
    want = orc.decode_boxes(DN[f"{name}/loc"], T, scales, exp_mode="cr")
    assert np.array_equal(bf, want), "float32 left/right boxes differ from the oracle"
    assert np.array_equal(bi, orc.clip_round(want, T))
    # ... and within NumPy's own float32-exp error (<= 2.5 ulp of the width) of the reference run
    assert np.abs(bf - DN[f"{name}/lr"]).max() <= 4e-6 * np.abs(DN[f"{name}/lr"]).max()
    # the correctly rounded exp reproduces the reference run's int32 boxes on every golden anchor (no edge of these cases
    # lies within NumPy's 2.5-ulp exp error of k + 0.5): hard requirement, so the NMS comparison below is unconditional
    same_int = np.array_equal(bi, DN[f"{name}/boxes_i32"])
    assert same_int, f"{name}: int32 boxes differ from the reference run"
    sc = DN[f"{name}/scores"]
    for thresh, suffix in ((0.5, ""), (0.3, "_t03")):
        r = model.nms_packed(scores, loc, batch, thresh)
        k = int(r["keep_count"].cpu()[0])
        ks, kb = r["keep_scores"][:k].cpu().numpy(), r["keep_boxes"][:k].cpu().numpy()
        rs, rb, _ = orc.nms_1d(sc, bi, thresh)
        assert np.array_equal(ks, rs) and np.array_equal(kb, rb)
        if same_int:       # identical integer boxes => must reproduce the reference's bbox_helper.nms output
            assert np.array_equal(ks, DN[f"{name}/keep_scores{suffix}"])
            assert np.array_equal(kb, DN[f"{name}/keep_boxes{suffix}"])


@pytest.mark.parametrize("name", CASES)
def test_predict_and_proposals_match_reference_golden(name):
    """fp32 path end to end: predict() boxes and post-NMS proposals vs the reference's.  Floating-point scores feed
    an integer pipeline, so exact equality needs the same rounding of every box edge; compare exactly where the
    decoded int boxes agree and require that to be (nearly) everywhere."""
    g, x, p = golden_case(FWD, name)
    T, scales = int(g["T"]), [int(s) for s in g["scales"]]
    model = make_model(p, scales, int(g["fc_depth"]), "fp32", DEV)
    with torch.no_grad():               # as evaluate.py:17 / infer.py:27 call it
        scores, boxes = model.predict(x[None].to(DEV))
    assert scores.shape == (T * len(scales),) and boxes.shape == (T * len(scales), 2)
    assert scores.dtype == np.float32 and boxes.dtype == np.float32
    assert np.abs(boxes - g["boxes_f32"]).max() < 1e-3 + 2e-6 * np.abs(g["boxes_f32"]).max()
    assert np.array_equal(boxes, orc.decode_boxes(model(x[None].to(DEV))[1].detach().cpu().numpy(), T, scales,
                                                  exp_mode="cr"))
    ib = orc.clip_round(boxes, T)
    frac_same = float((ib == g["boxes_i32"]).all(axis=1).mean())
    assert frac_same > 0.995
    with torch.no_grad():
        ks, kb = model.proposals(x[None].to(DEV), 0.5)
    # oracle NMS on OUR scores/boxes must agree bit for bit with the device NMS
    rs, rb, _ = orc.nms_1d(scores, ib, 0.5)
    assert np.array_equal(ks, rs) and np.array_equal(kb, rb)
    same = len(kb) == len(g["keep_boxes"]) and np.array_equal(kb, g["keep_boxes"])
    ref_scores = np.asarray(g["pred_cls"]).reshape(-1)
    same_order = np.array_equal(np.argsort(scores, kind="stable"), np.argsort(ref_scores, kind="stable"))
    _record_parity("predict_proposals", name, {"T": T, "scales": scales, "frac_int_boxes_equal_reference": frac_same,
                                               "score_order_equal_reference": bool(same_order),
                                               "kept_set_equal_reference": bool(same), "kept": int(len(kb)),
                                               "kept_reference": int(len(g["keep_boxes"]))})
    if frac_same == 1.0 and same_order:
        # identical integer boxes visited in the identical order: greedy NMS is then a pure integer function of its
        # input, so the kept boxes MUST equal what the reference's bbox_helper.nms returned
        assert same, f"{name}: kept set differs from the reference although boxes and score order agree"


def test_nms_packed_many_videos_vs_oracle():
    from edsnet_b200 import BatchPlan
    rng = np.random.default_rng(7)
    scales = [4, 8, 16, 32]
    S = 4
    lengths = [int(t) for t in rng.integers(1, 900, size=40)] + [1500]       # 1500*4 > 4096: scratch path
    p = orc.synth_params(1, "default")
    model = make_model(p, scales, 5, "fp32", DEV)
    batch = BatchPlan.build(lengths).to(DEV)
    R = sum(lengths)
    loc = np.stack([rng.normal(0, 0.6, R * S), rng.normal(0, 0.5, R * S)], 1).astype(np.float32)
    scores = rng.random(R * S).astype(np.float32)
    scores[::17] = scores[5]                                                 # plenty of exact ties
    r = model.nms_packed(torch.from_numpy(scores).to(DEV).reshape(R, S),
                         torch.from_numpy(loc).to(DEV).reshape(R, S, 2), batch, 0.5)
    counts = r["keep_count"].cpu().numpy()
    ks, kb, ki = r["keep_scores"].cpu().numpy(), r["keep_boxes"].cpu().numpy(), r["keep_idx"].cpu().numpy()
    o = 0
    for v, t in enumerate(lengths):
        sl = slice(o * S, (o + t) * S)
        boxes = orc.clip_round(orc.decode_boxes(loc[sl], t, scales, exp_mode="cr"), t)
        rs, rb, ridx = orc.nms_1d(scores[sl], boxes, 0.5)
        c = int(counts[v])
        assert c == len(rs), (v, t, c, len(rs))
        assert np.array_equal(ks[o * S:o * S + c], rs)
        assert np.array_equal(kb[o * S:o * S + c], rb)
        assert np.array_equal(ki[o * S:o * S + c], ridx.astype(np.int32))
        o += t


def test_nms_large_deep_suppression_chain():
    """Worst case for the fixpoint NMS of the > 4096-anchor path: identical boxes shifted by one position with scores that
    fall along the timeline form ONE suppression chain thousands of levels deep (box t is decided only after box t - 1 ...),
    far beyond the fixed number of parallel rounds -- the ordered clean-up pass must finish it, bit-exact."""
    from edsnet_b200 import BatchPlan
    T, scales = 6000, [10]
    p = orc.synth_params(1, "default")
    model = make_model(p, scales, 5, "fp32", DEV)
    batch = BatchPlan.build([T]).to(DEV)
    loc = np.zeros((T, 1, 2), dtype=np.float32)                      # every box = its anchor: [t - 5, t + 5) clipped
    scores = np.linspace(0.99, 0.01, T).astype(np.float32)
    for thresh in (0.5, 0.3):
        r = model.nms_packed(torch.from_numpy(scores).to(DEV).reshape(T, 1), torch.from_numpy(loc).to(DEV), batch, thresh)
        k = int(r["keep_count"].cpu()[0])
        boxes = orc.clip_round(orc.decode_boxes(loc, T, scales, exp_mode="cr"), T)
        rs, rb, ridx = orc.nms_1d(scores, boxes, thresh)
        assert k == len(rs) and k > 500
        assert np.array_equal(r["keep_idx"][:k].cpu().numpy(), ridx.astype(np.int32))
        assert np.array_equal(r["keep_boxes"][:k].cpu().numpy(), rb) and np.array_equal(r["keep_scores"][:k].cpu().numpy(), rs)


def test_nms_properties_at_full_size():
    """C5-size NMS (65536 anchors): kept boxes are mutually below the threshold, scores descend, every dropped
    valid box is suppressed by some kept box of higher-or-equal score, and NMS of the kept set is the identity."""
    from edsnet_b200 import BatchPlan
    rng = np.random.default_rng(11)
    T, scales, S = 16384, [4, 8, 16, 32], 4
    N = T * S
    p = orc.synth_params(1, "default")
    model = make_model(p, scales, 5, "fp32", DEV)
    batch = BatchPlan.build([T]).to(DEV)
    loc = np.stack([rng.normal(0, 0.6, N), rng.normal(0, 0.5, N)], 1).astype(np.float32)
    scores = (rng.permutation(N).astype(np.float32) + 1) / np.float32(N)
    r = model.nms_packed(torch.from_numpy(scores).to(DEV).reshape(T, S),
                         torch.from_numpy(loc).to(DEV).reshape(T, S, 2), batch, 0.5)
    k = int(r["keep_count"].cpu()[0])
    ks, kb = r["keep_scores"][:k].cpu().numpy(), r["keep_boxes"][:k].cpu().numpy().astype(np.int64)
    ki = r["keep_idx"][:k].cpu().numpy()
    assert k > 1000 and np.all(np.diff(ks) < 0)
    boxes = orc.clip_round(orc.decode_boxes(loc, T, scales, exp_mode="cr"), T)
    assert np.array_equal(boxes[ki], kb) and np.array_equal(scores[ki], ks)
    lo, hi = kb[:, 0], kb[:, 1]
    for i in range(0, k, max(1, k // 200)):                 # sampled pairwise check of the kept set
        inter = np.maximum(0, np.minimum(hi, hi[i]) - np.maximum(lo, lo[i]))
        hull = np.maximum(hi, hi[i]) - np.minimum(lo, lo[i])
        iou = inter / hull
        iou[i] = 0
        assert (iou < 0.5).all()
    # every dropped valid box overlaps >= thresh with a kept box of higher score (sampled)
    kept_mask = np.zeros(N, bool)
    kept_mask[ki] = True
    dropped = np.nonzero(~kept_mask & (boxes[:, 0] < boxes[:, 1]))[0]
    for j in dropped[:: max(1, len(dropped) // 300)]:
        higher = ks > scores[j]
        inter = np.maximum(0, np.minimum(hi[higher], boxes[j, 1]) - np.maximum(lo[higher], boxes[j, 0]))
        hull = np.maximum(hi[higher], boxes[j, 1]) - np.minimum(lo[higher], boxes[j, 0])
        assert (inter / hull >= 0.5).any()
    # full oracle NMS at this size takes a few seconds: exact comparison
    rs, rb, ridx = orc.nms_1d(scores, boxes, 0.5)
    assert np.array_equal(rs, ks) and np.array_equal(rb, kb) and np.array_equal(ridx, ki)


# ------------------------------------------------------------------------------------------------ error surface
def test_error_surface():
    from edsnet_b200 import DSNet, _capi
    p = orc.synth_params(1, "default")
    model = make_model(p, [4], 2, "fp32", DEV)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 8, 1024))                      # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 8, 1024, device=DEV))          # batch > 1, as the reference's view() fails
    odd = DSNet("nystromformer", 1024, 128, [5], 8, pooling_type="roi").to(DEV).eval()
    with pytest.raises(RuntimeError):
        odd(torch.zeros(1, 8, 1024, device=DEV))            # odd scale, as the reference's view() fails
    lib = _capi.lib()
    assert lib.edsnet_forward(None, None, None, None, None, None, None, 0, None) == _capi.E_ARG
    assert "config" in _capi.last_error()
    # the kernels read x with 32-byte vector loads: a misaligned view is refused, not faulted on
    buf = torch.zeros(8 * 1024 + 1, device=DEV)
    with pytest.raises(RuntimeError, match="aligned"):
        make_model(p, [4], 2, "fp16x2", DEV)(buf[1:].view(1, 8, 1024))


def test_pipeline_from_host_buffers():
    from edsnet_b200 import ScoringPipeline
    p = orc.synth_params(9, "xavier")
    scales = [4, 8]
    model = make_model(p, scales, 5, "fp32", DEV)
    lengths = [50, 300, 64, 129, 700, 33, 256, 90]
    xs = [orc.synth_features(t, 500 + i) for i, t in enumerate(lengths)]
    xh = torch.cat(xs).pin_memory()
    pipe = ScoringPipeline(model, chunk_rows=512)
    kc, ks, kb, cu = pipe.run(xh, lengths)
    assert pipe.h2d_bytes >= xh.numel() * 4 and pipe.d2h_bytes > 0
    for v, (x, t) in enumerate(zip(xs, lengths)):
        es, eb = model.proposals(x[None].to(DEV), 0.5)
        o, c = int(cu[v]) * 2, int(kc[v])
        assert np.array_equal(ks[o:o + c].numpy(), es) and np.array_equal(kb[o:o + c].numpy(), eb)


def test_gradients_match_the_oracle_graph():
    """eval()-mode call with grad enabled: values from the kernels, gradients of a cls+loc loss must equal the
    gradients of the oracle's torch graph (fp32, CPU) -- the parity check config 3's all-reduce builds on."""
    p = orc.synth_params(21, "xavier")
    scales, depth, T = [4, 8], 3, 90
    x = orc.synth_features(T, 5)
    g = torch.Generator().manual_seed(0)
    wc = torch.randn(T, 2, generator=g)
    wl = torch.randn(T, 2, 2, generator=g)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rc, rl = orc.dsnet_forward(x, pr, scales, depth)
    ((rc * wc).sum() + (rl * wl).sum()).backward()
    model = make_model(p, scales, depth, "fp32", DEV)
    cls, loc = model(x[None].to(DEV))
    assert cls.requires_grad and orc.rel_l2(cls.detach().cpu().numpy(), rc.detach().numpy()) < 1e-5
    ((cls * wc.to(DEV)).sum() + (loc * wl.to(DEV)).sum()).backward()
    names = {"base_model.to_qkv.weight": model.base_model.to_qkv.weight, "fc1.weight": model.fc1.weight,
             "fc_block.0.weight": model.fc_block[0].weight, "fc_loc.0.weight": model.fc_loc[0].weight,
             "base_model.res_conv.weight": model.base_model.res_conv.weight, "layer_norm.bias": model.layer_norm.bias}
    for k, q in names.items():
        e = orc.rel_l2(q.grad.cpu().numpy(), pr[k].grad.numpy())
        assert e < 2e-3, (k, e)          # fp32 backward through the pinv chain: CPU vs GPU summation order
    # train(): dropout active, outputs differ from eval and carry a graph
    model.train()
    c2, _ = model(x[None].to(DEV))
    assert c2.requires_grad and not torch.equal(c2.detach(), cls.detach())


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("rows,depth", [(1, 1), (129, 5), (1000, 7), (40000, 5)])
def test_fc_stack_stage(precision, rows, depth):
    """The shared fc block (anchor_based/dsnet.py:91-96,107-108) alone: CUDA-core and tcgen05 versions vs fp64."""
    capi, lib = _lib()
    p = orc.synth_params(41, "xavier")
    model = make_model(p, [4], depth, precision, DEV)
    g = torch.Generator().manual_seed(rows)
    u = torch.randn(rows, 128, generator=g) * 1.5
    ref = u.double()
    W, b = p["fc_block.0.weight"].double(), p["fc_block.0.bias"].double()
    gw, gb = p["fc_block.3.weight"].double(), p["fc_block.3.bias"].double()
    for _ in range(depth):
        ref = orc.layer_norm(torch.relu(ref @ W.t() + b), gw, gb)
    ud = u.to(DEV)
    out = torch.full((rows, 128), float("nan"), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    w = model._weights(torch.device(DEV), st)
    capi.check(lib.edsnet_fc_stack(model._config(), w, ud.data_ptr(), out.data_ptr(), rows, st))
    _no_tc_timeout()
    err = orc.rel_l2(out.cpu().numpy(), ref.numpy())
    print(precision, rows, depth, err)
    assert err < 3e-6, err


def test_training_step_single_gpu():
    """Config-3 step on one GPU (world_size 1) through the reference-shaped loop (DataParallelStep: model(x) in train()
    mode, torch losses, loss.backward() into the backward kernels, flat-bucket reduction, torch Adam): labels from a
    synthetic keyshot mask.  With Dropout switched off the loss must go down step after step; with the reference's
    Dropout(0.5) the steps stay finite and the updated weights reach the inference kernels."""
    from edsnet_b200 import training as tr
    scales = [4, 8, 16, 32]
    rng = np.random.default_rng(3)
    seqs, cls_l, loc_l = [], [], []
    for i, T in enumerate((120, 200)):
        mask = np.zeros(T, bool)
        mask[20:45] = True
        mask[90:110] = True
        c, l = tr.anchor_labels(mask, scales, rng)
        seqs.append(orc.synth_features(T, 900 + i).to(DEV))
        cls_l.append(torch.from_numpy(c).to(DEV))
        loc_l.append(torch.from_numpy(l).float().to(DEV))
    for p_drop in (0.0, 0.5):
        model = make_model(orc.synth_params(51, "xavier"), scales, 5, "fp16x3", DEV)
        model.fc_block[2].p = p_drop
        stepper = tr.DataParallelStep(model, lr=2e-4, world_size=1)
        torch.manual_seed(0)
        losses = [stepper.step(seqs, cls_l, loc_l) for _ in range(12)]
        print(p_drop, losses)
        assert all(np.isfinite(losses))
        if p_drop == 0.0:
            assert np.mean(losses[-3:]) < np.mean(losses[:3]) and losses[-1] < losses[0]
        # the updated weights flow back into the kernel path (weight cache keyed on parameter versions)
        model.eval()
        with torch.no_grad():
            c1, _ = model(seqs[0][None])
        assert torch.isfinite(c1).all()


# ------------------------------------------------------------------------------------------------ config 4: full MHA base
MHA = load_npz("forward_mha_golden.npz")


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("name", list(MHA["forward_cases"]))
def test_mha_base_matches_reference_golden(name, precision):
    """DSNet(base_model='attention') (modules/models.py:12-74), the T=2048 comparison config: flash attention without
    the (8, T, T) score tensor the reference materialises."""
    g = {k.split("/", 1)[1]: MHA[k] for k in MHA.files if k.startswith(name + "/")}
    x = orc.synth_features(int(g["T"]), int(g["x_seed"]))
    p = orc.synth_params_mha(int(g["w_seed"]), str(g["init"]))
    scales = [int(s) for s in g["scales"]]
    model = make_model(p, scales, int(g["fc_depth"]), precision, DEV, base="attention")
    with torch.no_grad():
        cls, loc = model(x[None].to(DEV))
    _no_tc_timeout()
    e = (orc.rel_l2(cls.cpu().numpy(), g["pred_cls"]), orc.rel_l2(loc.cpu().numpy(), g["pred_loc"]))
    print(name, precision, e)
    assert max(e) < TOL[precision], e


def test_mha_base_packed_and_gradients():
    p = orc.synth_params_mha(61, "xavier")
    scales = [4, 8]
    model = make_model(p, scales, 3, "fp32", DEV, base="attention")
    lengths = [50, 130, 64]
    xs = [orc.synth_features(t, 700 + i) for i, t in enumerate(lengths)]
    with torch.no_grad():
        cp, lp = model.forward_packed(torch.cat(xs).to(DEV), lengths)
    o = 0
    for x, t in zip(xs, lengths):
        with torch.no_grad():
            rc, rl = orc.dsnet_forward(x, p, scales, 3, base="attention")
        assert orc.rel_l2(cp[o:o + t].cpu().numpy(), rc.numpy()) < 1e-5
        assert orc.rel_l2(lp[o:o + t].cpu().numpy(), rl.numpy()) < 1e-5
        o += t
    cls, loc = model(xs[0][None].to(DEV))                      # eval + grad enabled: kernel values, torch-graph gradients
    (cls.sum() + loc.sum()).backward()
    assert model.base_model.Q.weight.grad is not None and torch.isfinite(model.base_model.Q.weight.grad).all()


# ------------------------------------------------------------------------------------------------ keyshot summary (f-1)
def _shots_for(T, rng, shot_lo=3, shot_hi=12, rate=15):
    """Synthetic shot structure shaped like the reference's h5 records (tests/test_train.py:16-45): picks every
    `rate` frames, contiguous shots covering all frames."""
    n_frames = T * rate - int(rng.integers(0, rate))
    bounds = [0]
    while bounds[-1] < n_frames:
        bounds.append(min(n_frames, bounds[-1] + int(rng.integers(shot_lo, shot_hi + 1)) * rate // 2))
    cps = np.stack([bounds[:-1], np.asarray(bounds[1:]) - 1], 1).astype(np.int32)
    nfps = (cps[:, 1] - cps[:, 0] + 1).astype(np.int32)
    picks = (np.arange(T) * rate).astype(np.int32)
    picks = picks[picks < n_frames]
    assert len(picks) == T
    return dict(cps=cps, nfps=nfps, picks=picks, n_frames=n_frames)


@pytest.mark.parametrize("name", list(DN["cases"]))
def test_keyshot_summary_matches_reference_golden(name):
    """The reference's own bbox2summary output (knapsack = the exact DP both sides share) on the reference's kept boxes."""
    from edsnet_b200 import BatchPlan, ShotPlan, keyshot_summaries, split_summaries
    T = int(DN[f"{name}/T"])
    scales = [int(s) for s in DN[f"{name}/scales"]]
    S = len(scales)
    model = make_model(orc.synth_params(1, "default"), scales, 5, "fp32", DEV)
    batch = BatchPlan.build([T]).to(DEV)
    ks, kb = DN[f"{name}/keep_scores"], DN[f"{name}/keep_boxes"]
    n = T * S
    nms = {"keep_count": torch.tensor([len(ks)], dtype=torch.int32, device=DEV),
           "keep_scores": torch.zeros(n, device=DEV), "keep_boxes": torch.zeros((n, 2), dtype=torch.int32, device=DEV)}
    nms["keep_scores"][:len(ks)] = torch.from_numpy(ks).to(DEV)
    nms["keep_boxes"][:len(ks)] = torch.from_numpy(kb).to(DEV)
    vd = dict(cps=DN[f"{name}/cps"], nfps=DN[f"{name}/nfps"], picks=DN[f"{name}/picks"], n_frames=T * 15)
    shots = ShotPlan([vd], DEV)
    out = keyshot_summaries(model, nms, batch, shots)
    got = split_summaries(out["summary"], shots)[0]
    assert np.array_equal(np.packbits(got), DN[f"{name}/summary"])


def test_keyshot_summary_packed_vs_oracle():
    """features -> proposals -> summaries for a packed batch, every stage on the device, against the oracle's host
    restatement fed with the same kept proposals (bit-exact: integer shot scores, same exact knapsack)."""
    from edsnet_b200 import BatchPlan, ShotPlan, keyshot_summaries, split_summaries
    rng = np.random.default_rng(5)
    p = orc.synth_params(71, "xavier")
    scales = [4, 8, 16, 32]
    model = make_model(p, scales, 5, "fp16x3", DEV)
    lengths = [int(t) for t in rng.integers(20, 700, size=12)] + [1, 64]
    xs = [orc.synth_features(t, 3000 + i) for i, t in enumerate(lengths)]
    vds = [_shots_for(t, rng) for t in lengths]
    batch = BatchPlan.build(lengths).to(DEV)
    shots = ShotPlan(vds, DEV)
    with torch.no_grad():
        cls, loc = model.forward_packed(torch.cat(xs).to(DEV), batch)
        nms = model.nms_packed(cls, loc, batch, 0.5)
    out = keyshot_summaries(model, nms, batch, shots)
    torch.cuda.synchronize()
    got = split_summaries(out["summary"], shots)
    counts = nms["keep_count"].cpu().numpy()
    ks, kb = nms["keep_scores"].cpu().numpy(), nms["keep_boxes"].cpu().numpy()
    seg = out["seg_scores"].cpu().numpy()
    o = 0
    for v, (t, vd) in enumerate(zip(lengths, vds)):
        a, c = o * 4, int(counts[v])
        want = orc.bbox_summary(t, ks[a:a + c], kb[a:a + c], vd["cps"], vd["n_frames"], vd["nfps"], vd["picks"])
        assert np.array_equal(got[v], want), v
        assert 0 < got[v].sum() <= int(vd["n_frames"] * 0.15) or got[v].sum() == 0
        o += t
    assert seg.max() <= 1000 and seg.min() >= 0


def test_keyshot_summary_long_uneven_shots():
    """Shots of 15 ... 1320 frames in one video (what KTS produces on real videos): the NumPy-exact pairwise shot means
    split ranges above 128 frames recursively; the device does that on an explicit stack (a device-recursive version
    crashed on such tables)."""
    from edsnet_b200 import BatchPlan, ShotPlan, keyshot_from_scores
    model = make_model(orc.synth_params(3, "default"), [4, 8], 5, "fp32", DEV)
    T, nf = 150, 2243
    g = np.random.default_rng(1).random(T).astype(np.float32)
    for cps in ([[0, 209], [210, 404], [405, 419], [420, 539], [540, 1739], [1740, 2242]],
                [[0, 209], [210, 404], [405, 419], [420, 1739], [1740, 2242]],
                [[0, 209], [210, 419], [420, 539], [540, 1739], [1740, 1769], [1770, 2242]], [[0, 2242]]):
        c = np.asarray(cps)
        w = c[:, 1] - c[:, 0] + 1
        vd = dict(cps=c, nfps=w, picks=np.arange(T) * 15, n_frames=nf)
        out = keyshot_from_scores(model, torch.from_numpy(g).to(DEV), BatchPlan.build([T]).to(DEV), ShotPlan([vd], DEV))
        torch.cuda.synchronize()
        want = orc.keyshot_summary(g, c, nf, w, vd["picks"])
        assert np.array_equal(out["summary"].cpu().numpy().astype(bool), want)


def test_training_targets_from_gtscore_vs_oracle():
    """anchor_based/train.py:79-84 on the device: get_keyshot_summ on ground-truth scores (no proposals) for a packed
    split, then downsample_summ -- bit-exact against the oracle's host restatement."""
    from edsnet_b200 import ShotPlan, training_targets
    rng = np.random.default_rng(21)
    model = make_model(orc.synth_params(3, "default"), [4, 8], 5, "fp32", DEV)
    lengths = [int(t) for t in rng.integers(30, 600, size=10)] + [1, 17]
    vds = [_shots_for(t, rng) for t in lengths]
    gts = [rng.random(t).astype(np.float32) for t in lengths]
    shots = ShotPlan(vds, DEV)
    got = training_targets(model, gts, shots, DEV)
    for t, vd, g, m in zip(lengths, vds, gts, got):
        want = orc.downsample_summ(orc.keyshot_summary(g, vd["cps"], vd["n_frames"], vd["nfps"], vd["picks"]))
        assert m.shape == (t,) and np.array_equal(m, want)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_forward_stays_inside_its_buffers(precision):
    """edsnet_forward called through the C ABI with guard bands around the workspace (sized EXACTLY by
    edsnet_workspace_bytes), the input and both outputs: no byte outside the declared extents may change, for ragged
    packed batches whose tiles end in partial rows."""
    capi, lib = _lib()
    p = orc.synth_params(13, "xavier")
    scales = [4, 8, 16, 32]
    model = make_model(p, scales, 5, precision, DEV)
    from edsnet_b200 import BatchPlan
    G = 1 << 16                                          # guard bytes on either side
    for lengths in ([1], [129, 64, 1, 37], [300, 255, 2]):
        R, S = sum(lengths), len(scales)
        batch = BatchPlan.build(lengths).to(DEV)
        cfg = model._config()
        st = torch.cuda.current_stream().cuda_stream
        w = model._weights(torch.device(DEV), st)
        need = lib.edsnet_workspace_bytes(cfg, R, len(lengths), None)
        sizes = {"ws": need, "x": R * 4096, "cls": R * S * 4, "loc": R * S * 8}
        bufs = {k: torch.full((n + 2 * G,), 0x5A, dtype=torch.uint8, device=DEV) for k, n in sizes.items()}
        x = torch.cat([orc.synth_features(t, 40 + i) for i, t in enumerate(lengths)]).to(DEV)
        bufs["x"][G:G + R * 4096] = x.view(-1).view(torch.uint8)
        ptr = {k: v.data_ptr() + G for k, v in bufs.items()}
        assert all(pp % 256 == 0 for pp in ptr.values())
        capi.check(lib.edsnet_forward(cfg, w, batch.struct, ptr["x"], ptr["cls"], ptr["loc"], ptr["ws"], need, st))
        _no_tc_timeout()
        for k, v in bufs.items():
            n = sizes[k]
            assert bool((v[:G] == 0x5A).all()) and bool((v[G + n:] == 0x5A).all()), f"{k}: write outside the buffer"
        assert torch.equal(bufs["x"][G:G + R * 4096].view(torch.float32).view(R, 1024), x), "input modified"
        cls = bufs["cls"][G:G + R * S * 4].view(torch.float32).view(R, S)
        ref_cls, _ = model.forward_packed(x, lengths)
        assert torch.equal(cls, ref_cls)


def test_graphed_forward_replays_identical_outputs():
    """CUDA-graph replay of the forward for a fixed shape: bit-identical to the eager call, for fresh inputs."""
    p = orc.synth_params(8, "xavier")
    model = make_model(p, [12], 5, "fp16x3", DEV)
    run = model.graphed_forward([320])
    for seed in (1, 2, 3):
        x = orc.synth_features(320, seed).to(DEV)
        with torch.no_grad():
            c0, l0 = model(x[None])
        c1, l1 = run(x[None])
        torch.cuda.synchronize()
        assert torch.equal(c0, c1) and torch.equal(l0, l1)
    _no_tc_timeout()


@pytest.mark.parametrize("scale", [1e-4, 30.0, 1e3])
def test_input_magnitude_robustness(scale):
    """Features scaled over seven orders of magnitude (the plane scales of the tensor-core path are partly taken from
    bounds, not maxima): the fp16x3 path stays within a small multiple of the error floor the reference's own fp32
    arithmetic has against float64 on the same input."""
    p = orc.synth_params(5, "xavier")
    scales = [4, 8, 16, 32]
    x = orc.synth_features(100, 77) * scale
    with torch.no_grad():
        c64, l64 = orc.dsnet_forward(x.double(), {k: v.double() for k, v in p.items()}, scales, 5)
        c32, l32 = orc.dsnet_forward(x, p, scales, 5)
    floor = max(orc.rel_l2(l32.numpy(), l64.numpy()), orc.rel_l2(c32.numpy(), c64.numpy()))
    model = make_model(p, scales, 5, "fp16x3", DEV)
    with torch.no_grad():
        c, l = model(x[None].to(DEV))
    _no_tc_timeout()
    e_cls = orc.rel_l2(c.cpu().numpy(), c64.numpy())
    e_loc = orc.rel_l2(l.cpu().numpy(), l64.numpy())
    print(scale, e_cls, e_loc, floor)
    # measured: 1e-4 and 30: at the floor; 1e3 (saturated softmaxes, the reference's own fp32 is 1.3e-5 off float64):
    # 9.3e-5 = 7 x floor with 32 truncating MMA steps per main accumulator (to_qkv, K = 1024), 6.5e-5 with 22
    assert max(e_cls, e_loc) <= 8 * floor + 5e-6


# ------------------------------------------------------------------------------------------------ evaluation metrics
EV = load_npz("eval_golden.npz")


def _eval_case(name):
    n_frames, nu = int(EV[f"{name}/n_frames"]), int(EV[f"{name}/users_frames"])
    pred = np.unpackbits(EV[f"{name}/pred"])[:n_frames]
    users = np.unpackbits(EV[f"{name}/users"], axis=1)[:, :nu]
    return int(EV[f"{name}/T"]), n_frames, pred, users, str(EV[f"{name}/metric"]), int(EV[f"{name}/x_seed"])


def test_eval_metrics_match_reference_golden():
    """F-score (bit-exact float64) and diversity of the reference's get_summ_f1score / get_summ_diversity, all golden
    cases in ONE packed launch (cut and zero-padded predictions, avg and max, empty and single-row selections)."""
    from edsnet_b200 import BatchPlan, ShotPlan, TruthPlan, eval_metrics
    names = [str(n) for n in EV["cases"]]
    cases = [_eval_case(n) for n in names]
    lengths = [c[0] for c in cases]
    x = torch.cat([orc.synth_features(c[0], c[5]) for c in cases]).to(DEV)
    batch = BatchPlan.build(lengths).to(DEV)
    vds = [dict(cps=np.array([[0, c[1] - 1]]), nfps=np.array([c[1]]), picks=np.arange(c[0]) * 15, n_frames=c[1]) for c in cases]
    shots = ShotPlan(vds, DEV)
    truth = TruthPlan([c[3] for c in cases], [c[4] for c in cases], DEV)
    summary = torch.from_numpy(np.concatenate([c[2] for c in cases]).astype(np.uint8)).to(DEV)
    out = eval_metrics(x, batch, shots, truth, summary)
    torch.cuda.synchronize()
    f, d = out["fscore"].cpu().numpy(), out["diversity"].cpu().numpy()
    uf = out["user_f1"].cpu().numpy()
    o = 0
    for i, n in enumerate(names):
        assert f[i] == float(EV[f"{n}/fscore"]), (n, f[i], float(EV[f"{n}/fscore"]))
        want_u = EV[f"{n}/user_f1"]
        assert np.array_equal(uf[o:o + len(want_u)], want_u), n
        o += len(want_u)
        want_d = float(EV[f"{n}/diversity"])
        assert abs(d[i] - want_d) <= 1e-5 * max(abs(want_d), 1e-12), (n, d[i], want_d)


def test_evaluate_loop_vs_oracle():
    """edsnet_b200.evaluate (the reference's evaluate() over a loader of 8-tuples) against the oracle's host chain fed
    with the device's kept proposals: F-scores identical, diversity to 1e-5."""
    from edsnet_b200 import evaluate
    from edsnet_b200 import BatchPlan, ShotPlan, TruthPlan, eval_metrics, keyshot_summaries, split_summaries
    rng = np.random.default_rng(11)
    p = orc.synth_params(5, "xavier")
    scales = [4, 8, 16, 32]
    model = make_model(p, scales, 5, "fp16x3", DEV)
    lengths = [int(t) for t in rng.integers(40, 500, size=9)]
    items = []
    for i, t in enumerate(lengths):
        vd = _shots_for(t, rng)
        nf = int(vd["n_frames"])
        users = (rng.random((int(rng.integers(3, 21)), nf)) < 0.2).astype(np.uint8)
        key = f"../datasets/eccv16_dataset_{'tvsum' if i % 2 == 0 else 'summe'}_google_pool5.h5/video_{i}"
        items.append((key, orc.synth_features(t, 6000 + i).numpy(), None, vd["cps"], nf, vd["nfps"], vd["picks"], users))
    fs, dv = evaluate(model, items, 0.5, DEV)
    # oracle chain on the device's proposals
    batch = BatchPlan.build(lengths).to(DEV)
    x = torch.from_numpy(np.concatenate([it[1] for it in items])).to(DEV)
    with torch.no_grad():
        cls, loc = model.forward_packed(x, batch)
        nms = model.nms_packed(cls, loc, batch, 0.5)
    counts = nms["keep_count"].cpu().numpy()
    ks, kb = nms["keep_scores"].cpu().numpy(), nms["keep_boxes"].cpu().numpy()
    o, want_f, want_d = 0, [], []
    for v, it in enumerate(items):
        t = lengths[v]
        a, c = o * 4, int(counts[v])
        summ = orc.bbox_summary(t, ks[a:a + c], kb[a:a + c], it[3], it[4], it[5], it[6])
        want_f.append(orc.summ_f1score(summ, it[7], "avg" if "tvsum" in it[0] else "max"))
        want_d.append(orc.summ_diversity(orc.downsample_summ(summ), it[1]))
        o += t
    assert fs == float(sum(want_f) / len(want_f)), (fs, want_f)
    assert abs(dv - sum(want_d) / len(want_d)) < 1e-5
    assert 0.0 < fs < 1.0


# ------------------------------------------------------------------------------------------------ temporal segmentation
KTS = load_npz("kts_golden.npz")


def _kts_features(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_kts_golden_helpers", os.path.join(os.path.dirname(__file__), "golden", "make_kts_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.piecewise_features(int(KTS[f"{name}/T"]), int(KTS[f"{name}/seed"]))


def test_kts_bit_exact_on_reference_kernel_matrices():
    """edsnet_kts on the float32 kernel matrices the reference computed: change points and float64 objective values of
    cpd_auto (and of cpd_nonlin with 3 change points) bit for bit, all golden videos in one packed call."""
    from edsnet_b200 import kts_change_points
    names = [str(n) for n in KTS["cases"]]
    lengths = [int(KTS[f"{n}/T"]) for n in names]
    x = torch.zeros((sum(lengths), 1024), device=DEV)
    kernels = [KTS[f"{n}/K"] for n in names]
    cps, obj = kts_change_points(x, lengths, kernels=kernels)
    for n, c, o in zip(names, cps, obj):
        assert np.array_equal(c, KTS[f"{n}/cps"]), (n, c, KTS[f"{n}/cps"])
        assert np.array_equal(o, KTS[f"{n}/scores"]), n
    cps3, obj3 = kts_change_points(x, lengths, kernels=kernels, m_fixed=3)
    for n, t, c, o in zip(names, lengths, cps3, obj3):
        if t >= 4:
            assert np.array_equal(c, KTS[f"{n}/cps3"]) and np.array_equal(o, KTS[f"{n}/scores3"]), n


def test_kts_from_features_and_shot_tables():
    """Features -> X X^T on the device -> segmentation: same change points as the reference found (its BLAS and the
    device sum X X^T in different orders, the objective agrees to float32 rounding), and the shot tables of
    VideoPreprocessor.kts."""
    from edsnet_b200 import kts_change_points, kts_shots
    names = [str(n) for n in KTS["cases"]]
    feats = [_kts_features(n) for n in names]
    for n, f in zip(names, feats):
        assert orc.rel_l2(np.matmul(f, f.T), KTS[f"{n}/K"]) < 1e-6          # the committed generator reproduces the inputs
    x = torch.from_numpy(np.concatenate(feats)).to(DEV)
    for gram in ("tensor", "fp32"):                                # X X^T on the tcgen05 GEMM / on CUDA cores
        cps, obj = kts_change_points(x, [len(f) for f in feats], gram=gram)
        for n, c, o in zip(names, cps, obj):
            assert np.array_equal(c, KTS[f"{n}/cps"]), (gram, n, c, KTS[f"{n}/cps"])
            # the objective is a difference of float32 prefix sums of K: last-bit differences of K (summation order /
            # split-fp16 products) show up at ~1e-6 of the LARGEST objective value
            assert np.allclose(o, KTS[f"{n}/scores"], rtol=1e-4, atol=5e-6 * float(np.max(KTS[f"{n}/scores"])) + 1e-4)
    n = "t150"
    f = feats[names.index(n)]
    cp, nfps, picks = kts_shots(int(KTS[f"{n}/n_frames"]), torch.from_numpy(f).to(DEV), int(KTS[f"{n}/rate"]))
    assert np.array_equal(cp, KTS[f"{n}/change_points"]) and np.array_equal(nfps, KTS[f"{n}/nfps"])
    assert np.array_equal(picks, np.arange(len(f)) * 15)


def test_gram_on_tensor_cores_matches_float32_matmul():
    """K = X X^T of video_helper.py:117 on the tcgen05 GEMM (three split-fp16 passes): within 2e-6 of the float64 product."""
    from edsnet_b200.kts import gram_tensor_core
    for n in (7, 128, 333):
        f = orc.synth_features(n, 50 + n)
        out = torch.empty((n, n), dtype=torch.float32, device=DEV)
        gram_tensor_core(f.to(DEV), out)
        ref = f.double() @ f.double().t()
        # all-positive features: the tensor core's truncating fp32 accumulation shows as a ~1e-6 low bias
        assert orc.rel_l2(out.cpu().numpy(), ref.numpy()) < 2e-6
    from edsnet_b200 import _capi
    assert _capi.lib().edsnet_debug_tc_status(0) == 0


def test_infer_chain_vs_oracle():
    """infer.py:22-36 from the sampled features on (segmentation -> scores -> NMS -> keyshot summary), packed over
    several videos, against the oracle's host chain fed with the device's kept proposals."""
    from edsnet_b200 import summarize, BatchPlan
    names = ["t60", "t150", "t257"]
    feats = [_kts_features(n) for n in names]
    lengths = [len(f) for f in feats]
    n_frames = [int(KTS[f"{n}/n_frames"]) for n in names]
    scales = [4, 8, 16, 32]
    model = make_model(orc.synth_params(9, "xavier"), scales, 5, "fp16x3", DEV)
    x = torch.from_numpy(np.concatenate(feats)).to(DEV)
    res = summarize(model, x, lengths, n_frames, 0.5)
    batch = BatchPlan.build(lengths).to(DEV)
    with torch.no_grad():
        cls, loc = model.forward_packed(x, batch)
        nms = model.nms_packed(cls, loc, batch, 0.5)
    counts = nms["keep_count"].cpu().numpy()
    ks, kb = nms["keep_scores"].cpu().numpy(), nms["keep_boxes"].cpu().numpy()
    o = 0
    for n, t, nf, r in zip(names, lengths, n_frames, res):
        assert np.array_equal(r["change_points"], KTS[f"{n}/change_points"]) and np.array_equal(r["nfps"], KTS[f"{n}/nfps"])
        a, c = o * len(scales), int(counts[names.index(n)])
        want = orc.bbox_summary(t, ks[a:a + c], kb[a:a + c], r["change_points"], nf, r["nfps"], r["picks"])
        assert r["summary"].shape == (nf,) and np.array_equal(r["summary"], want)
        assert r["summary"].sum() <= int(nf * 0.15)          # (no shot of t60 fits the 15 % budget: an empty summary)
        o += t
    assert sum(int(r["summary"].sum()) for r in res) > 0
