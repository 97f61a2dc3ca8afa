"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (pure Python/PyTorch) is imported from /root/reference/src with
empty stand-ins for third-party modules that are absent in this image and never
touched on the hot path.  Inputs and weights come from
the seeded generators in oracle/dsnet_oracle.py; each record stores a SHA-256
of the generated input and weights so the tests notice RNG drift, plus the
reference's outputs.  Nothing here runs on the GPU box.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dsnet_oracle as orc  # noqa: E402

REF = "/root/reference/src"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    dummy = type("Dummy", (), {})
    _stub("local_attention", LocalAttention=dummy)
    _stub("axial_positional_embedding", AxialPositionalEmbedding=dummy)
    _stub("performer_pytorch")
    _stub("performer_pytorch.reversible", ReversibleSequence=dummy, SequentialSequence=dummy)
    _stub("pywt")
    _stub("h5py", File=dummy)
    _stub("matplotlib")
    _stub("matplotlib.pyplot", plot=lambda *a, **k: None, show=lambda *a, **k: None)
    _stub("ortools")
    _stub("ortools.algorithms")
    _stub("ortools.algorithms.python")
    _stub("ortools.algorithms.python.knapsack_solver")
    sys.modules["ortools.algorithms.python"].knapsack_solver = sys.modules["ortools.algorithms.python.knapsack_solver"]
    sys.path.insert(0, REF)
    from anchor_based.dsnet import DSNet
    from anchor_based import anchor_helper
    from helpers import bbox_helper, vsumm_helper
    vsumm_helper.knapsack = orc.knapsack_dp      # ortools absent: same exact solver on both sides
    return DSNet, anchor_helper, bbox_helper, vsumm_helper


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def params_sha(p):
    return sha(*[p[k].numpy() for k in sorted(p)])


def ref_state_dict(p, fc_depth):
    sd = dict(p)
    for i in range(fc_depth):           # aliased shared block, anchor_based/dsnet.py:96
        for j in ("0.weight", "0.bias", "3.weight", "3.bias"):
            sd[f"fc.{i}.{j}"] = p[f"fc_block.{j}"]
    return sd


FORWARD_CASES = [
    # name, T, scales, fc_depth, init, x_seed, w_seed
    ("c1_T320_s12", 320, [12], 5, "default", 12345, 12345),
    ("c1x_T320_s12_xavier", 320, [12], 5, "xavier", 12345, 777),
    ("T100_s12", 100, [12], 5, "xavier", 1, 2),
    ("T450_s4_8_16_32_d7", 450, [4, 8, 16, 32], 7, "xavier", 3, 4),
    ("T64_s4_8", 64, [4, 8], 5, "default", 5, 6),
    ("T37_s4_32", 37, [4, 32], 3, "xavier", 7, 8),
    ("T800_s12", 800, [12], 5, "default", 9, 10),
    ("T1_s4", 1, [4], 2, "xavier", 11, 12),
    ("T2048_s4_8_16_32", 2048, [4, 8, 16, 32], 5, "xavier", 13, 14),
]


def main():
    torch.set_num_threads(1)          # single-thread => run-to-run deterministic reductions
    DSNet, anchor_helper, bbox_helper, vsumm_helper = import_reference()
    out = {}
    for name, T, scales, depth, init, xs, ws in FORWARD_CASES:
        x = orc.synth_features(T, xs)
        p = orc.synth_params(ws, init)
        model = DSNet("nystromformer", 1024, 128, list(scales), 8, fc_depth=depth,
                      orientation=None, pooling_type="roi").eval()
        model.load_state_dict(ref_state_dict(p, depth), strict=True)
        with torch.no_grad():
            cls, loc = model(x[None])
            scores, boxes = model.predict(x[None])
        ib = np.clip(boxes, 0, T).round().astype(np.int32)          # evaluate.py:26
        ks, kb = bbox_helper.nms(scores, ib, 0.5)
        # fp64 run of the same reference: error floor of the fp32 reference itself
        m64 = DSNet("nystromformer", 1024, 128, list(scales), 8, fc_depth=depth,
                    orientation=None, pooling_type="roi").eval()
        m64.load_state_dict(ref_state_dict(p, depth), strict=True)
        m64 = m64.double()
        with torch.no_grad():
            cls64, loc64 = m64(x[None].double())
        rec = dict(T=T, scales=np.asarray(scales), fc_depth=depth, x_seed=xs, w_seed=ws,
                   x_sha=sha(x.numpy()), w_sha=params_sha(p),
                   pred_cls=cls.numpy(), pred_loc=loc.numpy(),
                   pred_cls64=cls64.numpy(), pred_loc64=loc64.numpy(),
                   boxes_f32=boxes, boxes_i32=ib, keep_scores=ks, keep_boxes=kb)
        for k, v in rec.items():
            out[f"{name}/{k}"] = np.asarray(v)
        out[f"{name}/init"] = np.asarray(init)
        print(name, "ref fp32 vs fp64 rel-l2 cls/loc:",
              orc.rel_l2(cls.numpy(), cls64.numpy()), orc.rel_l2(loc.numpy(), loc64.numpy()),
              "kept", len(ks))
    out["forward_cases"] = np.asarray([c[0] for c in FORWARD_CASES])
    np.savez_compressed(os.path.join(ROOT, "tests/golden/forward_golden.npz"), **out)

    # ---- full multi-head attention base (config 4), reference DSNet(base_model='attention') ----
    mha = {}
    mha_cases = [("mha_T300_s4_8", 300, [4, 8], 5, "xavier", 21, 22), ("mha_T77_s12", 77, [12], 3, "default", 23, 24),
                 ("mha_T2048_s4_8_16_32", 2048, [4, 8, 16, 32], 5, "xavier", 25, 26)]
    for name, T, scales, depth, init, xs, ws in mha_cases:
        x = orc.synth_features(T, xs)
        p = orc.synth_params_mha(ws, init)
        model = DSNet("attention", 1024, 128, list(scales), 8, fc_depth=depth, orientation=None,
                      pooling_type="roi").eval()
        model.load_state_dict(ref_state_dict(p, depth), strict=True)
        with torch.no_grad():
            cls, loc = model(x[None])
        rec = dict(T=T, scales=np.asarray(scales), fc_depth=depth, x_seed=xs, w_seed=ws, x_sha=sha(x.numpy()),
                   w_sha=params_sha(p), pred_cls=cls.numpy(), pred_loc=loc.numpy())
        for k, v in rec.items():
            mha[f"{name}/{k}"] = np.asarray(v)
        mha[f"{name}/init"] = np.asarray(init)
        print(name, "done")
    mha["forward_cases"] = np.asarray([c[0] for c in mha_cases])
    np.savez_compressed(os.path.join(ROOT, "tests/golden/forward_mha_golden.npz"), **mha)

    # ---- stand-alone decode / NMS / summary vectors from the reference's host helpers ----
    rng = np.random.default_rng(2024)
    dn = {}
    for name, T, scales in [("n320", 320, [12]), ("n800x4", 800, [4, 8, 16, 32]), ("n3000x4", 3000, [4, 8, 16, 32])]:
        N = T * len(scales)
        loc = np.stack([rng.normal(0, 0.6, N), rng.normal(0, 0.5, N)], 1).astype(np.float32)
        scores = rng.permutation(N).astype(np.float32) / np.float32(N)            # tie-free
        anchors = anchor_helper.get_anchors(T, scales).reshape(-1, 2)
        cw = anchor_helper.offset2bbox(loc, anchors)
        lr = bbox_helper.cw2lr(cw)
        ib = np.clip(lr, 0, T).round().astype(np.int32)
        ks, kb = bbox_helper.nms(scores, ib, 0.5)
        ks3, kb3 = bbox_helper.nms(scores, ib, 0.3)
        # summary through the reference's own bbox2summary (knapsack = shared exact DP)
        shot = 5
        nseg = T // shot
        nfps = np.full(nseg, shot * 15, dtype=np.int32)
        nfps[-1] += (T - nseg * shot) * 15
        ends = np.cumsum(nfps)
        cps = np.stack([np.concatenate([[0], ends[:-1]]), ends - 1], 1).astype(np.int32)
        picks = (np.arange(T) * 15).astype(np.int32)
        summ = vsumm_helper.bbox2summary(T, ks, kb, cps, T * 15, nfps, picks)
        dn.update({f"{name}/T": T, f"{name}/scales": np.asarray(scales), f"{name}/loc": loc,
                   f"{name}/scores": scores, f"{name}/anchors": anchors, f"{name}/cw": cw, f"{name}/lr": lr,
                   f"{name}/boxes_i32": ib, f"{name}/keep_scores": ks, f"{name}/keep_boxes": kb,
                   f"{name}/keep_scores_t03": ks3, f"{name}/keep_boxes_t03": kb3,
                   f"{name}/cps": cps, f"{name}/nfps": nfps, f"{name}/picks": picks,
                   f"{name}/summary": np.packbits(summ)})
        print(name, "kept", len(ks), len(ks3), "summary frames", int(summ.sum()))
    dn["cases"] = np.asarray(["n320", "n800x4", "n3000x4"])
    np.savez_compressed(os.path.join(ROOT, "tests/golden/decode_nms_golden.npz"), **dn)


if __name__ == "__main__":
    main()
