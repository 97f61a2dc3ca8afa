"""Gradient goldens from the REAL reference (dev container only; needs /root/reference):

    python tests/golden/make_grad_golden.py

For each case the reference's DSNet (anchor_based/dsnet.py) is built in eval() mode (Dropout off: train-mode masks cannot
be matched across RNGs), scores one seeded video, the reference's OWN losses (anchor_based/losses.py
calc_cls_loss + calc_loc_loss, combined as anchor_based/train.py:119-123) are taken on seeded labels, and loss.backward()
gives the gradient of all 16 parameter tensors.  A 2.25 M-value gradient per case is too large to commit, so the record
keeps every small tensor in full and, for the four weight matrices, the Frobenius norm, a seeded random projection and a
strided sample of 4096 entries -- enough to pin the oracle's autograd (tests/test_oracle_golden.py), which the GPU
tests then compare against in full.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, ref_state_dict, sha, params_sha  # noqa: E402
from oracle import dsnet_oracle as orc  # noqa: E402

CASES = [  # name, T, scales, fc_depth, init, x_seed, w_seed, label_seed
    ("g_T320_s12", 320, [12], 5, "xavier", 31, 32, 33),
    ("g_T150_s4_8_16_32", 150, [4, 8, 16, 32], 5, "xavier", 41, 42, 43),
    ("g_T77_s4_8_default", 77, [4, 8], 3, "default", 51, 52, 53),
]
BIG = ("base_model.to_qkv.weight", "base_model.to_out.0.weight", "fc1.weight", "fc_block.0.weight")


def synth_labels(T, S, seed):
    """Seeded labels shaped like anchor_based/train.py:86-108 produces them: ~10 % positives, twice as many negatives."""
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(T, S, generator=g)
    cls = torch.zeros(T, S)
    cls[r < 0.1] = 1
    cls[(r >= 0.1) & (r < 0.3)] = -1
    loc = torch.randn(T, S, 2, generator=g) * 0.8
    return cls, loc


def digest(name, g):
    g = g.detach().double().numpy()
    if name not in BIG:
        return {"full": g.astype(np.float64)}
    flat = g.reshape(-1)
    rng = np.random.default_rng(12345)
    proj = rng.standard_normal(flat.size)
    idx = np.linspace(0, flat.size - 1, 4096).astype(np.int64)
    return {"norm": np.float64(np.linalg.norm(flat)), "proj": np.float64(flat @ proj), "sample": flat[idx]}


def main():
    DSNet, _, _, _ = import_reference()
    from anchor_based.losses import calc_cls_loss, calc_loc_loss
    out = {"grad_cases": np.array([c[0] for c in CASES])}
    for name, T, scales, depth, init, xs, ws, ls in CASES:
        x = orc.synth_features(T, xs)
        p = orc.synth_params(ws, init)
        model = DSNet("nystromformer", 1024, 128, scales, 8, fc_depth=depth, orientation=None, pooling_type="roi").eval()
        model.load_state_dict(ref_state_dict(p, depth), strict=True)
        cls_label, loc_label = synth_labels(T, len(scales), ls)
        pred_cls, pred_loc = model(x[None])
        loc_loss = calc_loc_loss(pred_loc, loc_label, cls_label)
        cls_loss = calc_cls_loss(pred_cls, cls_label)
        loss = cls_loss + 1.0 * loc_loss
        model.zero_grad()
        loss.backward()
        named = dict(model.named_parameters())
        pre = name + "/"
        out.update({pre + "T": T, pre + "scales": np.array(scales), pre + "fc_depth": depth, pre + "init": init,
                    pre + "x_seed": xs, pre + "w_seed": ws, pre + "label_seed": ls, pre + "x_sha": sha(x.numpy()),
                    pre + "w_sha": params_sha(p), pre + "loss": np.float64(loss.item()),
                    pre + "cls_loss": np.float64(cls_loss.item()), pre + "loc_loss": np.float64(loc_loss.item()),
                    pre + "cls_label": cls_label.numpy().astype(np.int8), pre + "loc_label": loc_label.numpy()})
        for k in p:
            for kk, v in digest(k, named[k].grad).items():
                out[f"{pre}grad/{k}/{kk}"] = v
        print(name, "loss", loss.item())
    np.savez_compressed(os.path.join(HERE, "grad_golden.npz"), **out)


if __name__ == "__main__":
    main()
