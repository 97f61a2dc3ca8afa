"""Golden vectors for the training-side host helpers, generated from the REAL reference (dev container only):
seq2bbox + lr2cw, get_pos_label at the three thresholds of anchor_based/train.py, bbox2offset.

    python tests/golden/make_labels_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import import_reference  # noqa: E402


def main():
    _, anchor_helper, bbox_helper, _ = import_reference()
    rng = np.random.default_rng(11)
    out = {}
    cases = []
    for name, T, scales, p_on in [("t40", 40, [4, 8], 0.6), ("t320", 320, [4, 8, 16, 32], 0.15), ("t97", 97, [12], 0.2),
                                  ("t64edge", 64, [4, 32], 0.5)]:
        # blocky random mask (runs of ones), like a down-sampled keyshot summary
        mask = np.zeros(T, dtype=bool)
        t = 0
        while t < T:
            run = int(rng.integers(1, 12))
            if rng.random() < p_on:
                mask[t:t + run] = True
            t += run
        if name == "t64edge":
            mask[0] = mask[-1] = True
        lr = bbox_helper.seq2bbox(mask)
        cw = bbox_helper.lr2cw(lr)
        anchors = anchor_helper.get_anchors(T, scales)
        rec = {"mask": mask, "scales": np.asarray(scales), "lr": lr, "cw": cw}
        for tag, th in (("pos", 0.6), ("neg", 0.0), ("inc", 0.3)):
            c, l = anchor_helper.get_pos_label(anchors, cw, th)
            rec[f"cls_{tag}"] = c
            rec[f"loc_{tag}"] = l
        for k, v in rec.items():
            out[f"{name}/{k}"] = np.asarray(v)
        cases.append(name)
        print(name, "segments", len(lr), "positives", int(rec["cls_pos"].sum()))
    out["cases"] = np.asarray(cases)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/labels_golden.npz"), **out)


if __name__ == "__main__":
    main()
