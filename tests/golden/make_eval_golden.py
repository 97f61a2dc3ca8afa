"""Golden vectors for the evaluation metrics (F-score, diversity) from the REAL reference helpers.

    python tests/golden/make_eval_golden.py        (dev container only: needs /root/reference)

Inputs are seeded; the outputs are what helpers/vsumm_helper.py:get_summ_f1score / downsample_summ /
get_summ_diversity return.  Also stores the literal of the reference's own unit test
(tests/helpers/test_vsumm_helper.py:36-40)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import dsnet_oracle as orc  # noqa: E402
import make_golden  # noqa: E402


def shot_summary(rng, n_frames, frac):
    """random keyshot-like 0/1 vector: shots of 30..120 frames, each selected with probability frac"""
    out = np.zeros(n_frames, dtype=np.uint8)
    f = 0
    while f < n_frames:
        ln = int(rng.integers(30, 121))
        if rng.random() < frac:
            out[f:f + ln] = 1
        f += ln
    return out


def main():
    _, _, _, vsumm_helper = make_golden.import_reference()
    rng = np.random.default_rng(777)
    g = {}
    cases = []
    # (name, T, users, delta of the users' frame count vs the prediction's, metric)
    for name, T, U, delta, metric in [("tv_avg", 320, 20, 0, "avg"), ("sm_max", 450, 15, 0, "max"),
                                      ("short_users", 200, 7, -9, "avg"), ("long_users", 130, 9, 11, "max"),
                                      ("u3", 100, 3, 0, "avg"), ("empty_pred", 64, 5, 0, "avg"),
                                      ("one_pos", 90, 4, 0, "max")]:
        n_frames = 15 * T - int(rng.integers(0, 15))          # ceil(n_frames / 15) == T
        pred = shot_summary(rng, n_frames, 0.15)
        if name == "empty_pred":
            pred[:] = 0
        if name == "one_pos":
            pred[:] = 0
            pred[15 * 7] = 1
        users = np.stack([shot_summary(rng, n_frames + delta, 0.15) for _ in range(U)])
        x = orc.synth_features(T, 4000 + len(cases)).numpy()
        f = vsumm_helper.get_summ_f1score(pred.astype(bool), users, metric)
        ds = vsumm_helper.downsample_summ(pred.astype(bool))
        d = vsumm_helper.get_summ_diversity(ds, x)
        per_user = [vsumm_helper.f1_score(u.astype(bool),
                                          np.pad(pred, (0, max(0, users.shape[1] - n_frames)))[:users.shape[1]].astype(bool))
                    for u in users]
        g.update({f"{name}/T": T, f"{name}/n_frames": n_frames, f"{name}/pred": np.packbits(pred),
                  f"{name}/users": np.packbits(users, axis=1), f"{name}/users_frames": users.shape[1],
                  f"{name}/metric": np.asarray(metric), f"{name}/x_seed": 4000 + len(cases),
                  f"{name}/fscore": np.float64(f), f"{name}/diversity": np.float64(d),
                  f"{name}/user_f1": np.asarray(per_user, dtype=np.float64)})
        cases.append(name)
        print(name, "fscore", f, "diversity", d, "selected", int(ds.sum()))
    # the reference's own known answer (tests/helpers/test_vsumm_helper.py:36-40)
    kat = vsumm_helper.f1_score(np.array([0, 1, 1, 0, 1], dtype=bool), np.array([1, 1, 0, 1, 1], dtype=bool))
    assert abs(kat - 4 / 7) < 1e-15
    g["kat_f1"] = np.float64(kat)
    g["cases"] = np.asarray(cases)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/eval_golden.npz"), **g)


if __name__ == "__main__":
    main()
