"""Golden vectors for the kernel temporal segmentation from the REAL reference (kts/cpd_auto.py, kts/cpd_nonlin.py,
helpers/video_helper.py:109-126 restated inline because video_helper imports cv2 / torchvision models at load).

    python tests/golden/make_kts_golden.py        (dev container only: needs /root/reference)

Inputs: seeded piecewise-stationary feature sequences (so that there are change points to find); stored: the float32
kernel matrix the reference computed (np.matmul), the change points, the objective values and the derived shot tables."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, "/root/reference/src")
from oracle import dsnet_oracle as orc  # noqa: E402
from kts.cpd_auto import cpd_auto  # noqa: E402
from kts.cpd_nonlin import cpd_nonlin  # noqa: E402


from make_kts_inputs import piecewise_features  # noqa: E402


def main():
    g = {}
    cases = []
    for name, T, seed, rate, tail in [("t8", 8, 1, 15, 3), ("t60", 60, 2, 15, 0), ("t150", 150, 3, 15, 7),
                                      ("t257", 257, 4, 15, 14), ("t400", 400, 5, 15, 1)]:
        f = piecewise_features(T, seed)
        n_frames = T * rate - tail
        K = np.matmul(f, f.T)
        cps, scores = cpd_auto(K, T - 1, 1, verbose=False)
        # helpers/video_helper.py:118-126
        cp = cps * rate
        cp = np.hstack((0, cp, n_frames))
        begin, end = cp[:-1], cp[1:]
        change_points = np.vstack((begin, end - 1)).T
        nfps = end - begin
        # fixed number of change points as well (cpd_nonlin on its own)
        cps3, sc3 = cpd_nonlin(K, min(3, T - 1), verbose=False)
        g.update({f"{name}/T": T, f"{name}/seed": seed, f"{name}/rate": rate, f"{name}/n_frames": n_frames,
                  f"{name}/K": K, f"{name}/cps": cps.astype(np.int64), f"{name}/scores": scores,
                  f"{name}/change_points": change_points.astype(np.int64), f"{name}/nfps": nfps.astype(np.int64),
                  f"{name}/cps3": cps3.astype(np.int64), f"{name}/scores3": sc3})
        cases.append(name)
        print(name, "change points:", len(cps), cps[:10])
    g["cases"] = np.asarray(cases)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/kts_golden.npz"), **g)


if __name__ == "__main__":
    main()
