"""Timing of the device temporal segmentation against the CPU restatement of the reference (oracle) on synthetic
TVSum/SumMe-shape videos.  Usage: python tools/kts_bench.py [n_videos]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from edsnet_b200 import kts_change_points  # noqa: E402
from oracle import dsnet_oracle as orc  # noqa: E402
from make_kts_inputs import piecewise_features  # noqa: E402


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rng = np.random.default_rng(5)
    lengths = [int(t) for t in rng.integers(100, 801, size=nv)]
    feats = [piecewise_features(t, 100 + i) for i, t in enumerate(lengths)]
    x = torch.from_numpy(np.concatenate(feats)).cuda()
    kts_change_points(x[:lengths[0]], lengths[:1])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cps, _ = kts_change_points(x, lengths)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"device: {nv} videos ({sum(lengths)} frames) in {dt * 1e3:.1f} ms = {nv / dt:.1f} videos/s")
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 20 and k < nv:
        f = feats[k]
        c, _ = orc.kts_auto(np.matmul(f, f.T), len(f) - 1, 1)
        assert np.array_equal(c, cps[k]), k
        k += 1
    dt = time.perf_counter() - t0
    print(f"cpu oracle: {k} videos in {dt:.1f} s = {k / dt:.2f} videos/s (change points identical)")


if __name__ == "__main__":
    main()
