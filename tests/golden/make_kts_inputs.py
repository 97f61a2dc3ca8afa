"""Seeded piecewise-stationary feature sequences for the temporal-segmentation goldens (no reference needed: used by
tests/golden/make_kts_golden.py to generate the vectors and by the GPU tests to rebuild the same inputs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import dsnet_oracle as orc  # noqa: E402


def piecewise_features(T, seed):
    rng = np.random.default_rng(seed)
    x = orc.synth_features(T, seed).numpy()
    n_cut = max(1, T // 25)
    cuts = np.sort(rng.choice(np.arange(2, max(3, T - 2)), size=min(n_cut, max(1, T - 4)), replace=False)) if T > 6 else np.array([], int)
    base = rng.standard_normal((len(cuts) + 1, 1024)).astype(np.float32)
    lab = np.searchsorted(cuts, np.arange(T), side="right")
    f = (0.6 * base[lab] + 0.4 * x).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    return f
