"""CPU checks of the two numerical claims the tcgen05 pseudo-inverse chain (csrc/pinv_tc.cuh) is built on
(transformer/nystroformer.py:13-28 is the iteration it implements):
  1. carrying XZ by the recurrence XZ' = XZ T2 / 4 instead of recomputing A Z every iteration leaves W = Z a3v as close to
     the float64 result as the reference's own float32 iteration;
  2. every iterate XZ_k is symmetric with spectrum in [0, 1], so |XZ| <= 1, |7I - XZ| <= 7, |15I - XZ T1| <= 15,
     |13I - XZ U| <= 13 entry-wise: the compile-time plane scales of the kernel cannot overflow."""
import numpy as np
import pytest
import torch

from oracle import dsnet_oracle as orc


def _attn2_cases():
    for T, seed, init, scale in [(320, 1, "xavier", 1.0), (450, 2, "default", 1.0), (100, 3, "xavier", 1.0),
                                 (800, 4, "xavier", 1e3), (257, 5, "default", 1e-4)]:
        x = orc.synth_features(T, seed) * scale
        p = orc.synth_params(seed + 10, init)
        st = {}
        with torch.no_grad():
            orc.nystrom_attention(x, p, stages=st)
        yield st["attn2"], st["a3v"]


def _iterate(a, recurrence, iters=6, track=None):
    mag = a.abs()
    z = a.transpose(-1, -2) / (mag.sum(-1).max() * mag.sum(-2).max())
    eye = torch.eye(a.shape[-1], dtype=a.dtype)
    xz = a @ z
    for _ in range(iters):
        if not recurrence:
            xz = a @ z
        t1 = 7 * eye - xz
        u = 15 * eye - xz @ t1
        t2 = 13 * eye - xz @ u
        if track is not None:
            track.append((xz, t1, u, t2))
        z = 0.25 * z @ t2
        xz = 0.25 * xz @ t2
    return z


def test_recurrence_keeps_w_at_float32_grade():
    worst_std, worst_rec = 0.0, 0.0
    for a2, a3v in _attn2_cases():
        w64 = _iterate(a2.double(), False) @ a3v.double()
        err = lambda z: float((z.double() @ a3v.double() - w64).norm() / w64.norm())
        worst_std = max(worst_std, err(_iterate(a2, False)))
        worst_rec = max(worst_rec, err(_iterate(a2, True)))
    assert worst_rec < 3e-6, worst_rec
    assert worst_rec < 4 * worst_std + 1e-6, (worst_rec, worst_std)


def test_recurrence_equals_the_reference_iteration_in_float64():
    for a2, _ in _attn2_cases():
        assert torch.allclose(_iterate(a2.double(), True), orc.pinv_iterative(a2.double()), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_entrywise_bounds_behind_the_fixed_plane_scales(dtype):
    for a2, _ in _attn2_cases():
        track = []
        _iterate(a2.to(dtype), True, track=track)
        for xz, t1, u, t2 in track:
            # 2x headroom in the kernel: scales 2^14 (XZ), 2^12 (T1), 2^11 (U, T2) against fp16's 65504
            assert float(xz.abs().max()) <= 1.0 + 1e-4
            assert float(t1.abs().max()) <= 7.0 + 1e-3
            assert float(u.abs().max()) <= 15.0 + 1e-3
            assert float(t2.abs().max()) <= 13.0 + 1e-3
            assert float((xz - xz.transpose(-1, -2)).abs().max()) < 1e-4          # symmetric up to rounding
