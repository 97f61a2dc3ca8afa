"""CPU checks of the index arithmetic behind the wide read-outs (gemm_tc.cuh wide2_slot and the recursive
halving of the row sums; attn_tc.cuh uses the same slot function for the value convolution and attn_out).  The kernels
are only trusted on the GPU against the oracle; these tests pin the PROPERTIES the design argues with: every slot of a
row is used exactly once, neither access pattern has a shared-memory bank conflict, and the reduction leaves every
(row, iteration) sum in exactly one known lane."""
import itertools

import numpy as np


def wide2_slot(s: int, r: int) -> int:
    """16-byte slot s (0..15: four fp32 columns each) of region row r -> physical slot (gemm_tc.cuh)."""
    return (s & 8) | (((s & 7) ^ (r & 7) ^ (s >> 3)) & 7)


def bank_group(row: int, phys_slot: int) -> int:
    """Which of the eight 16-byte bank groups of a 128-byte shared-memory line an access touches (region rows are 256 B:
    a multiple of 128, so only the slot decides)."""
    return (row * 256 + phys_slot * 16) // 16 % 8


def test_slot_map_is_a_permutation_per_row():
    for r in range(32):
        assert sorted(wide2_slot(s, r) for s in range(16)) == list(range(16))
        # the upper half of a row (slots 8..15 = the second warp's columns, or the lo-plane piece in attn_out) stays there
        assert all((wide2_slot(s, r) & 8) == (s & 8) for s in range(16))


def test_phase_a_thread_per_row_is_conflict_free():
    # a 128-bit warp access is served one quarter-warp (8 consecutive lanes = 8 consecutive rows) at a time
    for s, q in itertools.product(range(16), range(4)):
        groups = {bank_group(r, wide2_slot(s, r)) for r in range(8 * q, 8 * q + 8)}
        assert len(groups) == 8, (s, q)


def test_phase_b_eight_lanes_per_row_is_conflict_free():
    # lanes c = 0..7 of a quarter-warp read slots 2c (first load) and 2c + 1 (second load) of the SAME row
    for r, odd in itertools.product(range(32), range(2)):
        groups = {bank_group(r, wide2_slot(2 * c + odd, r)) for c in range(8)}
        assert len(groups) == 8, (r, odd)


def test_attn_out_pieces_keep_the_property():
    # attn_out borrows two 4 KB pieces of the dead q stage (rows of 128 B in the hi plane and in the lo plane, 16 KB
    # apart): slot >> 3 selects the piece, the low three bits the 16-byte chunk of the 128-byte row
    def addr(quarter, r, slot):
        return (slot >> 3) * 16384 + (quarter * 32 + r) * 128 + (((slot & 7) ^ (r & 7) ^ (slot >> 3)) & 7) * 16
    for quarter in range(4):
        seen = {addr(quarter, r, s) for r in range(32) for s in range(16)}
        assert len(seen) == 512                                            # 32 rows x 16 slots, no overlap
        lo, hi = quarter * 32 * 128, (quarter + 1) * 32 * 128
        assert all(lo <= a % 16384 < hi for a in seen)                      # only this pair's own P rows
        for s, q in itertools.product(range(16), range(4)):
            assert len({addr(quarter, r, s) // 16 % 8 for r in range(8 * q, 8 * q + 8)}) == 8
        for r, odd in itertools.product(range(32), range(2)):
            assert len({addr(quarter, r, 2 * c + odd) // 16 % 8 for c in range(8)}) == 8


def test_recursive_halving_of_the_row_sums():
    """gemm_tc.cuh, WIDE == 2: lane c of a row holds s[0..3] (iterations 0..3); after xor-1, xor-2, xor-4 exchanges
    lane c holds the sum over the eight lanes for iteration (c & 1) * 2 + ((c >> 1) & 1)."""
    rng = np.random.default_rng(0)
    s = rng.integers(-1000, 1000, size=(8, 4)).astype(np.int64)           # [lane][iteration], exact arithmetic
    b0 = np.arange(8) & 1
    b1 = (np.arange(8) >> 1) & 1
    keep = np.where(b0[:, None] == 1, s[:, 2:4], s[:, 0:2]).copy()         # [lane][2]
    send = np.where(b0[:, None] == 1, s[:, 0:2], s[:, 2:4])
    keep += send[np.arange(8) ^ 1]
    u = np.where(b1 == 1, keep[:, 1], keep[:, 0]).copy()
    snd = np.where(b1 == 1, keep[:, 0], keep[:, 1])
    u += snd[np.arange(8) ^ 2]
    u += u[np.arange(8) ^ 4]
    it = b0 * 2 + b1
    assert np.array_equal(u, s.sum(0)[it])
    assert sorted(it[:4]) == [0, 1, 2, 3]                                   # lanes 0..3 cover every iteration once
