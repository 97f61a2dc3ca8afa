"""Host-side batch planning for packed variable-length videos.

The reference scores one video per call (`for ... in val_loader`, evaluate.py:19; batch 1 is hard-wired by the
`.view(seq_len, num_scales)` at anchor_based/dsnet.py:114-115).  Here many videos are packed row-wise into one
[total_rows, 1024] matrix; the tables below tell the kernels where each video lives.  Pure NumPy, no device work:
`BatchPlan.to(device)` does one small H2D copy of all tables.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import _capi


def _tiles(lengths: np.ndarray, tile: int) -> np.ndarray:
    """[(video, first_row)] for every `tile`-row tile of every video, int32 [n_tiles, 2]."""
    counts = (lengths + tile - 1) // tile
    vid = np.repeat(np.arange(len(lengths), dtype=np.int64), counts)
    starts = np.cumsum(counts) - counts
    first = (np.arange(int(counts.sum()), dtype=np.int64) - np.repeat(starts, counts)) * tile
    return np.stack([vid, first], axis=1).astype(np.int32)


@dataclass
class BatchPlan:
    lengths: np.ndarray          # int64 [V]
    cu_rows: np.ndarray          # int32 [V+1]
    tiles64: np.ndarray          # int32 [n64, 2]
    tiles128: np.ndarray         # int32 [n128, 2]

    @staticmethod
    def build(lengths: Sequence[int]) -> "BatchPlan":
        ln = np.asarray(list(lengths), dtype=np.int64)
        if ln.ndim != 1 or ln.size == 0:
            raise ValueError("need at least one video")
        if (ln < 1).any():
            raise ValueError("every video needs at least one frame")
        if ln.size > 65535:
            raise ValueError("at most 65535 videos per packed batch")
        total = int(ln.sum())
        if total >= 2 ** 31 // 1536:
            raise ValueError("packed batch too large for 32-bit row indexing; split it")
        cu = np.zeros(ln.size + 1, dtype=np.int32)
        cu[1:] = np.cumsum(ln)
        return BatchPlan(ln, cu, _tiles(ln, 64), _tiles(ln, 128))

    @property
    def n_videos(self) -> int:
        return int(self.lengths.size)

    @property
    def total_rows(self) -> int:
        return int(self.cu_rows[-1])

    @property
    def max_rows(self) -> int:
        return int(self.lengths.max())

    def packed_tables(self) -> np.ndarray:
        """cu_rows (padded to an even count: the tile tables are read as 8-byte int2) | tiles64 | tiles128 in one
        int32 array (one H2D copy)."""
        pad = np.zeros(self.cu_rows.size % 2, dtype=np.int32)
        return np.concatenate([self.cu_rows, pad, self.tiles64.reshape(-1), self.tiles128.reshape(-1)])

    def nms_scratch(self, n_scales: int):
        """Byte offsets (int64 [V]) and total bytes of the global NMS scratch, needed only by videos with more than
        4096 anchors (edsnet_decode_nms contract: 48 bytes per anchor rounded up to a power of two)."""
        n = self.lengths * int(n_scales)
        p = np.where(n > 4096, 2 ** np.ceil(np.log2(np.maximum(n, 1))).astype(np.int64), 0)
        sizes = p * 48
        off = np.cumsum(sizes) - sizes
        return off.astype(np.int64), int(sizes.sum())

    def to(self, device) -> "DeviceBatch":
        return DeviceBatch(self, device)


class DeviceBatch:
    """BatchPlan with its tables resident on `device`; owns the ctypes edsnet_batch struct."""

    def __init__(self, plan: BatchPlan, device):
        import torch
        self.plan = plan
        host = torch.from_numpy(plan.packed_tables())
        if torch.device(device).type == "cuda":
            host = host.pin_memory()
        self._host = host                      # keeps the pinned staging buffer alive until the async copy ran
        self.tables = host.to(device, non_blocking=True)
        n_cu = plan.cu_rows.size + plan.cu_rows.size % 2
        n64 = plan.tiles64.shape[0]
        self.cu_rows = self.tables[:plan.cu_rows.size]
        self.tiles64 = self.tables[n_cu:n_cu + 2 * n64]
        self.tiles128 = self.tables[n_cu + 2 * n64:]
        b = _capi.Batch()
        b.n_videos = plan.n_videos
        b.total_rows = plan.total_rows
        b.max_rows = plan.max_rows
        b.cu_rows = self.cu_rows.data_ptr()
        b.tiles64 = self.tiles64.data_ptr()
        b.n_tiles64 = n64
        b.tiles128 = self.tiles128.data_ptr()
        b.n_tiles128 = plan.tiles128.shape[0]
        self._cu_host = np.ascontiguousarray(plan.cu_rows, dtype=np.int32)      # kept alive with the struct
        b.cu_rows_host = self._cu_host.ctypes.data
        self.struct = b


def shard_videos(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Video-wise partition for multi-GPU scoring (no collective on the data path): greedy longest-first
    assignment to the currently lightest rank, balancing the padded row count that drives the cost.
    Returns, per rank, the ascending list of video indices it owns.  Deterministic."""
    ln = np.asarray(list(lengths), dtype=np.int64)
    cost = ((ln + 63) // 64) * 64
    order = np.argsort(-cost, kind="stable")
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(int(i))
        loads[r] += int(cost[i])
    return [sorted(o) for o in out]
