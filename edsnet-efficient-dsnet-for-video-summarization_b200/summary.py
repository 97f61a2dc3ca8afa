"""Kept proposals -> keyshot summaries on the device (the step after NMS in evaluate.py:29 / infer.py:35:
`vsumm_helper.bbox2summary`, src/helpers/vsumm_helper.py:101-116 -> get_keyshot_summ :53-98 -> knapsack :26-45).

`ShotPlan` packs the per-video shot structure (change points, frames per shot, picks, n_frames) the reference reads
from its h5 files; `keyshot_summaries` runs `edsnet_keyshot_summary` on the output of `DSNet.nms_packed`.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch

from . import _capi
from .plan import DeviceBatch


class ShotPlan:
    def __init__(self, videos: Sequence[dict], device, proportion: float = 0.15):
        """videos: one dict per video with `cps` (n_seg, 2) int, `nfps` (n_seg,) int, `picks` (T,) int, `n_frames` int
        -- the arguments of vsumm_helper.bbox2summary."""
        cu_seg, cu_frames = [0], [0]
        cps, nfps, picks, cap, gcds, dp_off = [], [], [], [], [], []
        off = 0
        for vd in videos:
            c = np.asarray(vd["cps"], dtype=np.int32).reshape(-1, 2)
            w = np.asarray(vd["nfps"], dtype=np.int32).reshape(-1)
            if len(c) != len(w):
                raise ValueError("cps and nfps disagree")
            n_frames = int(vd["n_frames"])
            capacity = int(n_frames * proportion)                      # vsumm_helper.py:89
            pos = [int(x) for x in w if x > 0]
            g = math.gcd(capacity, *pos) if pos and capacity > 0 else 1
            g = max(g, 1)
            cps.append(c)
            nfps.append(w)
            picks.append(np.asarray(vd["picks"], dtype=np.int32).reshape(-1))
            cap.append(capacity // g)
            gcds.append(g)
            cu_seg.append(cu_seg[-1] + len(w))
            cu_frames.append(cu_frames[-1] + n_frames)
            dp_off.append(off)
            c1 = capacity // g + 1
            off += 4 * (2 * c1 + len(w) * ((c1 + 31) // 32))
            off = (off + 255) & ~255
        self.n_videos = len(videos)
        self.total_seg, self.total_frames = cu_seg[-1], cu_frames[-1]
        self.dp_bytes = max(off, 256)
        self.cu_seg_host = np.asarray(cu_seg, dtype=np.int32)
        self.cu_frames_host = np.asarray(cu_frames, dtype=np.int64)
        self.lengths = [len(p) for p in picks]

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
        self.t = {
            "cu_seg": dev(cu_seg, np.int32), "cps": dev(np.concatenate(cps) if cps else np.zeros((0, 2)), np.int32),
            "nfps": dev(np.concatenate(nfps), np.int32), "picks": dev(np.concatenate(picks), np.int32),
            "cu_frames": dev(cu_frames, np.int64), "capacity": dev(cap, np.int32), "gcd": dev(gcds, np.int32),
            "dp_off": dev(dp_off, np.int64),
        }
        s = _capi.Shots()
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        self.struct = s


def keyshot_summaries(model, nms_out: dict, batch: DeviceBatch, shots: ShotPlan) -> dict:
    """Device tensors: summary uint8 [total_frames] (video v = cu_frames[v]..cu_frames[v+1]), picked uint8
    [total_seg], seg_scores int32 [total_seg], frame_scores / pos_scores float32."""
    if shots.lengths != [int(t) for t in batch.plan.lengths]:
        raise ValueError("picks do not match the batch's video lengths")
    dev = nms_out["keep_count"].device
    out = {
        "pos_scores": torch.empty(batch.plan.total_rows, dtype=torch.float32, device=dev),
        "frame_scores": torch.empty(max(shots.total_frames, 1), dtype=torch.float32, device=dev),
        "seg_scores": torch.empty(max(shots.total_seg, 1), dtype=torch.int32, device=dev),
        "picked": torch.empty(max(shots.total_seg, 1), dtype=torch.uint8, device=dev),
        "summary": torch.empty(max(shots.total_frames, 1), dtype=torch.uint8, device=dev),
    }
    scratch = torch.empty(shots.dp_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().edsnet_keyshot_summary(
            model._config(), batch.struct, shots.struct, nms_out["keep_count"].data_ptr(),
            nms_out["keep_scores"].data_ptr(), nms_out["keep_boxes"].data_ptr(), out["pos_scores"].data_ptr(),
            out["frame_scores"].data_ptr(), out["seg_scores"].data_ptr(), out["picked"].data_ptr(),
            out["summary"].data_ptr(), scratch.data_ptr(), stream))
    out["_scratch"] = scratch
    return out


def keyshot_from_scores(model, scores: torch.Tensor, batch: DeviceBatch, shots: ShotPlan) -> dict:
    """vsumm_helper.get_keyshot_summ (helpers/vsumm_helper.py:53-98) on given per-position scores, packed over the
    videos of `batch` (e.g. the ground-truth importance scores the training loop turns into targets,
    anchor_based/train.py:79).  scores: float32 [total_rows] on the device.  Same outputs as keyshot_summaries."""
    if not scores.is_cuda:
        raise RuntimeError("keyshot_from_scores needs CUDA tensors (there is no CPU fallback)")
    if shots.lengths != [int(t) for t in batch.plan.lengths]:
        raise ValueError("picks do not match the batch's video lengths")
    dev = scores.device
    out = {
        "pos_scores": scores.to(torch.float32).contiguous().clone(),
        "frame_scores": torch.empty(max(shots.total_frames, 1), dtype=torch.float32, device=dev),
        "seg_scores": torch.empty(max(shots.total_seg, 1), dtype=torch.int32, device=dev),
        "picked": torch.empty(max(shots.total_seg, 1), dtype=torch.uint8, device=dev),
        "summary": torch.empty(max(shots.total_frames, 1), dtype=torch.uint8, device=dev),
    }
    scratch = torch.empty(shots.dp_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().edsnet_keyshot_summary(
            model._config(), batch.struct, shots.struct, None, None, None, out["pos_scores"].data_ptr(),
            out["frame_scores"].data_ptr(), out["seg_scores"].data_ptr(), out["picked"].data_ptr(),
            out["summary"].data_ptr(), scratch.data_ptr(), stream))
    out["_scratch"] = scratch
    return out


def training_targets(model, gtscores: Sequence[np.ndarray], shots: ShotPlan, device) -> List[np.ndarray]:
    """The per-video target masks of anchor_based/train.py:79-84 (get_keyshot_summ on the ground-truth scores, then
    downsample_summ) for a whole split in one launch.  The reference recomputes them (ortools knapsack included)
    every step of every epoch although they only depend on the dataset: compute once, cache."""
    from .plan import BatchPlan
    lengths = [len(g) for g in gtscores]
    batch = BatchPlan.build(lengths).to(device)
    s = torch.from_numpy(np.concatenate([np.asarray(g, dtype=np.float32) for g in gtscores])).to(device)
    out = keyshot_from_scores(model, s, batch, shots)
    return [m[::15] for m in split_summaries(out["summary"], shots)]


def split_summaries(summary: torch.Tensor, shots: ShotPlan) -> List[np.ndarray]:
    s = summary.cpu().numpy().astype(bool)
    _capi.raise_on_tc_timeout()
    cf = shots.cu_frames_host
    return [s[cf[v]:cf[v + 1]] for v in range(shots.n_videos)]
