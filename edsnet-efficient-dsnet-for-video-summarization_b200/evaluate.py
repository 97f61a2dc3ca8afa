"""Evaluation loop of the reference on the device (src/evaluate.py:14-40): features -> scores -> proposals -> NMS ->
keyshot summary -> F-score against the user summaries and diversity of the selected features, for a whole packed batch
of videos without a host round trip between the stages.

`TruthPlan` packs the user summaries the reference reads from its h5 files (`user_summary`, shape [users, n_frames]);
`eval_metrics` runs `edsnet_eval_metrics` on the output of `keyshot_summaries`; `evaluate` is the drop-in for the
reference's `evaluate(model, val_loader, nms_thresh, device)` over an iterable with the same 8-tuples.
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from .plan import BatchPlan, DeviceBatch
from .summary import ShotPlan, keyshot_summaries


class TruthPlan:
    def __init__(self, user_summaries: Sequence[np.ndarray], metrics: Sequence[str], device):
        """user_summaries: per video a 0/1 array [users, n_frames]; metrics: per video 'avg' or 'max'
        (evaluate.py:31 picks 'avg' for keys containing 'tvsum')."""
        if len(user_summaries) != len(metrics):
            raise ValueError("one metric per video")
        cu_users, user_off, frames, met, rows = [0], [], [], [], []
        off = 0
        for us, m in zip(user_summaries, metrics):
            us = np.asarray(us)
            if us.ndim != 2:
                raise ValueError("user_summary must be [users, n_frames]")
            if m not in ("avg", "max"):
                raise ValueError(f"Invalid eval metric {m}")            # vsumm_helper.py:170
            u, n = us.shape
            pitch = (n + 3) & ~3
            buf = np.zeros((u, pitch), dtype=np.uint8)
            buf[:, :n] = us.astype(bool)
            for r in range(u):
                user_off.append(off + r * pitch)
            off += u * pitch
            rows.append(buf.reshape(-1))
            cu_users.append(cu_users[-1] + u)
            frames.append(n)
            met.append(0 if m == "avg" else 1)
        self.n_videos = len(metrics)
        self.total_users = cu_users[-1]

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
        flat = np.concatenate(rows) if rows and off > 0 else np.zeros(4, dtype=np.uint8)
        self.t = {"cu_users": dev(cu_users, np.int32), "user_off": dev(user_off if user_off else [0], np.int64),
                  "user_frames": dev(frames, np.int32), "user_summ": dev(flat, np.uint8), "metric": dev(met, np.int32)}
        s = _capi.EvalTruth()
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        self.struct = s


def eval_metrics(x: torch.Tensor, batch: DeviceBatch, shots: ShotPlan, truth: TruthPlan, summary: torch.Tensor) -> dict:
    """Device tensors: fscore / diversity float64 [n_videos], user_f1 float64 [total_users], counts int32 [n_videos, 2]."""
    if not x.is_cuda:
        raise RuntimeError("eval_metrics needs CUDA tensors (there is no CPU fallback)")
    V = batch.plan.n_videos
    if truth.n_videos != V or shots.n_videos != V:
        raise ValueError("shot / truth tables do not match the batch")
    for v, t in enumerate(batch.plan.lengths):
        n_frames = int(shots.cu_frames_host[v + 1] - shots.cu_frames_host[v])
        # get_summ_diversity asserts len(summ[::15]) == len(features) (vsumm_helper.py:128)
        assert (n_frames + 14) // 15 == int(t), "down-sampled summary and features disagree in length"
    dev = x.device
    out = {"fscore": torch.empty(V, dtype=torch.float64, device=dev),
           "diversity": torch.empty(V, dtype=torch.float64, device=dev),
           "user_f1": torch.empty(max(truth.total_users, 1), dtype=torch.float64, device=dev),
           "counts": torch.empty((V, 2), dtype=torch.int32, device=dev)}
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().edsnet_eval_metrics(
            batch.struct, shots.t["cu_frames"].data_ptr(), summary.data_ptr(), truth.struct, x.data_ptr(),
            out["fscore"].data_ptr(), out["diversity"].data_ptr(), out["user_f1"].data_ptr(),
            out["counts"].data_ptr(), stream))
    return out


def evaluate(model, val_loader: Iterable, nms_thresh: float, device) -> Tuple[float, float]:
    """src/evaluate.py:14-40 for a loader yielding (key, seq, gtscore, cps, n_frames, nfps, picks, user_summary):
    mean F-score and mean diversity over the videos.  All videos are scored in one packed batch."""
    items = list(val_loader)
    if not items:
        return 0.0, 0.0
    model.eval()
    dev = torch.device(device)
    lengths = [len(it[1]) for it in items]
    x = torch.from_numpy(np.concatenate([np.asarray(it[1], dtype=np.float32) for it in items])).to(dev)
    batch = BatchPlan.build(lengths).to(dev)
    shots = ShotPlan([{"cps": it[3], "n_frames": it[4], "nfps": it[5], "picks": it[6]} for it in items], dev)
    truth = TruthPlan([it[7] for it in items], ["avg" if "tvsum" in str(it[0]) else "max" for it in items], dev)
    with torch.no_grad():
        cls, loc = model._forward_nograd(x, batch)
        nms = model.nms_packed(cls, loc, batch, nms_thresh)
        summ = keyshot_summaries(model, nms, batch, shots)
        met = eval_metrics(x, batch, shots, truth, summ["summary"])
    f = met["fscore"].cpu().numpy()
    d = met["diversity"].cpu().numpy()
    _capi.raise_on_tc_timeout()
    # data_helper.AverageMeter: running sums of python floats divided by the count
    return float(sum(float(v) for v in f) / len(f)), float(sum(float(v) for v in d) / len(d))
