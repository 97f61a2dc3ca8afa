"""The inference chain of the reference's infer.py (src/infer.py:22-36) from the sub-sampled features onwards, on the
device: temporal segmentation (VideoPreprocessor.kts) -> scores -> decode / clip / round -> NMS -> keyshot summary;
`summarize_frames` starts one step earlier, at the sampled, preprocessed frames (GoogLeNet pool5 features on the device,
features.py).  Video decoding / PIL preprocessing in front (helpers/video_helper.py:28-33,82-100) and the video writer
behind it are host I/O and not part of this package.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from .kts import kts_change_points
from .plan import BatchPlan
from .summary import ShotPlan, keyshot_summaries, split_summaries


def summarize(model, features: torch.Tensor, lengths: Sequence[int], n_frames: Sequence[int], nms_thresh: float = 0.5,
              sample_rate: int = 15) -> List[dict]:
    """features: packed float32 [sum(lengths), 1024] on the model's CUDA device (one row per sampled frame);
    n_frames: original frame count of every video.  Returns per video a dict with `summary` (bool [n_frames], what
    infer.py:35-36 calls pred_summ), `change_points`, `nfps`, `picks` (video_helper.py:119-126)."""
    lengths = [int(t) for t in lengths]
    if not features.is_cuda:
        raise RuntimeError("summarize needs CUDA tensors (there is no CPU fallback)")
    dev = features.device
    model.eval()
    cps, _ = kts_change_points(features, lengths)
    shots = []
    for t, nf, cp in zip(lengths, n_frames, cps):
        cp = np.hstack((0, cp * sample_rate, int(nf)))
        begin, end = cp[:-1], cp[1:]
        shots.append({"cps": np.vstack((begin, end - 1)).T, "nfps": end - begin,
                      "picks": np.arange(0, t) * sample_rate, "n_frames": int(nf)})
    batch = BatchPlan.build(lengths).to(dev)
    plan = ShotPlan(shots, dev)
    with torch.no_grad():
        cls, loc = model._forward_nograd(features, batch)
        nms = model.nms_packed(cls, loc, batch, nms_thresh)
        out = keyshot_summaries(model, nms, batch, plan)
    summ = split_summaries(out["summary"], plan)
    return [{"summary": s, "change_points": sh["cps"], "nfps": sh["nfps"], "picks": sh["picks"]}
            for s, sh in zip(summ, shots)]


def summarize_frames(model, extractor, frames: Sequence[torch.Tensor], n_frames: Sequence[int], nms_thresh: float = 0.5,
                     sample_rate: int = 15, frame_batch: int = 64) -> List[dict]:
    """VideoPreprocessor.run + the rest of infer.py (video_helper.py:128-131, infer.py:26-36) for a list of videos given
    as their SAMPLED, preprocessed frames: frames[v] is (T_v, 3, H, W) float32 on the device (every sample_rate-th frame,
    resized / cropped / normalised as video_helper.py:28-33 does).  extractor: a features.GoogLeNetPool5.  Returns what
    `summarize` returns plus the `features` of every video."""
    lengths = [int(f.shape[0]) for f in frames]
    feats = []
    for f in frames:
        for s in range(0, f.shape[0], frame_batch):
            feats.append(extractor(f[s:s + frame_batch]))
    features = torch.cat(feats, dim=0)
    out = summarize(model, features, lengths, n_frames, nms_thresh, sample_rate)
    cu = np.concatenate([[0], np.cumsum(lengths)])
    for v, d in enumerate(out):
        d["features"] = features[int(cu[v]):int(cu[v + 1])]
    return out
