"""Host-buffer -> proposals throughput path.

The reference's evaluation loop (evaluate.py:19-28) moves one video to the device, scores it, copies the
scores back and runs decode + NMS on the host, one video at a time.  `ScoringPipeline.run` does the same job for
a whole list of videos held in HOST memory: the videos are cut into chunks of at most `chunk_rows` feature rows;
chunk i+1's host->device copy (copy stream, pinned memory) overlaps chunk i's kernels (compute stream) and chunk
i-1's device->host copy of the kept proposals (copy-back stream).  Nothing is computed on the host.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .plan import BatchPlan


def bind_host_to_gpu_numa_node(device_index: int) -> List[int]:
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (same socket / PCIe root).
    Pinned host buffers allocated afterwards land in that socket's memory, so the host->device stream of every rank of
    a multi-GPU box stays off the inter-socket link.  Returns the core list ([] if NVML or the affinity call is
    unavailable: the pipeline then runs unbound, as before)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in range(n_cpu) if (int(words[c // 64]) >> (c % 64)) & 1 and c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


class ScoringPipeline:
    def __init__(self, model, chunk_rows: int = 32768, nms_thresh: float = 0.5, depth: int = 2):
        self.model = model
        self.chunk_rows = int(chunk_rows)
        self.nms_thresh = float(nms_thresh)
        self.depth = max(2, int(depth))
        self._dev_x = None
        self._streams = None
        self._out = None                 # pinned result buffers, reused while the batch geometry is unchanged
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.kernel_launches = 0

    # ---- chunking (host logic, unit-tested on CPU) ----
    @staticmethod
    def chunk_videos(lengths: Sequence[int], chunk_rows: int) -> List[Tuple[int, int]]:
        """Consecutive [first_video, last_video) ranges whose row sums do not exceed chunk_rows (a single video
        longer than chunk_rows forms its own chunk)."""
        out, start, rows = [], 0, 0
        for i, t in enumerate(lengths):
            t = int(t)
            if rows and rows + t > chunk_rows:
                out.append((start, i))
                start, rows = i, 0
            rows += t
        if rows:
            out.append((start, len(lengths)))
        return out

    def _setup(self, device, max_rows: int):
        if self._streams is None or self._streams[0].device != device:
            self._streams = (torch.cuda.Stream(device), torch.cuda.Stream(device), torch.cuda.Stream(device))
        if self._dev_x is None or self._dev_x[0].device != device or self._dev_x[0].shape[0] < max_rows:
            self._dev_x = [torch.empty((max_rows, 1024), dtype=torch.float32, device=device)
                           for _ in range(self.depth)]

    def run_copies_only(self, x_host: torch.Tensor, lengths: Sequence[int], device=None) -> None:
        """The host<->device traffic of `run` without any kernel: the same chunks on the same copy streams (H2D of
        the features and the batch tables, D2H of result-sized buffers).  Its duration is the floor the link and the
        host memory system put under the end-to-end number; bench.py reports e2e as a fraction of it."""
        model = self.model
        device = torch.device(device) if device is not None else next(model.parameters()).device
        lengths = [int(t) for t in lengths]
        S = model.num_scales
        V, R = len(lengths), int(sum(lengths))
        chunks = self.chunk_videos(lengths, self.chunk_rows)
        cu = np.zeros(V + 1, dtype=np.int64)
        cu[1:] = np.cumsum(lengths)
        self._setup(device, max(int(cu[b] - cu[a]) for a, b in chunks))
        s_in, _, s_out = self._streams
        if self._out is None or self._out[0].numel() != V or self._out[1].numel() != R * S:
            self._out = (torch.empty(V, dtype=torch.int32).pin_memory(),
                         torch.empty(R * S, dtype=torch.float32).pin_memory(),
                         torch.empty((R * S, 2), dtype=torch.int32).pin_memory())
        keep_count, keep_scores, keep_boxes = self._out
        max_rows = self._dev_x[0].shape[0]
        if getattr(self, "_floor_src", None) is None or self._floor_src[1].numel() < max_rows * S:
            self._floor_src = (torch.zeros(max(V, 1), dtype=torch.int32, device=device),
                               torch.zeros(max_rows * S, dtype=torch.float32, device=device),
                               torch.zeros((max_rows * S, 2), dtype=torch.int32, device=device))
        d_cnt, d_sc, d_bx = self._floor_src
        cur = torch.cuda.current_stream(device)
        for s in (s_in, s_out):
            s.wait_stream(cur)
        free_ev = [None] * self.depth
        pending = []
        for ci, (a, b) in enumerate(chunks):
            r0, r1 = int(cu[a]), int(cu[b])
            buf = self._dev_x[ci % self.depth]
            plan = BatchPlan.build(lengths[a:b])
            with torch.cuda.stream(s_in):
                if free_ev[ci % self.depth] is not None:
                    s_in.wait_event(free_ev[ci % self.depth])
                buf[: r1 - r0].copy_(x_host[r0:r1], non_blocking=True)
                dbatch = plan.to(device)
                ready = torch.cuda.Event()
                ready.record(s_in)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ready)
                keep_count[a:b].copy_(d_cnt[: b - a], non_blocking=True)
                keep_scores[r0 * S:r1 * S].copy_(d_sc[: (r1 - r0) * S], non_blocking=True)
                keep_boxes[r0 * S:r1 * S].copy_(d_bx[: (r1 - r0) * S], non_blocking=True)
                done = torch.cuda.Event()
                done.record(s_out)
                free_ev[ci % self.depth] = done
            pending.append(dbatch)
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)
        torch.cuda.current_stream(device).synchronize()

    def run(self, x_host: torch.Tensor, lengths: Sequence[int], device=None):
        """x_host: [sum(lengths), 1024] float32 in (preferably pinned) host memory.
        Returns (keep_count int32 [V], keep_scores float32 [R*S], keep_boxes int32 [R*S, 2], cu_rows int32 [V+1]) as
        pinned host tensors (owned by the pipeline and overwritten by the next run); video v's proposals are entries cu_rows[v]*S .. +keep_count[v], descending score --
        the (keep_scores, keep_boxes) pair evaluate.py:28 gets from bbox_helper.nms."""
        model = self.model
        device = torch.device(device) if device is not None else next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("ScoringPipeline needs the model on a CUDA device (no CPU path)")
        if x_host.is_cuda:
            raise RuntimeError("ScoringPipeline.run takes host features; use DSNet.proposals_packed for device data")
        lengths = [int(t) for t in lengths]
        S = model.num_scales
        V, R = len(lengths), int(sum(lengths))
        if x_host.shape[0] != R:
            raise RuntimeError("x_host rows do not match sum(lengths)")
        chunks = self.chunk_videos(lengths, self.chunk_rows)
        cu = np.zeros(V + 1, dtype=np.int64)
        cu[1:] = np.cumsum(lengths)
        max_rows = max(int(cu[b] - cu[a]) for a, b in chunks)
        self._setup(device, max_rows)
        s_in, s_cmp, s_out = self._streams
        if self._out is None or self._out[0].numel() != V or self._out[1].numel() != R * S:
            self._out = (torch.empty(V, dtype=torch.int32).pin_memory(),
                         torch.empty(R * S, dtype=torch.float32).pin_memory(),
                         torch.empty((R * S, 2), dtype=torch.int32).pin_memory())
        keep_count, keep_scores, keep_boxes = self._out
        free_ev = [None] * self.depth           # buffer reusable once its chunk's kernels are done
        self.h2d_bytes = self.d2h_bytes = self.kernel_launches = 0
        cur = torch.cuda.current_stream(device)
        for s in (s_in, s_cmp, s_out):
            s.wait_stream(cur)
        pending = []
        for ci, (a, b) in enumerate(chunks):
            r0, r1 = int(cu[a]), int(cu[b])
            buf = self._dev_x[ci % self.depth]
            plan = BatchPlan.build(lengths[a:b])
            with torch.cuda.stream(s_in):
                if free_ev[ci % self.depth] is not None:
                    s_in.wait_event(free_ev[ci % self.depth])
                xd = buf[: r1 - r0]
                xd.copy_(x_host[r0:r1], non_blocking=True)
                dbatch = plan.to(device)
                ready = torch.cuda.Event()
                ready.record(s_in)
            self.h2d_bytes += (r1 - r0) * 4096 + dbatch.tables.numel() * 4
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ready)
                with torch.no_grad():
                    cls, loc = model._forward_nograd(xd, dbatch)
                    res = model.nms_packed(cls, loc, dbatch, self.nms_thresh)
                done = torch.cuda.Event()
                done.record(s_cmp)
                free_ev[ci % self.depth] = done
            self.kernel_launches += model.launches_per_forward() + 2
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                keep_count[a:b].copy_(res["keep_count"], non_blocking=True)
                keep_scores[r0 * S:r1 * S].copy_(res["keep_scores"], non_blocking=True)
                keep_boxes[r0 * S:r1 * S].copy_(res["keep_boxes"], non_blocking=True)
                for t in (xd, cls, loc, *[v for v in res.values() if isinstance(v, torch.Tensor)]):
                    t.record_stream(s_out)
            self.d2h_bytes += (b - a) * 4 + (r1 - r0) * S * 12
            pending.append((dbatch, res, cls, loc))
        cur.wait_stream(s_out)
        cur.wait_stream(s_cmp)
        cur.wait_stream(s_in)
        torch.cuda.current_stream(device).synchronize()
        from . import _capi
        _capi.raise_on_tc_timeout()
        del pending
        return keep_count, keep_scores, keep_boxes, torch.from_numpy(cu.astype(np.int32))
