"""Kernel temporal segmentation on the device: the shot boundaries the reference derives from the sub-sampled features
before it scores a raw video (src/helpers/video_helper.py:109-126 -> src/kts/cpd_auto.py:6-33 ->
src/kts/cpd_nonlin.py:4-92).  `kts_change_points` runs `edsnet_kts` over a packed list of videos (chunked so that the
per-video n x n scratch blocks fit a byte budget); `kts_shots` returns what `VideoPreprocessor.kts` returns.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from .plan import BatchPlan

_VIDEO_DT = np.dtype([("scratch_off", np.int64), ("row0", np.int32), ("n", np.int32)])


def gram_tensor_core(xv: torch.Tensor, out: torch.Tensor) -> None:
    """out (n x n float32 view, row-major) <- X X^T of one video's features (video_helper.py:117) on the tcgen05 GEMM:
    the rows go through the operand-plane gather (`edsnet_cnn_im2col` as a 1 x 1 gather: fp16 hi / lo planes, one
    power-of-two scale per row) and ONE product with A = B = those planes, three split-fp16 passes (the entries agree
    with the float64 product to ~1e-6: the truncating fp32 accumulation of the tensor core).  Rows are padded to
    the GEMM's 128-column granularity; torch only moves data here."""
    import ctypes as C
    from .features import _Act
    lib = _capi.lib()
    n, f = int(xv.shape[0]), int(xv.shape[1])
    npad = (n + 127) // 128 * 128
    dev = xv.device
    xp = torch.zeros((npad, f), dtype=torch.float32, device=dev)
    xp[:n] = xv
    planes = torch.empty(int(lib.edsnet_split_f16_bytes(npad, f)), dtype=torch.uint8, device=dev)
    cpad = torch.empty((npad, npad), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    ci = _Act([(xp, f, 0, f)], 1, npad, 1, relu=False).struct()
    _capi.check(lib.edsnet_cnn_im2col(C.byref(ci), 1, npad, 1, 1, 1, 1, 0, f, planes.data_ptr(), st))
    _capi.check(lib.edsnet_gemm(_capi.PREC_FP16X3, 0, None, planes.data_ptr(), None, planes.data_ptr(), cpad.data_ptr(),
                                npad, npad, f, None, None, 0, st))
    out.copy_(cpad[:n, :n])


def kts_change_points(x: torch.Tensor, lengths: Sequence[int], kernels: Optional[Sequence[np.ndarray]] = None,
                      ncp_cap: int = -1, m_fixed: int = -1, vmax: float = 1.0, desc_rate: int = 1, lmin: int = 1,
                      lmax: int = 100000, scratch_budget: int = 4 << 30, gram: str = "tensor"
                      ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """x: packed float32 [sum(lengths), 1024] features on a CUDA device.  kernels: optional per-video float32 n x n
    kernel matrices to segment instead of X X^T (bit-exact comparison against the reference needs the reference's own
    np.matmul result).  gram: 'tensor' = X X^T on the tcgen05 GEMM (gram_tensor_core), 'fp32' = the CUDA-core kernel inside
    edsnet_kts.  Returns (change points per video, objective values for 0..len(cps) change points per video)."""
    if gram not in ("tensor", "fp32"):
        raise ValueError("gram must be 'tensor' or 'fp32'")
    if not x.is_cuda:
        raise RuntimeError("kts_change_points needs CUDA tensors (there is no CPU fallback)")
    lengths = [int(t) for t in lengths]
    if x.shape[0] != sum(lengths):
        raise RuntimeError("x rows do not match sum(lengths)")
    for n in lengths:
        m = m_fixed if m_fixed >= 0 else (min(ncp_cap, n - 1) if ncp_cap >= 0 else n - 1)
        # cpd_nonlin.py:50-51
        assert (m + 1) * lmin <= n <= (m + 1) * lmax, "Kernel matrix awaited / segment length bounds"
    lib = _capi.lib()
    dev = x.device
    x = x.detach().contiguous()
    cps_out: List[np.ndarray] = [None] * len(lengths)
    obj_out: List[np.ndarray] = [None] * len(lengths)
    cu = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    start = 0
    while start < len(lengths):
        end, total = start, 0
        while end < len(lengths):
            b = int(lib.edsnet_kts_scratch_bytes(lengths[end]))
            if end > start and total + b > scratch_budget:
                break
            total += b
            end += 1
        sub = lengths[start:end]
        plan = BatchPlan.build(sub)
        batch = plan.to(dev)
        vids = np.zeros(len(sub), dtype=_VIDEO_DT)
        off = 0
        for i, n in enumerate(sub):
            vids[i] = (off, int(plan.cu_rows[i]), n)
            off += int(lib.edsnet_kts_scratch_bytes(n))
        scratch = torch.empty(max(off, 256), dtype=torch.uint8, device=dev)
        if kernels is not None:
            for i, n in enumerate(sub):
                k = torch.from_numpy(np.ascontiguousarray(kernels[start + i], dtype=np.float32)).to(dev)
                o = int(vids[i]["scratch_off"])
                scratch[o:o + n * n * 4].view(torch.float32).copy_(k.reshape(-1))
        x_arg = None
        if kernels is None and gram == "tensor" and x.shape[1] % 64 == 0:
            with torch.cuda.device(dev):
                for i, n in enumerate(sub):
                    o, r0 = int(vids[i]["scratch_off"]), int(cu[start]) + int(plan.cu_rows[i])
                    gram_tensor_core(x[r0:r0 + n], scratch[o:o + n * n * 4].view(torch.float32).view(n, n))
        elif kernels is None:
            x_arg = "features"
        vids_dev = torch.from_numpy(vids.view(np.uint8)).to(dev)
        R = int(sum(sub))
        n_cps = torch.zeros(len(sub), dtype=torch.int32, device=dev)
        cps = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
        obj = torch.zeros(max(R, 1), dtype=torch.float64, device=dev)
        xs = x[int(cu[start]):int(cu[end])]
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(lib.edsnet_kts(batch.struct, vids_dev.data_ptr(), xs.data_ptr() if x_arg else None,
                                       ncp_cap, m_fixed, float(vmax), desc_rate, lmin, lmax, n_cps.data_ptr(),
                                       cps.data_ptr(), obj.data_ptr(), scratch.data_ptr(), stream))
        nc, cp, ob = n_cps.cpu().numpy(), cps.cpu().numpy(), obj.cpu().numpy()
        for i in range(len(sub)):
            o = int(plan.cu_rows[i])
            cps_out[start + i] = cp[o:o + int(nc[i])].astype(np.int64)
            obj_out[start + i] = ob[o:o + int(nc[i]) + 1].copy()
        start = end
    return cps_out, obj_out


def kts_shots(n_frames: int, features: torch.Tensor, sample_rate: int = 15):
    """VideoPreprocessor.kts (helpers/video_helper.py:109-126): (change_points [n_seg, 2] inclusive frame ranges,
    frames per segment, picks) from the sub-sampled features of one video."""
    T = int(features.shape[0])
    picks = np.arange(0, T) * sample_rate
    (cps,), _ = kts_change_points(features, [T])
    cps = cps * sample_rate
    cps = np.hstack((0, cps, n_frames))
    begin, end = cps[:-1], cps[1:]
    return np.vstack((begin, end - 1)).T, end - begin, picks
