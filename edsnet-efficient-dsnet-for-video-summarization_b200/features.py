"""GoogLeNet pool5 frame features on the device: the step in front of the scoring path (second half).

What it mirrors in the reference (paths relative to /root/reference/src): `helpers/video_helper.py:27-73` --
`FeatureExtractor('google-net')`: torchvision's googlenet without its last two children (Dropout, fc) in eval mode,
`run()` = one preprocessed frame -> flatten -> `feat / (|feat| + 1e-10)`.  `GoogLeNetPool5` takes the state dict of that
torchvision module (same names: `conv1.conv.weight`, `inception3a.branch2.1.bn.running_var`, ...) and a batch of
preprocessed frames (N, 3, 224, 224) float32 on a CUDA device, and returns (N, 1024) normalised features.  Video decoding
and the PIL resize / crop / normalise of `video_helper.py:28-33,82-100` are host I/O and stay with the caller.

How it runs (no torch op computes any part of the network; torch provides device memory and streams):
  * every BasicConv2d = conv (no bias) + BatchNorm2d(eps 1e-3) + ReLU becomes ONE product on the tcgen05 GEMM
    (`edsnet_gemm`, three split-fp16 passes, fp32-grade): BatchNorm is folded into the weights and a bias once, the
    weights are laid out [C_out padded to 128][kh * kw * C_in padded to 64] with the channel fastest, the ReLU is applied by
    whoever reads the output;
  * `edsnet_cnn_im2col` gathers the patches of a layer's input straight into the GEMM's operand planes.  Its input is a
    virtual channel concatenation, so an inception module's four branch outputs are never copied together, and the
    three 1x1 convolutions that read the module input (branch1, branch2.0, branch3.0) are one product with their
    weights stacked;
  * `edsnet_cnn_maxpool` (ceil_mode) and `edsnet_cnn_avgpool_l2norm` are the remaining kernels.
57 convolutions = 39 products, 13 max-pools, 1 average-pool per batch of frames.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import torch

from . import _capi

BN_EPS = 0.001
_PREC = {"fp32": 0, "fp16x3": 1, "fp16": 2, "fp16x2": 3}

# name, in, ch1x1, ch3x3red, ch3x3, ch5x5red, ch5x5, pool_proj  (torchvision/models/googlenet.py; the "5x5" branch is 3x3)
_INCEPTIONS = [
    ("inception3a", 192, 64, 96, 128, 16, 32, 32),
    ("inception3b", 256, 128, 128, 192, 32, 96, 64),
    ("inception4a", 480, 192, 96, 208, 16, 48, 64),
    ("inception4b", 512, 160, 112, 224, 24, 64, 64),
    ("inception4c", 512, 128, 128, 256, 24, 64, 64),
    ("inception4d", 512, 112, 144, 288, 32, 64, 64),
    ("inception4e", 528, 256, 160, 320, 32, 128, 128),
    ("inception5a", 832, 256, 160, 320, 32, 128, 128),
    ("inception5b", 832, 384, 192, 384, 48, 128, 128),
]


def _ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def fold_batchnorm(sd: Dict[str, torch.Tensor], name: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """conv (no bias) + BatchNorm2d(eval) == conv with w * g / sqrt(var + eps) and bias b - mean * g / sqrt(var + eps).
    Returns (weight [C_out][kh][kw][C_in], bias [C_out]) in float64-folded float32."""
    w = sd[f"{name}.conv.weight"].double()
    g = sd[f"{name}.bn.weight"].double() / torch.sqrt(sd[f"{name}.bn.running_var"].double() + BN_EPS)
    b = sd[f"{name}.bn.bias"].double() - sd[f"{name}.bn.running_mean"].double() * g
    w = (w * g[:, None, None, None]).permute(0, 2, 3, 1).contiguous()
    return w.float(), b.float()


class _Act:
    """An activation: virtual channel concat of (buffer, ld, col0, channels) pieces of [pixels][ld] fp32 buffers."""

    def __init__(self, pieces, n, h, w, relu, nchw=False):
        self.pieces, self.n, self.h, self.w, self.relu, self.nchw = pieces, n, h, w, relu, nchw

    @property
    def channels(self) -> int:
        return sum(p[3] for p in self.pieces)

    def struct(self) -> _capi.CnnInput:
        ci = _capi.CnnInput()
        ci.n_src = len(self.pieces)
        ci.relu = 1 if self.relu else 0
        hw = self.h * self.w
        for i, (t, ld, col0, ch) in enumerate(self.pieces):
            s = ci.src[i]
            s.p = t.data_ptr()
            if self.nchw:                     # (N, C, H, W) contiguous
                s.image_stride, s.pixel_stride, s.channel_stride = ld * hw, 1, hw
            else:
                s.image_stride, s.pixel_stride, s.channel_stride = hw * ld, ld, 1
            s.col0, s.channels = col0, ch
        return ci


class _Conv:
    """One product: the stacked, BatchNorm-folded weights of 1..3 convolutions over the same input as operand planes."""

    def __init__(self, sd, names: Sequence[str], k: int, stride: int, pad: int, device):
        ws, bs = zip(*(fold_batchnorm(sd, n) for n in names))
        self.k, self.stride, self.pad = k, stride, pad
        self.splits = [int(w.shape[0]) for w in ws]
        cout = sum(self.splits)
        kk = int(ws[0][0].numel())
        self.kpad, self.npad = _ceil_to(kk, 64), _ceil_to(cout, 128)
        wm = torch.zeros(self.npad, self.kpad, dtype=torch.float32)
        wm[:cout, :kk] = torch.cat([w.reshape(w.shape[0], -1) for w in ws], dim=0)
        bias = torch.zeros(self.npad, dtype=torch.float32)
        bias[:cout] = torch.cat(bs)
        self.bias = bias.to(device)
        # weight planes through the same gather kernel: a 1x1 "convolution" over an npad-pixel image of kpad channels
        wm = wm.to(device)
        lib = _capi.lib()
        self.planes = torch.empty(int(lib.edsnet_split_f16_bytes(self.npad, self.kpad)), dtype=torch.uint8, device=device)
        act = _Act([(wm, self.kpad, 0, self.kpad)], 1, self.npad, 1, relu=False)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            ci = act.struct()
            _capi.check(lib.edsnet_cnn_im2col(C.byref(ci), 1, self.npad, 1, 1, 1, 1, 0, self.kpad, self.planes.data_ptr(), st))
        torch.cuda.current_stream(device).synchronize()      # wm may be freed


class GoogLeNetPool5:
    """Drop-in for the reference's FeatureExtractor('google-net') on preprocessed frames (see the module docstring)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, precision: str = "fp16x3"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("edsnet_b200 has no CPU path: GoogLeNetPool5 needs a CUDA device")
        if precision not in _PREC or precision == "fp32":
            raise RuntimeError("GoogLeNetPool5 runs on the tcgen05 GEMM: precision must be fp16x3, fp16x2 or fp16")
        self.precision = precision
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        d = self.device
        self.conv1 = _Conv(sd, ["conv1"], 7, 2, 3, d)
        self.conv2 = _Conv(sd, ["conv2"], 1, 1, 0, d)
        self.conv3 = _Conv(sd, ["conv3"], 3, 1, 1, d)
        self.inc = {}
        for name, *_ in _INCEPTIONS:
            self.inc[name] = (
                _Conv(sd, [f"{name}.branch1", f"{name}.branch2.0", f"{name}.branch3.0"], 1, 1, 0, d),
                _Conv(sd, [f"{name}.branch2.1"], 3, 1, 1, d),
                _Conv(sd, [f"{name}.branch3.1"], 3, 1, 1, d),
                _Conv(sd, [f"{name}.branch4.1"], 1, 1, 0, d),
            )
        self.launches = 0

    # ------------------------------------------------------------------ building blocks
    def _conv(self, act: _Act, cv: _Conv) -> Tuple[torch.Tensor, int, int]:
        """act -> raw (pre-ReLU) output buffer [pixels][npad], OH, OW."""
        lib = _capi.lib()
        oh = (act.h + 2 * cv.pad - cv.k) // cv.stride + 1
        ow = (act.w + 2 * cv.pad - cv.k) // cv.stride + 1
        m = act.n * oh * ow
        if cv.k * cv.k * act.channels > cv.kpad:
            raise RuntimeError("input channels do not match the convolution's weights")
        planes = torch.empty(int(lib.edsnet_split_f16_bytes(m, cv.kpad)), dtype=torch.uint8, device=self.device)
        out = torch.empty((m, cv.npad), dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        ci = act.struct()
        _capi.check(lib.edsnet_cnn_im2col(C.byref(ci), act.n, act.h, act.w, cv.k, cv.k, cv.stride, cv.pad, cv.kpad,
                                          planes.data_ptr(), st))
        _capi.check(lib.edsnet_gemm(_PREC[self.precision], 2, None, planes.data_ptr(), None, cv.planes.data_ptr(),
                                    out.data_ptr(), m, cv.npad, cv.kpad, cv.bias.data_ptr(), None, 0, st))
        self.launches += 2
        return out, oh, ow

    def _maxpool(self, act: _Act, k: int, stride: int, pad: int) -> _Act:
        lib = _capi.lib()

        def osz(h):
            o = (h + 2 * pad - k + stride - 1) // stride + 1
            return o - 1 if (o - 1) * stride >= h + pad else o
        oh, ow = osz(act.h), osz(act.w)
        c = act.channels
        out = torch.empty((act.n * oh * ow, c), dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        ci = act.struct()
        _capi.check(lib.edsnet_cnn_maxpool(C.byref(ci), act.n, act.h, act.w, k, stride, pad, out.data_ptr(), st))
        self.launches += 1
        return _Act([(out, c, 0, c)], act.n, oh, ow, relu=False)      # the ReLU went in with the values read

    def _inception(self, act: _Act, name: str) -> _Act:
        fused, c2, c3, c4 = self.inc[name]
        c1, c2r, c3r = fused.splits
        f, _, _ = self._conv(act, fused)
        b2, _, _ = self._conv(_Act([(f, fused.npad, c1, c2r)], act.n, act.h, act.w, relu=True), c2)
        b3, _, _ = self._conv(_Act([(f, fused.npad, c1 + c2r, c3r)], act.n, act.h, act.w, relu=True), c3)
        b4, _, _ = self._conv(self._maxpool(act, 3, 1, 1), c4)
        return _Act([(f, fused.npad, 0, c1), (b2, c2.npad, 0, c2.splits[0]), (b3, c3.npad, 0, c3.splits[0]),
                     (b4, c4.npad, 0, c4.splits[0])], act.n, act.h, act.w, relu=True)

    # ------------------------------------------------------------------ the network
    def __call__(self, frames: torch.Tensor) -> torch.Tensor:
        """frames: (N, 3, H, W) float32 on the module's CUDA device (preprocessed as video_helper.py:28-33 does)
        -> (N, 1024) float32, every row divided by its L2 norm + 1e-10 (video_helper.py:66-72)."""
        if not isinstance(frames, torch.Tensor) or not frames.is_cuda:
            raise RuntimeError("edsnet_b200 has no CPU path: move the frames to a CUDA device")
        if frames.dtype != torch.float32 or frames.dim() != 4 or frames.shape[1] != 3:
            raise RuntimeError("expected float32 frames of shape (N, 3, H, W)")
        frames = frames.contiguous()
        n, _, h, w = (int(v) for v in frames.shape)
        lib = _capi.lib()
        self.launches = 0
        with torch.cuda.device(self.device):
            act = _Act([(frames, 3, 0, 3)], n, h, w, relu=False, nchw=True)
            y, h, w = self._conv(act, self.conv1)
            act = self._maxpool(_Act([(y, self.conv1.npad, 0, 64)], n, h, w, relu=True), 3, 2, 0)
            y, h, w = self._conv(act, self.conv2)
            y, h, w = self._conv(_Act([(y, self.conv2.npad, 0, 64)], n, h, w, relu=True), self.conv3)
            act = self._maxpool(_Act([(y, self.conv3.npad, 0, 192)], n, h, w, relu=True), 3, 2, 0)
            act = self._inception(act, "inception3a")
            act = self._inception(act, "inception3b")
            act = self._maxpool(act, 3, 2, 0)
            for name in ("inception4a", "inception4b", "inception4c", "inception4d", "inception4e"):
                act = self._inception(act, name)
            act = self._maxpool(act, 2, 2, 0)
            act = self._inception(act, "inception5a")
            act = self._inception(act, "inception5b")
            out = torch.empty((n, 1024), dtype=torch.float32, device=self.device)
            st = torch.cuda.current_stream(self.device).cuda_stream
            ci = act.struct()
            _capi.check(lib.edsnet_cnn_avgpool_l2norm(C.byref(ci), n, act.h * act.w, out.data_ptr(), st))
            self.launches += 1
        return out
