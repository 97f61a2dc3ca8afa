"""CPU restatement of the frame-feature extractor in front of the scoring path (TEST INFRASTRUCTURE ONLY: imported by
tests/, never by the product package).

What it follows (paths relative to /root/reference/src): helpers/video_helper.py:27-73 -- `FeatureExtractor('google-net')`
wraps torchvision's `models.googlenet` without its last two children (Dropout, fc) in eval mode, and `run()` flattens the
(1, 1024, 1, 1) output and divides it by its L2 norm + 1e-10.  The network itself is the third-party torchvision
(0.26 in this image; the reference pins nothing): GoogLeNet / Inception / BasicConv2d of
torchvision/models/googlenet.py -- conv (no bias) + BatchNorm2d(eps=0.001) + ReLU; the "5x5" branch really is 3x3; every
MaxPool2d has ceil_mode=True; `transform_input` is NOT applied because the reference calls the children through an
nn.Sequential.  Restated here with torch.nn.functional ops on the parameter names of torchvision's state dict.

Parity: PINNED to the real torchvision module as the reference wraps it (tests/golden/make_golden_googlenet.py runs
`nn.Sequential(*list(models.googlenet(...).children())[:-2])` on seeded inputs and weights and commits the features).
The PIL resize / crop / normalise in front (video_helper.py:28-33) is host-side image I/O and is not restated.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 0.001

# name, in, ch1x1, ch3x3red, ch3x3, ch5x5red, ch5x5, pool_proj   (torchvision/models/googlenet.py)
INCEPTIONS: List[Tuple[str, int, int, int, int, int, int, int]] = [
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


def conv_shapes() -> Dict[str, Tuple[int, int, int]]:
    """BasicConv2d name -> (out channels, in channels, kernel size)."""
    shapes = {"conv1": (64, 3, 7), "conv2": (64, 64, 1), "conv3": (192, 64, 3)}
    for name, cin, c1, c3r, c3, c5r, c5, pp in INCEPTIONS:
        shapes[f"{name}.branch1"] = (c1, cin, 1)
        shapes[f"{name}.branch2.0"] = (c3r, cin, 1)
        shapes[f"{name}.branch2.1"] = (c3, c3r, 3)
        shapes[f"{name}.branch3.0"] = (c5r, cin, 1)
        shapes[f"{name}.branch3.1"] = (c5, c5r, 3)
        shapes[f"{name}.branch4.1"] = (pp, cin, 1)
    return shapes


def synth_googlenet_params(seed: int) -> Dict[str, torch.Tensor]:
    """Seeded parameters under torchvision's state-dict names (He-scaled convolutions so that activations keep their
    magnitude through 22 layers; non-trivial BatchNorm statistics so that the folding is exercised)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for name, (co, ci, k) in conv_shapes().items():
        p[f"{name}.conv.weight"] = torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5
        p[f"{name}.bn.weight"] = 0.5 + torch.rand(co, generator=g)
        p[f"{name}.bn.bias"] = 0.1 * torch.randn(co, generator=g)
        p[f"{name}.bn.running_mean"] = 0.1 * torch.randn(co, generator=g)
        p[f"{name}.bn.running_var"] = 0.5 + torch.rand(co, generator=g)
    return p


def synth_frames(n: int, seed: int) -> torch.Tensor:
    """n preprocessed frames (3, 224, 224): smooth random images, roughly the range ImageNet normalisation produces."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(n, 3, 14, 14, generator=g)
    x = F.interpolate(low, size=(224, 224), mode="bilinear", align_corners=False)
    return (x + 0.25 * torch.randn(n, 3, 224, 224, generator=g)).contiguous()


def basic_conv(x: torch.Tensor, p: Dict[str, torch.Tensor], name: str, stride: int = 1, padding: int = 0) -> torch.Tensor:
    y = F.conv2d(x, p[f"{name}.conv.weight"], None, stride=stride, padding=padding)
    y = F.batch_norm(y, p[f"{name}.bn.running_mean"], p[f"{name}.bn.running_var"], p[f"{name}.bn.weight"],
                     p[f"{name}.bn.bias"], training=False, eps=BN_EPS)
    return F.relu(y)


def inception(x: torch.Tensor, p: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    b1 = basic_conv(x, p, f"{name}.branch1")
    b2 = basic_conv(basic_conv(x, p, f"{name}.branch2.0"), p, f"{name}.branch2.1", padding=1)
    b3 = basic_conv(basic_conv(x, p, f"{name}.branch3.0"), p, f"{name}.branch3.1", padding=1)
    b4 = basic_conv(F.max_pool2d(x, 3, stride=1, padding=1, ceil_mode=True), p, f"{name}.branch4.1")
    return torch.cat([b1, b2, b3, b4], dim=1)


def pool5_raw(x: torch.Tensor, p: Dict[str, torch.Tensor], stages: dict | None = None) -> torch.Tensor:
    """(N, 3, H, W) -> (N, 1024): the children of googlenet up to avgpool, in order."""
    x = basic_conv(x, p, "conv1", stride=2, padding=3)
    x = F.max_pool2d(x, 3, stride=2, ceil_mode=True)
    if stages is not None:
        stages["maxpool1"] = x
    x = basic_conv(x, p, "conv2")
    x = basic_conv(x, p, "conv3", padding=1)
    x = F.max_pool2d(x, 3, stride=2, ceil_mode=True)
    x = inception(x, p, "inception3a")
    if stages is not None:
        stages["inception3a"] = x
    x = inception(x, p, "inception3b")
    x = F.max_pool2d(x, 3, stride=2, ceil_mode=True)
    for name in ("inception4a", "inception4b", "inception4c", "inception4d", "inception4e"):
        x = inception(x, p, name)
    x = F.max_pool2d(x, 2, stride=2, ceil_mode=True)
    x = inception(x, p, "inception5a")
    x = inception(x, p, "inception5b")
    return F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)


def pool5_features(x: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    """FeatureExtractor.run for a batch of preprocessed frames: pool5, then feat / (|feat| + 1e-10) per frame."""
    f = pool5_raw(x, p)
    return f / (f.norm(dim=1, keepdim=True) + 1e-10)
