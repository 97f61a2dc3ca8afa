"""Frame-feature extractor (GoogLeNet pool5, features.py) throughput: frames per second on cuda:0 for a batch of
preprocessed 224 x 224 frames (CUDA events, warm-up, inputs resident), the CPU oracle on the host cores next to it.
    python tools/pool5_bench.py [frames per batch] [precision] [--cpu]
Algorithmic work: 1.50 GMAC = 3.0 GFLOP per frame (57 convolutions)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edsnet_b200 import GoogLeNetPool5  # noqa: E402
from oracle import googlenet_oracle as gno  # noqa: E402


def flops_per_frame():
    # output pixels of every convolution at 224 x 224 input
    hw = {"conv1": 112 * 112, "conv2": 56 * 56, "conv3": 56 * 56}
    for name, *_ in gno.INCEPTIONS:
        side = 28 if name.startswith("inception3") else 14 if name.startswith("inception4") else 7
        for b in ("branch1", "branch2.0", "branch2.1", "branch3.0", "branch3.1", "branch4.1"):
            hw[f"{name}.{b}"] = side * side
    return sum(2.0 * co * ci * k * k * hw[n] for n, (co, ci, k) in gno.conv_shapes().items())


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 64
    prec = args[1] if len(args) > 1 else "fp16x3"
    dev = torch.device("cuda", 0)
    p = gno.synth_googlenet_params(1)
    net = GoogLeNetPool5(p, dev, precision=prec)
    x = gno.synth_frames(n, 2).to(dev)
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fl = flops_per_frame()
    line = {"metric": "pool5_frames_per_sec", "value": n / ms * 1e3, "unit": "frames/s", "ms_per_batch": ms,
            "frames_per_batch": n, "dtype": prec, "launches_per_batch": net.launches,
            "algorithmic_gflop_per_frame": fl / 1e9, "achieved_TFLOPs": fl * n / ms / 1e9}
    if "--cpu" in sys.argv:
        m = 8
        xc = gno.synth_frames(m, 2)
        with torch.no_grad():
            gno.pool5_features(xc[:2], p)
            t0 = time.perf_counter()
            gno.pool5_features(xc, p)
            dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{m} frames, torch-CPU restatement of the torchvision module"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
