"""One forward + decode + NMS of a bench-shaped packed batch between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full --import-source on -o ... python tools/profile_forward.py [videos] [precision]`
(the kernel list is then exactly one forward, independent of how many launches the weight preparation needs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from edsnet_b200 import BatchPlan  # noqa: E402


def main():
    videos = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    lengths = bench.workload_lengths(0, videos)
    R = int(sum(lengths))
    model = bench.xavier_state([12]).to(dev).eval()
    model.precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
    x = bench.synth_features_device(R, dev, bench.SEED + 1000)
    plan = BatchPlan.build(lengths).to(dev)
    with torch.no_grad():
        for _ in range(3):
            cls, loc = model._forward_nograd(x, plan)
            model.nms_packed(cls, loc, plan, bench.NMS_THRESH)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        cls, loc = model._forward_nograd(x, plan)
        model.nms_packed(cls, loc, plan, bench.NMS_THRESH)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print(f"profiled one forward of {videos} videos, {R} rows, precision {model.precision}")


if __name__ == "__main__":
    main()
