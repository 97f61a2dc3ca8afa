"""Per-stage CUDA-event times of the packed forward + decode + NMS (bench-shaped batch), a quick A/B tool:
    python tools/stage_times.py [videos] [precision] [steps]
Prints ms per forward for every stage and the un-instrumented forward time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from edsnet_b200 import BatchPlan, _capi  # noqa: E402


def main():
    videos = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda", 0)
    lengths = bench.workload_lengths(0, videos)
    R = int(sum(lengths))
    model = bench.xavier_state([12]).to(dev).eval()
    model.precision = prec
    x = bench.synth_features_device(R, dev, bench.SEED + 1000)
    plan = BatchPlan.build(lengths).to(dev)
    lib = _capi.lib()

    def step():
        cls, loc = model._forward_nograd(x, plan)
        model.nms_packed(cls, loc, plan, bench.NMS_THRESH)

    with torch.no_grad():
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        plain = e0.elapsed_time(e1) / steps
        _capi.check(lib.edsnet_debug_stage_timing(1))
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        st = _capi.stage_times()
        _capi.check(lib.edsnet_debug_stage_timing(0))
    assert lib.edsnet_debug_tc_status(0) == 0, "tcgen05 pipeline timeout flag is set"
    print(f"{videos} videos, {R} rows, {prec}: {plain:.3f} ms per forward+decode+NMS un-instrumented")
    tot = 0.0
    for name, (ms, n) in st.items():
        if n:
            print(f"  {name:18s} {ms / steps:8.4f} ms  ({n // steps} launches)")
            tot += ms / steps
    print(f"  {'sum':18s} {tot:8.4f} ms")


if __name__ == "__main__":
    main()
