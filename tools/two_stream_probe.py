"""Device-resident C2 step with its two launch sequences on ONE stream (the bench's way) against TWO streams (the kernels
of the two chunks may then share the machine wherever their resources allow): ms per step and identity of the kept sets.
python tools/two_stream_probe.py [videos] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from edsnet_b200 import BatchPlan, ScoringPipeline  # noqa: E402


def main():
    videos = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda", 0)
    lengths = bench.workload_lengths(0, videos)
    R = int(sum(lengths))
    model = bench.xavier_state([12]).to(dev).eval()
    model.precision = "fp16x2"
    x = bench.synth_features_device(R, dev, bench.SEED + 1000)
    pipe = ScoringPipeline(model, nms_thresh=bench.NMS_THRESH)
    chunks = pipe.chunk_videos(lengths, 1048576)
    cu = np.concatenate([[0], np.cumsum(lengths)])
    plans = [BatchPlan.build(lengths[a:b]).to(dev) for a, b in chunks]
    streams = [torch.cuda.Stream(dev) for _ in chunks]

    def step(two):
        outs = []
        main_s = torch.cuda.current_stream(dev)
        with torch.no_grad():
            for i, ((a, b), dp) in enumerate(zip(chunks, plans)):
                xd = x[int(cu[a]):int(cu[b])]
                if two:
                    streams[i].wait_stream(main_s)
                    with torch.cuda.stream(streams[i]):
                        cls, loc = model._forward_nograd(xd, dp)
                        outs.append(model.nms_packed(cls, loc, dp, bench.NMS_THRESH))
                else:
                    cls, loc = model._forward_nograd(xd, dp)
                    outs.append(model.nms_packed(cls, loc, dp, bench.NMS_THRESH))
            if two:
                for s in streams:
                    main_s.wait_stream(s)
        return outs

    res = {}
    for two in (False, True, False, True):
        for _ in range(3):
            outs = step(two)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            outs = step(two)
        e1.record()
        torch.cuda.synchronize()
        kept = [int(o["keep_count"].sum()) for o in outs]
        res.setdefault(two, []).append(e0.elapsed_time(e1) / steps)
        print(f"{'two streams' if two else 'one stream '}: {e0.elapsed_time(e1) / steps:.3f} ms per step, kept {kept}", flush=True)


if __name__ == "__main__":
    main()
