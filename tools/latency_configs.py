"""Latency of the non-bench configurations of BASELINE.json on one B200 (CUDA events, median of N runs):
C1 (T=320, scales [12]), C4 (full-MHA base, T=2048), C5 (T=16384, scales [4,8,16,32]) -- forward and decode+NMS.
Weights / inputs: seeded synthetic (bench.py generators).  Usage: python tools/latency_configs.py [--cpu]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from edsnet_b200 import BatchPlan, DSNet
    dev = torch.device("cuda:0")
    out = {}
    for name, base, T, scales in (("C1", "nystromformer", 320, [12]), ("C4", "attention", 2048, [4, 8, 16, 32]),
                                  ("C5", "nystromformer", 16384, [4, 8, 16, 32]),
                                  ("C5_s12", "nystromformer", 16384, [12])):
        torch.manual_seed(bench.SEED)
        m = DSNet(base, 1024, 128, scales, 8, fc_depth=5, pooling_type="roi").to(dev).eval()
        x = bench.synth_features_device(T, dev, 1)
        db = BatchPlan.build([T]).to(dev)
        res = {}
        with torch.no_grad():
            for what in ("forward", "decode_nms"):
                ts = []
                for it in range(13):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    if what == "forward":
                        e0.record()
                        cls, loc = m._forward_nograd(x, db)
                        e1.record()
                    else:
                        e0.record()
                        r = m.nms_packed(cls, loc, db, 0.5)
                        e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                res[what + "_ms"] = float(np.median(ts[3:]))
            res["kept"] = int(r["keep_count"].cpu()[0])
            # the same forward replayed as a CUDA graph (launch latency removed)
            run = m.graphed_forward([T])
            ts = []
            for it in range(13):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run.graph.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res["forward_graph_ms"] = float(np.median(ts[3:]))
        out[name] = res
        print(name, res, flush=True)
    if "--cpu" in sys.argv:
        from oracle import dsnet_oracle as orc
        torch.set_num_threads(os.cpu_count())
        for name, base, T, scales in (("C1", "nystromformer", 320, [12]), ("C5", "nystromformer", 16384, [4, 8, 16, 32])):
            x = orc.synth_features(T, 1)
            p = orc.synth_params_mha(2, "xavier") if base == "attention" else orc.synth_params(2, "xavier")
            ts = []
            for it in range(4):
                t0 = time.perf_counter()
                with torch.no_grad():
                    c, l = orc.dsnet_forward(x, p, scales, 5, base=base)
                t1 = time.perf_counter()
                b = orc.clip_round(orc.decode_boxes(l.numpy(), T, scales), T)
                orc.nms_1d(c.numpy().reshape(-1), b, 0.5)
                t2 = time.perf_counter()
                ts.append((t1 - t0, t2 - t1))
            out[name + "_cpu"] = {"forward_ms": 1e3 * float(np.median([a for a, _ in ts[1:]])),
                                  "decode_nms_ms": 1e3 * float(np.median([b for _, b in ts[1:]])), "cores": os.cpu_count()}
            print(name + "_cpu", out[name + "_cpu"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
