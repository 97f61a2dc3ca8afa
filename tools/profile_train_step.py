"""One native training step (BASELINE.json config 3: one tvsum-shaped video, scales [4, 8, 16, 32], Dropout on) between
cudaProfilerStart/Stop, launched kernel by kernel (no CUDA graph) so that
`ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/profile_train_step.py [T]` lists every
launch of the forward, the loss gradient, the backward and Adam."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from edsnet_b200 import training as tr
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    scales = [4, 8, 16, 32]
    model = bench.xavier_state(scales).to(dev)
    x = bench.synth_features_device(T, dev, 7)
    mask = np.zeros(T, bool)
    mask[T // 8:T // 8 + T // 20] = True
    mask[T // 2:T // 2 + T // 16] = True
    cache = tr.LabelCache(scales)
    cache.add(0, mask)
    rng = np.random.default_rng(1)
    stepper = tr.NativeDataParallelStep(model, use_graphs=False)
    for _ in range(3):
        c, l = cache.labels(0, rng)
        stepper.step([x], [c], [l])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    c, l = cache.labels(0, rng)
    stepper.step([x], [c], [l])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profiled one training step, T = {T}, loss {stepper.last_loss():.4f}")


if __name__ == "__main__":
    main()
