"""SURVEY 8(e) parity check of the data-parallel training step on W GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 tools/train_parity_ddp.py

Every rank takes ONE seeded video, computes its gradient with the native kernels (Dropout off), the flat 2 248 843-value
buffer is all-reduced over NCCL and divided by W; the result must equal the mean of the W single-video gradients of the
staged float64 model (oracle/backward_model.py, pinned to the reference's autograd) on all 16 parameter tensors to 1e-4.
Prints one JSON line (rank 0) -- copied to profiles/.  The oracle is used as the checker only."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import backward_model as bm      # noqa: E402
from oracle import dsnet_oracle as orc       # noqa: E402
from tests.test_gpu_training import FIELD2NAME, _labels, _oracle_mean_grads      # noqa: E402
from tests.util import make_model            # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from edsnet_b200 import training as tr
    scales, depth = [4, 8, 16, 32], 5
    p = orc.synth_params(301, "xavier")
    lengths = [int(t) for t in np.random.default_rng(7).integers(100, 801, size=world)]
    xs = [orc.synth_features(t, 310 + i) for i, t in enumerate(lengths)]
    labels = [_labels(t, len(scales), 320 + i) for i, t in enumerate(lengths)]
    model = make_model(p, scales, depth, "fp16x3", dev)
    stepper = tr.NativeDataParallelStep(model, world_size=world, dropout=False)
    loss = stepper.backward_only([xs[rank].to(dev)], [labels[rank][0].numpy()], [labels[rank][1].numpy()])
    if world > 1:
        dist.all_reduce(stepper.flat_grad, op=dist.ReduceOp.SUM)
    stepper.flat_grad /= world
    torch.cuda.synchronize(dev)
    torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
    want, losses, _ = _oracle_mean_grads(xs, p, scales, depth, labels)          # mean of the W single-video gradients
    errs = {FIELD2NAME[f]: orc.rel_l2(stepper.grad_views[f].cpu().numpy(), want[FIELD2NAME[f]].numpy())
            for f in stepper.grad_views}
    worst = max(errs.values())
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world_size": world, "video_lengths": lengths, "allreduce_bytes": stepper.n_params * 4,
                          "loss_rank0": float(loss[0, 0]), "oracle_loss_rank0": losses[0][0],
                          "worst_rel_l2_over_ranks": float(t.item()), "rel_l2_per_tensor_rank0": errs, "bar": 1e-4,
                          "ok": bool(t.item() < 1e-4)}))
    if world > 1:
        dist.destroy_process_group()
    assert worst < 1e-4, errs


if __name__ == "__main__":
    main()
