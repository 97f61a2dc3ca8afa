"""BASELINE.json config 3: anchor-based training step (cls + loc loss) on a synthetic tvsum.yml-shaped split, data
parallel over the GPUs of one box (one video per GPU per optimiser step, ONE flat NCCL gradient all-reduce, identical
Adam update).  The reference's loop body is anchor_based/train.py:78-128; it has no collective (batch 1 on one device).

    python tools/train_step_bench.py [--steps 40] [--videos-per-rank 1] [--cpu]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py

Prints one JSON line (rank 0): ms per optimiser step (CUDA events, max over ranks), videos/s, the share of host label
generation, and with --cpu the same step on the host cores (torch CPU ops, the reference's arithmetic)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def synth_split(n_videos, rng):
    """tvsum.yml-shaped fold: 40 training videos, T ~ U[100, 800], ground-truth keyshot masks of 2-6 segments covering
    about 15 % of the positions (what get_keyshot_summ + downsample_summ produce, anchor_based/train.py:79-86)."""
    vids = []
    for i in range(n_videos):
        T = int(rng.integers(100, 801))
        mask = np.zeros(T, bool)
        budget = int(0.15 * T)
        for _ in range(int(rng.integers(2, 7))):
            ln = max(2, int(budget / 4 * rng.uniform(0.5, 1.5)))
            a = int(rng.integers(0, max(1, T - ln)))
            mask[a:a + ln] = True
        vids.append((T, mask))
    return vids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--videos-per-rank", type=int, default=1)
    ap.add_argument("--scales", type=int, nargs="+", default=[4, 8, 16, 32])
    ap.add_argument("--cpu", action="store_true", help="also time the same step on the host cores")
    ap.add_argument("--graphs", action="store_true", help="replay the step as CUDA graphs (GraphedDataParallelStep)")
    ap.add_argument("--cache-labels", action="store_true",
                    help="IoU sweeps once per video (training.LabelCache); only the negatives are drawn per step")
    args = ap.parse_args()
    rank, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from edsnet_b200 import training as tr

    model = bench.xavier_state(args.scales).to(dev)
    rng = np.random.default_rng(bench.SEED + rank)
    vids = synth_split(40, rng)
    feats = [bench.synth_features_device(T, dev, 7000 + 100 * rank + i) for i, (T, _) in enumerate(vids)]
    stepper = (tr.GraphedDataParallelStep if args.graphs else tr.DataParallelStep)(model, world_size=world)
    k = args.videos_per_rank
    label_s = [0.0]
    cache = tr.LabelCache(args.scales) if args.cache_labels else None
    if cache is not None:
        for i, (_, mask) in enumerate(vids):
            cache.add(i, mask)

    def one_step(i):
        sel = [(i * k + j) % len(vids) for j in range(k)]
        t0 = time.perf_counter()
        if cache is not None:
            labs = [cache.labels(s, rng) for s in sel]
        else:
            labs = [tr.anchor_labels(vids[s][1], args.scales, rng) for s in sel]   # host NumPy, as the reference
        label_s[0] += time.perf_counter() - t0
        cls_l = [torch.from_numpy(c).to(dev, non_blocking=True) for c, _ in labs]
        loc_l = [torch.from_numpy(l).float().to(dev, non_blocking=True) for _, l in labs]
        return stepper.step([feats[s] for s in sel], cls_l, loc_l)

    for i in range(max(args.warmup, 2 * len(vids) // k + 2 if args.graphs else 0)):     # graphs: every video seen twice
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    label_s[0] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    losses = [one_step(args.warmup + i) for i in range(args.steps)]
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    line = {"config": "C3: training step, tvsum.yml-shaped synthetic split (40 videos, T~U[100,800]), scales "
                      f"{args.scales}, {k} video(s) per GPU per step, Adam lr 5e-5 wd 1e-5",
            "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
            "videos_per_sec": world * k * args.steps / (ms * 1e-3),
            "host_label_ms_per_step": 1e3 * label_s[0] / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
            "loss_first": float(np.mean(losses[:5])), "loss_last": float(np.mean(losses[-5:])),
            "grad_allreduce": "one flat fp32 bucket (NCCL)" if world > 1 else "none (1 GPU)",
            "cuda_graphs": bool(args.graphs)}
    if args.cpu and rank == 0 and world == 1:
        cpu_model = bench.xavier_state(args.scales)
        cpu_model.train()
        torch.set_num_threads(os.cpu_count())
        from edsnet_b200 import autograd as ag
        opt = torch.optim.Adam(cpu_model.parameters(), lr=5e-5, weight_decay=1e-5)
        rng2 = np.random.default_rng(1)
        xs = [f.cpu() for f in feats[:8]]
        ts = []
        for i in range(10):
            s = i % 8
            t1 = time.perf_counter()
            c, l = tr.anchor_labels(vids[s][1], args.scales, rng2)
            pc, pl = ag.scoring_with_grad(cpu_model, xs[s], [vids[s][0]])
            loss = tr.cls_loss(pc, torch.from_numpy(c)) + tr.loc_loss(pl, torch.from_numpy(l).float(), torch.from_numpy(c))
            opt.zero_grad()
            loss.backward()
            opt.step()
            ts.append(time.perf_counter() - t1)
        line["cpu_ms_per_step"] = 1e3 * float(np.median(ts[2:]))
        line["cpu_cores"] = os.cpu_count()
    if rank == 0:
        bench.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
