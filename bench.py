#!/usr/bin/env python
"""Benchmark of the EDSNet anchor-based scoring path on B200 (BASELINE.json configs[1]):
batched inference (forward + decode + temporal NMS) over 4096 synthetic TVSum/SumMe-shape videos per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  `value`: videos/s with the features already resident in HBM (CUDA events,
max over ranks).  `e2e`: the same job through ScoringPipeline.run from pinned HOST buffers, H2D of the features and
D2H of the kept proposals inside the timed region.  `roofline`: the dominant kernel (to_qkv tcgen05 GEMM) timed
with CUDA events on its own stream during extra instrumented steps.  `cpu_baseline`: the CPU restatement of the
reference (oracle/, torch-CPU + NumPy, all host cores) on a bounded sample of the same videos.
`--impl reference` times that CPU path alone and prints the same line shape.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VIDEOS = 4096
T_LO, T_HI = 100, 800
SEED = 12345                     # reference default seed (src/helpers/init_helper.py:49)
FC_DEPTH = 5
NMS_THRESH = 0.5


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_lengths(rank: int, n_videos: int):
    rng = np.random.default_rng(SEED + rank)
    return [int(t) for t in rng.integers(T_LO, T_HI + 1, size=n_videos)]


def synth_features_device(total_rows: int, device, seed: int):
    """relu(randn) rows, L2-normalised: stand-in for GoogLeNet pool5 features (src/helpers/video_helper.py:72)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((total_rows, 1024), dtype=torch.float32, device=device)
    step = 1 << 16
    for r0 in range(0, total_rows, step):
        r1 = min(total_rows, r0 + step)
        x = torch.relu(torch.randn((r1 - r0, 1024), generator=g, device=device))
        out[r0:r1] = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return out


def xavier_state(scales, seed=SEED):
    """Random-init weights in the reference's training start state (xavier_init, anchor_based/train.py:19-24)."""
    from edsnet_b200 import DSNet
    torch.manual_seed(seed)
    m = DSNet("nystromformer", 1024, 128, list(scales), 8, fc_depth=FC_DEPTH, pooling_type="roi")

    def xavier_init(module):
        name = module.__class__.__name__
        if "Linear" in name or "Conv" in name:
            torch.nn.init.xavier_uniform_(module.weight, gain=np.sqrt(2.0))
            if module.bias is not None:
                torch.nn.init.constant_(module.bias, 0.1)
    m.apply(xavier_init)
    return m


class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int, period=0.02):
        super().__init__(daemon=True)
        self.period = period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:      # pragma: no cover
            log("clock sampling unavailable:", e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_ev.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()
        if self.ok:
            self.join(timeout=2)
        return {"sm_mhz": int(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_setup(state_dict, scales):
    from oracle import dsnet_oracle as orc
    p = {k: v.detach().cpu().float().contiguous() for k, v in state_dict.items() if not k.startswith("fc.")}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    def score(x_cpu):
        return orc.proposals(x_cpu, p, scales, FC_DEPTH, NMS_THRESH)
    return score, cores


def run_cpu_sample(score, x_host, lengths, budget_s, max_videos):
    """Score videos one by one (the reference's loop, evaluate.py:19-28) until the time budget is used."""
    cu = np.concatenate([[0], np.cumsum(lengths)])
    t0 = time.perf_counter()
    n = frames = 0
    while n < min(max_videos, len(lengths)):
        score(x_host[cu[n]:cu[n + 1]])
        frames += lengths[n]
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n, frames, dt


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, warnings from
    libraries that print to fd 1) is redirected to stderr for the whole run."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos", type=int, default=N_VIDEOS, help="videos per GPU per step")
    ap.add_argument("--scales", type=int, nargs="+", default=[12])
    ap.add_argument("--precision", default="fp16x3", choices=["fp32", "fp16x3", "fp16"])
    ap.add_argument("--chunk-rows", type=int, default=32768, help="rows per chunk of the host->device pipeline (e2e)")
    ap.add_argument("--device-chunk-rows", type=int, default=1048576, help="rows per launch sequence, device-resident arm")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    scales = list(args.scales)
    S = len(scales)
    workload = (f"C2: {args.videos} synthetic TVSum/SumMe-shape videos per GPU, T~U[{T_LO},{T_HI}], 1024-d fp32 pool5-like "
                f"features, nystromformer+roi, anchor_scales {scales}, fc_depth {FC_DEPTH}, forward+decode+NMS({NMS_THRESH})")

    # ------------------------------------------------------------------ reference arm: CPU path alone
    if args.impl == "reference":
        if rank != 0:
            return
        lengths = workload_lengths(0, args.videos)
        sample = 32
        model = xavier_state(scales)
        from oracle import dsnet_oracle as orc
        xs = torch.cat([orc.synth_features(t, SEED + 1 + i) for i, t in enumerate(lengths[:sample])])
        score, cores = cpu_reference_setup(model.state_dict(), scales)
        for _ in range(args.warmup):
            run_cpu_sample(score, xs, lengths[:sample], 1e9, sample)
        t0 = time.perf_counter()
        vids = frames = 0
        for _ in range(args.steps):
            n, f, _ = run_cpu_sample(score, xs, lengths[:sample], 1e9, sample)
            vids += n
            frames += f
        dt = time.perf_counter() - t0
        v = vids / dt
        desc = f"first {sample} videos of the workload per step ({sum(lengths[:sample])} frames), one video per call"
        emit(({
            "impl": "reference", "metric": "videos_per_sec", "value": v, "unit": "videos/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "frames_per_sec": frames / dt, "config": {"workload": workload, "sample": desc},
            "cpu_baseline": {"value": v, "unit": "videos/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ B200 arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from edsnet_b200 import BatchPlan, ScoringPipeline, _capi
    from edsnet_b200.pipeline import bind_host_to_gpu_numa_node
    lib = _capi.lib()
    # host threads and the pinned staging buffers of this rank on the socket its GPU hangs off
    orig_affinity = os.sched_getaffinity(0)
    numa_cpus = bind_host_to_gpu_numa_node(local_rank)
    log(f"[rank {rank}] bound to {len(numa_cpus)} GPU-local cores" if numa_cpus else f"[rank {rank}] no NUMA binding")

    lengths = workload_lengths(rank, args.videos)
    R = int(sum(lengths))
    model = xavier_state(scales).to(dev).eval()
    model.precision = args.precision
    log(f"[rank {rank}] {args.videos} videos, {R} frames, x = {R * 4096 / 2**30:.2f} GiB")
    x_dev = synth_features_device(R, dev, SEED + 1000 + rank)
    pipe = ScoringPipeline(model, chunk_rows=args.chunk_rows, nms_thresh=NMS_THRESH)
    chunks = pipe.chunk_videos(lengths, args.device_chunk_rows)
    cu = np.concatenate([[0], np.cumsum(lengths)])
    dplans = [BatchPlan.build(lengths[a:b]).to(dev) for a, b in chunks]
    launches_per_step = len(chunks) * (model.launches_per_forward() + 2)

    def step_device():
        out = None
        with torch.no_grad():
            for (a, b), dp in zip(chunks, dplans):
                xd = x_dev[int(cu[a]):int(cu[b])]
                cls, loc = model._forward_nograd(xd, dp)
                out = model.nms_packed(cls, loc, dp, NMS_THRESH)
        return out

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_device()
    sampler.start()
    ms_dev = timed(step_device, args.steps)
    assert lib.edsnet_debug_tc_status(0) == 0, "tcgen05 pipeline wait timed out during the run"

    # e2e: host pinned buffers -> proposals on the host
    x_host = torch.empty((R, 1024), dtype=torch.float32).pin_memory()
    x_host.copy_(x_dev)
    torch.cuda.synchronize(dev)
    for _ in range(args.warmup):
        res = pipe.run(x_host, lengths, dev)
    ms_e2e = timed(lambda: pipe.run(x_host, lengths, dev), args.steps)
    clocks = sampler.stop()
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    if world > 1:                      # bytes per step of the whole job, like `value`
        t = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        h2d, d2h = int(t[0].item()), int(t[1].item())
    kept_total = int(res[0].sum())

    # instrumented steps: per-stage CUDA events on the launching stream (not part of `value`)
    _capi.check(lib.edsnet_debug_stage_timing(1))
    inst_steps = 2
    for _ in range(inst_steps):
        step_device()
    torch.cuda.synchronize(dev)
    stages = _capi.stage_times()
    _capi.check(lib.edsnet_debug_stage_timing(0))
    total_stage_ms = sum(v[0] for v in stages.values())
    qkv_ms, qkv_n = stages["to_qkv_gemm"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    flops_per_launch = 2.0 * (R / len(chunks)) * 1024 * 1536          # algorithmic: one pass, real rows
    avg_ms = qkv_ms / max(qkv_n, 1)
    achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
    # DRAM traffic of this kernel from the committed `ncu --set full` capture (profiles/r01m_kernels_full.md:
    # dram__bytes_read 0.967633 GB + dram__bytes_write 1.392453 GB for a 229 489-row launch), scaled to this launch's
    # rows; algorithmic bytes are 4 KB (x planes in) + 6 KB (q|k|v planes out) + 96 B (scales) per row.
    rows_per_launch = R / len(chunks)
    traffic = (0.967633e9 + 1.392453e9) * rows_per_launch / 229489.0
    roofline = {"kernel": "gemm_tc_kernel<BN128,BK64,3 stages,3 passes,QKV_PLANES> (to_qkv, fp16 hi/lo split = 3 tcgen05 "
                          "passes, epilogue writes the q|k|v operand planes)", "bound": "tensor",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": peak_src, "traffic": traffic,
                "traffic_source": "ncu --set full, profiles/r01m_kernels_full.md (bytes per row x rows of this launch)",
                "algorithmic_bytes_per_launch": rows_per_launch * (4096 + 6144 + 96),
                "mma_passes": 3, "frac_counting_passes": 3 * achieved / peak_tf,
                "algorithmic_flops_per_launch": flops_per_launch, "avg_launch_ms": avg_ms,
                "share_of_step": qkv_ms / total_stage_ms if total_stage_ms else None,
                "stage_ms_per_step": {k: round(v[0] / inst_steps, 4) for k, v in stages.items()}}

    # per-stage roofline fractions from the same instrumented steps: ALGORITHMIC bytes / flops per feature row (SURVEY
    # 8d; split-precision passes counted once) x rows per step / stage time, against the measured HBM / sustained bf16 peaks
    peak_hbm = float(peaks.get("hbm_gbs", 6500.0))
    per_row = {                                   # stage -> ("hbm", bytes per row) or ("tensor", flops per row)
        "split_f16": ("hbm", 4096 + 4096 + 4), "to_qkv_gemm": ("tensor", 2.0 * 1024 * 1536),
        "landmarks": ("hbm", 4096), "value_conv": ("hbm", 2048 + 2048 + 2048 + 4),
        "to_out_gemm": ("tensor", 2.0 * 512 * 1024), "layernorm1024": ("hbm", 4096 + 4096 + 4),
        "fc1_gemm": ("tensor", 2.0 * 1024 * 128), "fc_stack": ("tensor", FC_DEPTH * 2.0 * 128 * 128),
        # ROI pooling + heads: SURVEY's figure is 4*128 B read + 12 S B written per row; in the tcgen05 modes the three head
        # projections are emitted by the fc stack's last layer (16 B per row), so the hidden rows are never written or
        # re-read and this stage only moves 16 + 12 S bytes per row (both figures are reported)
        "roi_pool_heads": ("hbm", 16 + 12 * S), "decode_boxes": ("hbm", 8 * S + 16 * S), "nms": ("hbm", 25 * S),
    }
    stage_roof = {}
    for name, (kind, amount) in per_row.items():
        if name in stages and stages[name][0] > 0:
            sec = stages[name][0] / inst_steps * 1e-3
            if kind == "hbm":
                ach = amount * R / sec / 1e9
                stage_roof[name] = {"bound": "hbm", "achieved_GBps": round(ach, 1), "frac": round(ach / peak_hbm, 3)}
            else:
                ach = amount * R / sec / 1e12
                stage_roof[name] = {"bound": "tensor", "achieved_TFLOPs": round(ach, 1), "frac": round(ach / peak_tf, 3)}
    if "roi_pool_heads" in stage_roof:
        sec = stages["roi_pool_heads"][0] / inst_steps * 1e-3
        stage_roof["roi_pool_heads"]["vs_unfused_algorithmic_GBps"] = round((4 * 128 + 12 * S) * R / sec / 1e9, 1)
        stage_roof["roi_pool_heads"]["note"] = "head projections fused into the fc stack: 16 B per row in instead of 512"
    roofline["stages"] = stage_roof
    roofline["stages_note"] = ("algorithmic work counted once; the tensor stages run 3 split-fp16 MMA passes (fp32-grade "
                               "accuracy), so their ceiling is 1/3; the attention-core stages (landmark softmaxes, "
                               "pseudo-inverse chain) are latency bound and not listed")

    vids_all = args.videos * world
    frames_all = R * world            # every rank has its own seeded lengths; close enough for the aggregate
    if world > 1:
        t = torch.tensor([float(R)], device=dev)
        dist.all_reduce(t)
        frames_all = int(t.item())
    value = vids_all * args.steps / (ms_dev * 1e-3)
    e2e_v = vids_all * args.steps / (ms_e2e * 1e-3)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, orig_affinity)          # the CPU leg gets every host core again
        score, cores = cpu_reference_setup(model.state_dict(), scales)
        n, frames, dt = run_cpu_sample(score, x_host, lengths, args.cpu_budget, 2048)
        cpu_base = {"value": n / dt, "unit": "videos/s", "cores": cores, "kind": "port",
                    "frames_per_sec": frames / dt,
                    "sample": f"first {n} videos of the workload ({frames} frames), one video per call as "
                              f"evaluate.py:19-28, {dt:.1f} s of CPU work"}
    if rank == 0:
        line = {
            "metric": "videos_per_sec", "value": value, "unit": "videos/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "frames_per_sec": frames_all * args.steps / (ms_dev * 1e-3),
            "config": {"workload": workload, "videos_per_gpu": args.videos, "frames_per_gpu": R,
                       "device_chunk_rows": args.device_chunk_rows, "chunks_per_step": len(chunks),
                       "e2e_chunk_rows": args.chunk_rows, "parallelism": f"video-wise x{world}", "host_numa_binding": bool(numa_cpus),
                       "l2": "inputs (7.5 GB/GPU) exceed L2; no flush needed", "weights": "xavier random init",
                       "kept_proposals": kept_total},
            "e2e": {"value": e2e_v, "unit": "videos/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "frames_per_sec": frames_all * args.steps / (ms_e2e * 1e-3)},
            "gpu_launches": launches_per_step * args.steps,
            "e2e_gpu_launches": pipe.kernel_launches * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_base, "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
