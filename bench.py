#!/usr/bin/env python
"""Benchmark of the EDSNet anchor-based scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config c2|c1|c3|c4|c5] [--strong] [--precision fp16x2|fp16x3|fp16|fp32]

Default (BASELINE.json configs[1], "c2"): batched inference (forward + decode + temporal NMS) over 4096 synthetic
TVSum/SumMe-shape videos per GPU.  One JSON line on stdout (rank 0):
  value         videos/s with the features already resident in HBM (CUDA events, max over ranks)
  e2e           the same job through ScoringPipeline.run from pinned HOST buffers, H2D of the features and D2H of the
                kept proposals inside the timed region, plus the copies-only floor of the same chunks
  roofline      the dominant kernel (to_qkv tcgen05 GEMM), timed with CUDA events on its own stream during extra
                instrumented steps, plus per-stage fractions
  modes         device-resident videos/s of every arithmetic mode in the same run and the parity bar each one meets
  other_configs short measurements of c1 / c3 / c4 / c5 in the same run (so that one driver run sees every config)
  cpu_baseline  the CPU restatement of the reference (oracle/, torch-CPU + NumPy, all host cores) on a bounded sample
--strong: configs[1] as written -- the 4096 videos are the TOTAL, partitioned video-wise by plan.shard_videos.
--config c1 / c4 / c5: one video per step (T=320; full-MHA base T=2048; T=16384 x 4 scales); c3: the data-parallel
training step (k videos per GPU per optimiser step, ONE flat NCCL gradient all-reduce, native forward/backward/Adam).
`--impl reference` times the CPU path of the selected config alone and prints the same line shape.
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
# parity bar each arithmetic mode is held to on pred_cls / pred_loc (tests/util.py TOL) and what BASELINE.json asks
MODE_BARS = {"fp32": "1e-5 (FP32 mode bar)", "fp16x3": "1e-5 (FP32 mode bar, on tensor cores)",
             "fp16x2": "5e-4 (inside the 1e-3 tensor-core bar)", "fp16": "3e-3 (does NOT meet the 1e-3 bar; opt-in only)"}
MODE_PASSES = {"fp16x3": 3, "fp16x2": 2, "fp16": 1}
SINGLE = {"c1": ("nystromformer", 320, [12]), "c4": ("attention", 2048, [4, 8, 16, 32]),
          "c5": ("nystromformer", 16384, [4, 8, 16, 32])}


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


def xavier_state(scales, seed=SEED, base="nystromformer"):
    """Random-init weights in the reference's training start state (xavier_init, anchor_based/train.py:19-24)."""
    from edsnet_b200 import DSNet
    torch.manual_seed(seed)
    m = DSNet(base, 1024, 128, list(scales), 8, fc_depth=FC_DEPTH, pooling_type="roi")

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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def cpu_reference_setup(state_dict, scales, base="nystromformer"):
    from oracle import dsnet_oracle as orc
    p = {k: v.detach().cpu().float().contiguous() for k, v in state_dict.items() if not k.startswith("fc.")}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    def score(x_cpu):
        T = x_cpu.shape[0]
        with torch.no_grad():
            cls, loc = orc.dsnet_forward(x_cpu, p, scales, FC_DEPTH, base=base)
        boxes = orc.clip_round(orc.decode_boxes(loc.numpy(), T, scales), T)
        return orc.nms_1d(cls.numpy().reshape(-1), boxes, NMS_THRESH)
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


class Dist:
    """Rank bookkeeping + the three collectives the bench itself needs (barrier, max, sum)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dev = None

    def init(self):
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def reduce(self, values, op="sum"):
        if self.world == 1:
            return [float(v) for v in values]
        import torch.distributed as dist
        t = torch.tensor([float(v) for v in values], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def timed(self, fn, steps):
        """EXACTLY `steps` calls between two CUDA events, barrier + synchronize on both sides, max over ranks (ms)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.reduce([e0.elapsed_time(e1)], "max")[0]

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def median_ms(fn, n=13, skip=3):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[skip:]))


def dense_flops(T, S, depth=FC_DEPTH):
    """algorithmic dense FLOPs of one Nystrom-base forward (split-precision passes counted once)."""
    n = ((T + 63) // 64) * 64
    return (2.0 * n * 1024 * 1536 + 2.0 * n * 512 * 1024 + 4.0 * 8 * n * 64 * 64 + 2.0 * 8 * 64 ** 3 +
            6 * 4 * 2.0 * 64 ** 3 * 8 + 3 * 2.0 * 8 * n * 64 * 64 + 2.0 * T * 1024 * 128 + depth * 2.0 * T * 128 * 128 +
            2.0 * T * S * 128 * 3)


# ---------------------------------------------------------------------------------------------------------------------
# single-video configurations (c1, c4, c5): latency of forward + decode + NMS, resident and from host buffers
# ---------------------------------------------------------------------------------------------------------------------
def measure_single(cfg_name, dev, precision, steps=20, warmup=3, with_e2e=True, with_stages=True):
    from edsnet_b200 import BatchPlan, _capi
    base, T, scales = SINGLE[cfg_name]
    S = len(scales)
    lib = _capi.lib()
    model = xavier_state(scales, base=base).to(dev).eval()
    model.precision = precision
    x = synth_features_device(T, dev, 1)
    db = BatchPlan.build([T]).to(dev)
    out = {}

    def step():
        with torch.no_grad():
            cls, loc = model._forward_nograd(x, db)
            return model.nms_packed(cls, loc, db, NMS_THRESH)

    for _ in range(warmup):
        r = step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = step()
    e1.record()
    torch.cuda.synchronize(dev)
    out["ms_per_step"] = e0.elapsed_time(e1) / steps
    out["kept"] = int(r["keep_count"].cpu()[0])
    with torch.no_grad():
        out["forward_ms"] = median_ms(lambda: model._forward_nograd(x, db))
        cls, loc = model._forward_nograd(x, db)
        out["decode_nms_ms"] = median_ms(lambda: model.nms_packed(cls, loc, db, NMS_THRESH))
        run = model.graphed_forward([T])
        out["forward_graph_ms"] = median_ms(lambda: run.graph.replay())
    if with_e2e:
        xh = x.cpu().pin_memory()
        xd = torch.empty_like(x)

        def e2e_step():
            xd.copy_(xh, non_blocking=True)
            with torch.no_grad():
                c, l = model._forward_nograd(xd, db)
                rr = model.nms_packed(c, l, db, NMS_THRESH)
            k = int(rr["keep_count"].cpu()[0])          # the host needs the count before it can read the proposals
            return rr["keep_scores"][:k].cpu(), rr["keep_boxes"][:k].cpu()

        for _ in range(warmup):
            ks, kb = e2e_step()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            ks, kb = e2e_step()
        e1.record()
        torch.cuda.synchronize(dev)
        out["e2e_ms_per_step"] = e0.elapsed_time(e1) / steps
        out["h2d_bytes_per_step"] = T * 4096
        out["d2h_bytes_per_step"] = 4 + int(ks.numel()) * 4 + int(kb.numel()) * 4
    if with_stages:
        _capi.check(lib.edsnet_debug_stage_timing(1))
        n_inst = 10
        for _ in range(n_inst):
            step()
        torch.cuda.synchronize(dev)
        stages = _capi.stage_times()
        _capi.check(lib.edsnet_debug_stage_timing(0))
        out["stage_ms"] = {k: round(v[0] / n_inst, 5) for k, v in stages.items() if v[1]}
    assert lib.edsnet_debug_tc_status(1) == 0, "tcgen05 pipeline wait timed out"
    out["launches_per_step"] = model.launches_per_forward() + 2
    out["T"], out["scales"], out["base"] = T, scales, base
    return out, model, x


def single_roofline(cfg_name, res, precision, peaks):
    """Dominant tensor kernel of a single-video config: the q|k|v projection GEMM (by FLOPs and by time)."""
    base, T, scales = SINGLE[cfg_name]
    peak_tf = float(peaks.get("bf16_tflops", 1687.0))      # a kernel timed alone: the burst figure
    qkv_ms = res.get("stage_ms", {}).get("to_qkv_gemm")
    ncols = 3072 if base == "attention" else 1536
    flops = 2.0 * T * 1024 * ncols
    roof = {"kernel": f"gemm_tc_kernel ({'Q|K|V' if base == 'attention' else 'to_qkv'} projection, {T} x 1024 x {ncols})",
            "bound": "tensor", "peak": peak_tf, "unit": "TFLOP/s", "traffic": None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks else "fallback 1687",
            "algorithmic_flops_per_launch": flops,
            "mma_passes": 3 if base == "attention" and precision == "fp16x2" else MODE_PASSES.get(precision, 0)}
    if qkv_ms:
        ach = flops / (qkv_ms * 1e-3) / 1e12
        roof.update(achieved=ach, frac=ach / peak_tf, avg_launch_ms=qkv_ms,
                    share_of_step=qkv_ms / max(sum(res["stage_ms"].values()), 1e-9))
    if base != "attention":
        whole = dense_flops(T, len(scales)) / (res["forward_ms"] * 1e-3) / 1e12
        roof["whole_forward"] = {"algorithmic_gflop": dense_flops(T, len(scales)) / 1e9, "achieved_TFLOPs": whole,
                                 "frac": whole / peak_tf}
    roof["stage_ms_per_step"] = res.get("stage_ms")
    return roof


def cpu_single(cfg_name, budget_s=15.0):
    from oracle import dsnet_oracle as orc
    base, T, scales = SINGLE[cfg_name]
    model = xavier_state(scales, base=base)
    score, cores = cpu_reference_setup(model.state_dict(), scales, base)
    x = orc.synth_features(T, 1)
    score(x)
    t0 = time.perf_counter()
    n = 0
    while True:
        score(x)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 200:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "videos/s", "cores": cores, "kind": "port", "ms_per_video": 1e3 * dt / n,
            "sample": f"{n} calls of the CPU restatement on the same video (T={T}, scales {scales}), forward + decode + NMS, "
                      f"{dt:.1f} s of CPU work"}


def bench_single(args, D):
    cfg = args.config
    base, T, scales = SINGLE[cfg]
    workload = (f"{cfg.upper()}: one synthetic video per step, T={T}, 1024-d fp32 pool5-like features, "
                f"{'full-MHA' if base == 'attention' else 'nystromformer'}+roi, anchor_scales {scales}, fc_depth {FC_DEPTH}, "
                f"forward+decode+NMS({NMS_THRESH})")
    if args.impl == "reference":
        if D.rank != 0:
            return
        cb = cpu_single(cfg, budget_s=max(5.0, 2.0 * args.steps))
        emit({"impl": "reference", "metric": "videos_per_sec", "value": cb["value"], "unit": "videos/s", "n_gpus": args.gpus,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_video"], "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
              "frames_per_sec": cb["value"] * T, "config": {"workload": workload}, "cpu_baseline": cb,
              "e2e": {"value": cb["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return
    D.init()
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    steps = max(args.steps, 10)
    res, _, _ = measure_single(cfg, D.dev, args.precision, steps=steps, warmup=args.warmup)
    clocks = sampler.stop()
    ms = D.reduce([res["ms_per_step"]], "max")[0]
    ms_e2e = D.reduce([res["e2e_ms_per_step"]], "max")[0]
    peaks = load_peaks()
    cb = cpu_single(cfg) if (D.rank == 0 and D.world == 1 and not args.no_cpu_baseline) else None
    if D.rank == 0:
        emit({"metric": "videos_per_sec", "value": D.world * 1e3 / ms, "unit": "videos/s", "n_gpus": D.world,
              "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
              "frames_per_sec": D.world * T * 1e3 / ms,
              "config": {"workload": workload, "parallelism": f"replicas x{D.world} (a single video is not split)",
                         "l2": "single video: inputs are L2 resident by nature of the config (latency measurement)",
                         "weights": "xavier random init", "kept_proposals": res["kept"],
                         "forward_ms": res["forward_ms"], "forward_graph_ms": res["forward_graph_ms"],
                         "decode_nms_ms": res["decode_nms_ms"], "parity_bar": MODE_BARS[args.precision]},
              "e2e": {"value": D.world * 1e3 / ms_e2e, "unit": "videos/s", "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": res["h2d_bytes_per_step"] * D.world,
                      "d2h_bytes_per_step": res["d2h_bytes_per_step"] * D.world},
              "gpu_launches": res["launches_per_step"] * steps,
              "roofline": single_roofline(cfg, res, args.precision, peaks), "cpu_baseline": cb, "clocks": clocks})
    D.close()


# ---------------------------------------------------------------------------------------------------------------------
# c3: data-parallel training step
# ---------------------------------------------------------------------------------------------------------------------
def synth_split(n_videos, rng):
    """tvsum.yml-shaped fold: T ~ U[100, 800], ground-truth keyshot masks of 2-6 segments covering about 15 % of the
    positions (what get_keyshot_summ + downsample_summ produce, anchor_based/train.py:79-86)."""
    vids = []
    for _ in range(n_videos):
        T = int(rng.integers(T_LO, T_HI + 1))
        mask = np.zeros(T, bool)
        budget = int(0.15 * T)
        for _ in range(int(rng.integers(2, 7))):
            ln = max(2, int(budget / 4 * rng.uniform(0.5, 1.5)))
            a = int(rng.integers(0, max(1, T - ln)))
            mask[a:a + ln] = True
        vids.append((T, mask))
    return vids


def measure_c3(D, scales, videos_per_rank, steps, warmup, precision="fp16x3"):
    """k videos per GPU per optimiser step: native train-mode forward (Dropout active) + losses + native backward,
    ONE flat NCCL all-reduce of the 2 248 843-value gradient, native Adam.  Labels come from the per-video cache
    (training.LabelCache: the random negatives are drawn per step on the host, as anchor_based/train.py:91-108)."""
    from edsnet_b200 import BatchPlan, training as tr
    dev = D.dev
    model = xavier_state(scales).to(dev)
    model.precision = precision
    rng = np.random.default_rng(SEED + D.rank)
    vids = synth_split(40, rng)
    feats = [synth_features_device(T, dev, 7000 + 100 * D.rank + i) for i, (T, _) in enumerate(vids)]
    cache = tr.LabelCache(scales)
    for i, (_, mask) in enumerate(vids):
        cache.add(i, mask)
    stepper = tr.NativeDataParallelStep(model, world_size=D.world)
    k = videos_per_rank
    host_s = [0.0]

    def one_step(i):
        sel = [(i * k + j) % len(vids) for j in range(k)]
        t0 = time.perf_counter()
        labs = [cache.labels(s, rng) for s in sel]
        host_s[0] += time.perf_counter() - t0
        return stepper.step([feats[s] for s in sel], [c for c, _ in labs], [l for _, l in labs])

    # every video of the split twice before timing: the second sighting of a length tuple captures its CUDA graph
    n_warm = max(warmup, 2 * ((len(vids) + k - 1) // k))
    for i in range(n_warm):
        one_step(i)
    it = [n_warm]

    def fn():
        one_step(it[0])
        it[0] += 1

    def fn_readback():                     # e2e: labels H2D and the loss back on the host every step
        loss_dev = one_step(it[0])
        it[0] += 1
        return float(loss_dev[:, 0].mean())
    host_s[0] = 0.0
    ms = D.timed(fn, steps)
    label_ms = 1e3 * host_s[0] / steps
    ms_e2e = D.timed(fn_readback, steps)
    loss = stepper.last_loss()
    # share of the collective: the same steps with the all-reduce skipped (gradients stay local)
    t_ar = None
    if D.world > 1:
        stepper.skip_allreduce = True
        ms_no = D.timed(fn, steps)
        stepper.skip_allreduce = False
        t_ar = max(0.0, ms - ms_no) / steps
    # where the step's time goes: per-stage CUDA events over a few instrumented steps (not part of the timed region)
    stage_ms = None
    try:
        from edsnet_b200 import _capi
        lib = _capi.lib()
        _capi.check(lib.edsnet_debug_stage_timing(1))
        n_inst = 10
        for _ in range(n_inst):
            fn()
        torch.cuda.synchronize(dev)
        st = _capi.stage_times()
        _capi.check(lib.edsnet_debug_stage_timing(0))
        stage_ms = {k: round(v[0] / n_inst, 5) for k, v in st.items() if v[1]}
    except Exception as e:          # pragma: no cover
        stage_ms = {"error": repr(e)}
    # forward-only latency of the same videos (inference kernels) for the "step <= 3 x forward" comparison
    model.eval()
    with torch.no_grad():
        sel = [j % len(vids) for j in range(k)]
        xx = torch.cat([feats[s] for s in sel])
        db = BatchPlan.build([vids[s][0] for s in sel]).to(dev)
        fwd_ms = median_ms(lambda: model._forward_nograd(xx, db))
    model.train()
    rows = float(np.mean([vids[j % len(vids)][0] for j in range(n_warm * k, (n_warm + steps) * k)]))
    return {"ms_per_step": ms / steps, "videos_per_sec": D.world * k * steps / (ms * 1e-3), "videos_per_rank_per_step": k,
            "e2e_ms_per_step": ms_e2e / steps, "e2e_videos_per_sec": D.world * k * steps / (ms_e2e * 1e-3),
            "cuda_graph_replays": stepper.graph_replays,
            "host_label_ms_per_step": label_ms, "loss": loss,
            "allreduce_ms_per_step": t_ar, "allreduce_bytes_per_step": stepper.n_params * 4 if D.world > 1 else 0,
            "forward_only_ms": fwd_ms, "step_over_forward": (ms / steps) / fwd_ms, "mean_rows_per_video": rows,
            "launches_per_step": stepper.launches_per_step, "n_params": stepper.n_params, "scales": list(scales),
            "stage_ms_per_step": stage_ms}


def cpu_c3(scales, k, budget_s=15.0):
    """The reference's training loop body (anchor_based/train.py:110-128) on the host cores: torch-CPU autograd of the
    restated model in train() mode (Dropout active), the reference's losses, Adam."""
    from oracle import dsnet_oracle as orc
    from edsnet_b200 import training as tr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = xavier_state(scales)
    p = {n: v.detach().clone().requires_grad_(True) for n, v in model.state_dict().items() if not n.startswith("fc.")}
    opt = torch.optim.Adam(list(p.values()), lr=5e-5, weight_decay=1e-5)
    rng = np.random.default_rng(SEED)
    vids = synth_split(8, rng)
    xs = [orc.synth_features(T, 7000 + i) for i, (T, _) in enumerate(vids)]
    labs = [tr.anchor_labels(m, scales, rng) for _, m in vids]

    def step(i):
        opt.zero_grad()
        tot = 0.0
        for j in range(k):
            s = (i * k + j) % len(vids)
            cls, loc = orc.dsnet_forward(xs[s], p, scales, FC_DEPTH, dropout_gen=torch.Generator().manual_seed(i))
            cl = torch.from_numpy(labs[s][0])
            tot = tot + tr.cls_loss(cls, cl) + tr.loc_loss(loc, torch.from_numpy(labs[s][1]).float(), cl)
        (tot / k).backward()
        opt.step()
    step(0)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s and n < 400:
        step(n + 1)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * k / dt, "unit": "videos/s", "cores": cores, "kind": "port", "ms_per_step": 1e3 * dt / n,
            "sample": f"{n} optimiser steps of {k} video(s) (8 synthetic tvsum-shaped videos, labels precomputed), torch-CPU "
                      f"autograd of the restated model + Adam, {dt:.1f} s of CPU work"}


def bench_c3(args, D):
    scales = list(args.scales) if args.scales_given else [4, 8, 16, 32]
    k = args.videos_per_rank
    workload = (f"C3: anchor-based training step (cls + loc loss, Adam) on a synthetic tvsum.yml-shaped split, T~U[{T_LO},{T_HI}], "
                f"anchor_scales {scales}, fc_depth {FC_DEPTH}, {k} video(s) per GPU per optimiser step, data parallel, one "
                f"flat NCCL gradient all-reduce")
    if args.impl == "reference":
        if D.rank != 0:
            return
        cb = cpu_c3(scales, k, budget_s=max(5.0, 2.0 * args.steps))
        emit({"impl": "reference", "metric": "videos_per_sec", "value": cb["value"], "unit": "videos/s", "n_gpus": args.gpus,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
              "config": {"workload": workload}, "cpu_baseline": cb,
              "e2e": {"value": cb["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return
    D.init()
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    steps = max(args.steps, 20)
    prec = args.precision if args.precision != "fp16x2" else "fp16x3"      # gradients are held to 1e-4: full split
    res = measure_c3(D, scales, k, steps, max(args.warmup, 5), prec)
    clocks = sampler.stop()
    peaks = load_peaks()
    cb = cpu_c3(scales, k) if (D.rank == 0 and D.world == 1 and not args.no_cpu_baseline) else None
    if D.rank == 0:
        T = res["mean_rows_per_video"]
        flops = 3.0 * dense_flops(int(T), len(scales)) * k         # forward + backward (dX and dW): 3 x the forward
        peak_tf = float(peaks.get("bf16_tflops", 1687.0))
        ach = flops / (res["ms_per_step"] * 1e-3) / 1e12
        emit({"metric": "videos_per_sec", "value": res["videos_per_sec"], "unit": "videos/s", "n_gpus": D.world, "steps": steps,
              "warmup": max(args.warmup, 5), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": prec, "data": "synthetic",
              "config": {"workload": workload, "parallelism": f"data parallel x{D.world}", "weights": "xavier random init",
                         "l2": "one video per step fits L2 (latency-bound step by nature of the config)", **res},
              # the features of the training split live on the device (as the reference keeps them after .to(device));
              # per step the host sends the labels and reads the loss back
              "e2e": {"value": res["e2e_videos_per_sec"], "unit": "videos/s", "ms_per_step": res["e2e_ms_per_step"],
                      "h2d_bytes_per_step": int(D.world * k * T * len(scales) * 12), "d2h_bytes_per_step": 4 * D.world,
                      "note": "every step uploads its labels from pinned host memory and reads its loss back; the features "
                              "of the training split are resident (the reference keeps them on the device as well)"},
              "gpu_launches": res["launches_per_step"] * steps,
              "roofline": {"kernel": "whole training step (launch / latency bound at one video per GPU)", "bound": "tensor",
                           "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                           "algorithmic_flops_per_launch": flops},
              "cpu_baseline": cb, "clocks": clocks})
    D.close()


# ---------------------------------------------------------------------------------------------------------------------
# c2: batched inference (the headline configuration)
# ---------------------------------------------------------------------------------------------------------------------
def bench_c2(args, D):
    scales = list(args.scales)
    S = len(scales)
    per = "in TOTAL (strong scaling)" if args.strong else "per GPU"
    workload = (f"C2: {args.videos} synthetic TVSum/SumMe-shape videos {per}, T~U[{T_LO},{T_HI}], 1024-d fp32 pool5-like "
                f"features, nystromformer+roi, anchor_scales {scales}, fc_depth {FC_DEPTH}, forward+decode+NMS({NMS_THRESH})")

    # ------------------------------------------------------------------ reference arm: CPU path alone
    if args.impl == "reference":
        if D.rank != 0:
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
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "frames_per_sec": frames / dt, "config": {"workload": workload, "sample": desc},
            "cpu_baseline": {"value": v, "unit": "videos/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ B200 arm
    D.init()
    dev, rank, world = D.dev, D.rank, D.world
    from edsnet_b200 import BatchPlan, ScoringPipeline, _capi, shard_videos
    from edsnet_b200.pipeline import bind_host_to_gpu_numa_node
    lib = _capi.lib()
    # host threads and the pinned staging buffers of this rank on the socket its GPU hangs off
    orig_affinity = os.sched_getaffinity(0)
    numa_cpus = bind_host_to_gpu_numa_node(D.local_rank)
    log(f"[rank {rank}] bound to {len(numa_cpus)} GPU-local cores" if numa_cpus else f"[rank {rank}] no NUMA binding")

    if args.strong:
        all_lengths = workload_lengths(0, args.videos)          # ONE global list, the same on every rank
        mine = shard_videos(all_lengths, world)[rank]           # balanced on padded rows, no collective on the data path
        lengths = [all_lengths[i] for i in mine]
        feat_seed = SEED + 1000
    else:
        lengths = workload_lengths(rank, args.videos)
        feat_seed = SEED + 1000 + rank
    n_local = len(lengths)
    R = int(sum(lengths))
    model = xavier_state(scales).to(dev).eval()
    model.precision = args.precision
    log(f"[rank {rank}] {n_local} videos, {R} frames, x = {R * 4096 / 2**30:.2f} GiB")
    x_dev = synth_features_device(R, dev, feat_seed)
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

    sampler = ClockSampler(D.local_rank)
    for _ in range(args.warmup):
        step_device()
    sampler.start()
    ms_dev = D.timed(step_device, args.steps)
    assert lib.edsnet_debug_tc_status(0) == 0, "tcgen05 pipeline wait timed out during the run"

    # e2e: host pinned buffers -> proposals on the host
    x_host = torch.empty((R, 1024), dtype=torch.float32).pin_memory()
    x_host.copy_(x_dev)
    torch.cuda.synchronize(dev)
    for _ in range(args.warmup):
        res = pipe.run(x_host, lengths, dev)
    ms_e2e = D.timed(lambda: pipe.run(x_host, lengths, dev), args.steps)
    clocks = sampler.stop()
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_launches = pipe.kernel_launches
    kept_total = int(res[0].sum())
    # the floor under e2e: the same chunks' copies with no kernel in between (PCIe link + host memory system)
    floor_steps = max(3, args.steps // 2)
    pipe.run_copies_only(x_host, lengths, dev)
    ms_floor = D.timed(lambda: pipe.run_copies_only(x_host, lengths, dev), floor_steps) / floor_steps
    h2d, d2h = (int(v) for v in D.reduce([h2d, d2h]))            # bytes per step of the whole job, like `value`

    # the other arithmetic modes, device-resident, same run (fewer steps)
    modes = {args.precision: {"value": None, "ms_per_step": ms_dev / args.steps, "parity_bar": MODE_BARS[args.precision],
                              "mma_passes_big_gemms": MODE_PASSES.get(args.precision, 0)}}
    for other_mode in ([] if args.no_modes else [m for m in ("fp16x3", "fp16x2") if m != args.precision]):
        model.precision = other_mode
        for _ in range(2):
            step_device()
        k = max(3, args.steps // 2)
        modes[other_mode] = {"value": None, "ms_per_step": D.timed(step_device, k) / k, "parity_bar": MODE_BARS[other_mode],
                             "mma_passes_big_gemms": MODE_PASSES.get(other_mode, 0)}
    model.precision = args.precision
    step_device()

    # instrumented steps: per-stage CUDA events on the launching stream (not part of `value`)
    _capi.check(lib.edsnet_debug_stage_timing(1))
    inst_steps = 10
    for _ in range(inst_steps):
        step_device()
    torch.cuda.synchronize(dev)
    stages = _capi.stage_times()
    _capi.check(lib.edsnet_debug_stage_timing(0))
    total_stage_ms = sum(v[0] for v in stages.values())
    qkv_ms, qkv_n = stages["to_qkv_gemm"]
    peaks = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    passes = MODE_PASSES.get(args.precision, 0)
    rows_per_launch = R / len(chunks)
    flops_per_launch = 2.0 * rows_per_launch * 1024 * 1536          # algorithmic: one pass, real rows
    avg_ms = qkv_ms / max(qkv_n, 1)
    achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
    # algorithmic bytes per row of this kernel: the x planes it reads (hi, + lo with three passes) and the q|k|v planes
    # + scales it writes.  `traffic` is NOT measured in this run: it is the dram__bytes of the committed ncu --set full
    # capture of the three-pass kernel (profiles/r01m_kernels_full.md, 229 489-row launch) scaled to this launch's rows,
    # and only quoted for that mode.
    alg_bytes_row = (4096 if passes == 3 else 2048) + 6144 + 96
    # per 229 489-row launch: three passes 0.968 GB read + 1.392 GB written (r01m), two passes 0.478 + 1.382 GB (r02s: the
    # x lo plane is neither written nor read)
    cap = {3: (0.967633e9 + 1.392453e9, "profiles/r01m_kernels_full.md"),
           2: (0.478e9 + 1.382e9, "profiles/r02s_kernels_full_fp16x2.md")}.get(passes)
    traffic = cap[0] * rows_per_launch / 229489.0 if cap else None
    roofline = {"kernel": f"gemm_tc_kernel<BN128,BK64,{passes} passes,QKV_PLANES> (to_qkv, fp16 hi/lo split operands, epilogue "
                          "writes the q|k|v operand planes)", "bound": "tensor",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": peak_src, "traffic": traffic,
                "traffic_source": (f"static: ncu --set full capture of a 229 489-row launch of the {passes}-pass kernel "
                                   f"({cap[1]}), bytes per row x rows of this launch") if cap else "null: no capture of this mode",
                "algorithmic_bytes_per_launch": rows_per_launch * alg_bytes_row,
                "mma_passes": passes, "frac_counting_passes": passes * achieved / peak_tf,
                "algorithmic_flops_per_launch": flops_per_launch, "avg_launch_ms": avg_ms, "instrumented_steps": inst_steps,
                "share_of_step": qkv_ms / total_stage_ms if total_stage_ms else None,
                "stage_sum_ms_per_step": total_stage_ms / inst_steps,
                "stage_ms_per_step": {k: round(v[0] / inst_steps, 4) for k, v in stages.items()}}

    # per-stage roofline fractions from the same instrumented steps: ALGORITHMIC bytes / flops per feature row (SURVEY
    # 8d; split-precision passes counted once) x rows per step / stage time, against the measured HBM / sustained bf16 peaks
    peak_hbm = float(peaks.get("hbm_gbs", 6500.0))
    per_row = {                                   # stage -> ("hbm", bytes per row) or ("tensor", flops per row)
        "split_f16": ("hbm", 4096 + (2048 if passes == 2 else 4096) + 4), "to_qkv_gemm": ("tensor", 2.0 * 1024 * 1536),
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
    roofline["stages_note"] = ("algorithmic work counted once; to_qkv / to_out run `mma_passes` split-fp16 MMA passes, fc1 and "
                               "the fc block always 3, so their ceilings are 1/passes; the attention-core stages (landmark "
                               "softmaxes, pseudo-inverse chain) are latency bound and not listed")

    vids_all, frames_all = (int(v) for v in D.reduce([n_local, R]))
    value = vids_all * args.steps / (ms_dev * 1e-3)
    e2e_v = vids_all * args.steps / (ms_e2e * 1e-3)
    for m in modes.values():
        m["value"] = vids_all / (m["ms_per_step"] * 1e-3)

    other = {}
    if not args.no_other_configs:
        del x_dev
        torch.cuda.empty_cache()
        if world == 1:
            for name in ("c1", "c4", "c5"):
                try:
                    r, _, _ = measure_single(name, dev, args.precision, steps=10, warmup=3, with_e2e=False, with_stages=False)
                    other[name] = {k: r[k] for k in ("T", "scales", "base", "ms_per_step", "forward_ms", "forward_graph_ms",
                                                     "decode_nms_ms", "kept")}
                except Exception as e:            # a secondary measurement must not take the headline line down
                    other[name] = {"error": repr(e)}
        try:                                      # every rank takes part: the step holds a collective
            other["c3_train_step"] = measure_c3(D, [4, 8, 16, 32], 1, 20, 5)
        except Exception as e:
            other["c3_train_step"] = {"error": repr(e)}

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
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "frames_per_sec": frames_all * args.steps / (ms_dev * 1e-3),
            "config": {"workload": workload, "videos_total": vids_all, "videos_rank0": n_local, "frames_rank0": R,
                       "device_chunk_rows": args.device_chunk_rows, "chunks_per_step": len(chunks),
                       "e2e_chunk_rows": args.chunk_rows, "parallelism": f"video-wise x{world}", "host_numa_binding": bool(numa_cpus),
                       "l2": "inputs (GBs per GPU) exceed L2; no flush needed", "weights": "xavier random init",
                       "kept_proposals": kept_total, "parity_bar": MODE_BARS[args.precision]},
            "e2e": {"value": e2e_v, "unit": "videos/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "frames_per_sec": frames_all * args.steps / (ms_e2e * 1e-3),
                    "copies_only_ms_per_step": ms_floor, "frac_of_copy_floor": ms_floor / (ms_e2e / args.steps),
                    "h2d_GBps_per_gpu_in_e2e": (h2d / world) / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "h2d_GBps_per_gpu_copies_only": (h2d / world) / (ms_floor * 1e-3) / 1e9},
            "gpu_launches": launches_per_step * args.steps,
            "e2e_gpu_launches": e2e_launches * args.steps,
            "modes": modes, "other_configs": other or None,
            "roofline": roofline, "cpu_baseline": cpu_base, "clocks": clocks,
        }
        emit(line)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c1", "c3", "c4", "c5"])
    ap.add_argument("--strong", action="store_true", help="c2: --videos is the TOTAL, sharded by plan.shard_videos")
    ap.add_argument("--videos", type=int, default=N_VIDEOS, help="videos per GPU per step (total with --strong)")
    ap.add_argument("--videos-per-rank", type=int, default=1, help="c3: videos per GPU per optimiser step")
    ap.add_argument("--scales", type=int, nargs="+", default=None)
    ap.add_argument("--precision", default="fp16x2", choices=["fp32", "fp16x3", "fp16", "fp16x2"])
    ap.add_argument("--chunk-rows", type=int, default=32768, help="rows per chunk of the host->device pipeline (e2e)")
    ap.add_argument("--device-chunk-rows", type=int, default=1048576, help="rows per launch sequence, device-resident arm")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3
    args.scales_given = args.scales is not None
    if args.scales is None:
        args.scales = [12]
    D = Dist()
    if args.config == "c2":
        bench_c2(args, D)
    elif args.config == "c3":
        bench_c3(args, D)
    else:
        bench_single(args, D)


if __name__ == "__main__":
    main()
