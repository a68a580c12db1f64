#!/usr/bin/env python
"""bench.py -- BiCodec detokenize throughput (audio-seconds decoded per wall-second) on B200.

    python bench.py --gpus N --steps K --warmup W          # ours (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N ...          # reference CPU detokenize, rank 0 only

One "step" = one detokenize pass over one batch of synthetic token streams with the synthetic
(random-init) BiCodec checkpoint.  Headline workload (every N): BASELINE config 2 = batch 64 x 10 s per GPU
(500 semantic + 32 global tokens per utterance), fp32 parity mode; the batch is sharded by utterance,
no data-path collective ("scaling": "weak").  Prints ONE JSON line (rank 0).

The same line carries sub-records for the other BASELINE.json configs so that the driver's run sees them:
  "bf16"  config 2 in bf16 mode (value, roofline of its dominant kernel)
  "c1"    config 1's shape on the GPU: one 10 s utterance, batch 1, p50 / p95 latency per call, fp32 and bf16 (N = 1 only)
  "c3"    batch 1024 x 30 s utterance-sharded = 128 utterances per GPU at every N
  "c4"    256 concurrent streams x 50-token chunks: p50/p99 chunk latency, fp32 and bf16 (N = 1 only)
  "c5"    batch 32 x 120 s time-sharded over the N GPUs (N = 1: the same 8-window schedule run by one process),
          halo-exchange and recompute variants
(`--sub none` skips them, `--sub bf16,c4` selects.)
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GFLOP_PER_FRAME = 1.175364          # algorithmic 2*MAC per token frame
GFLOP_PER_UTT = 0.0289
FRAME_RATE = 50.0
METRIC = "audio-sec decoded/sec (BiCodec detokenize)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def _ncu_traffic(kernel_name: str, precision: str, batch: int, frames: int):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernel from the committed
    `ncu --set full` capture of this command: profiles/r2_traffic.json (else r1_traffic.json), a table
    {"<precision> b<batch> t<frames>": {"<kernel name>": bytes}} written by tools/ncu_traffic.py.  DRAM counters
    cannot be read outside a profiler, so this is the capture's number, not one measured in this run."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            v = d[f"{precision} b{batch} t{frames}"].get(kernel_name)
            if v is not None:
                return v, f"profiles/{name} (ncu --set full capture of this command)"
        except (OSError, KeyError, ValueError):
            continue
    return None, None


class ClockSampler:
    """Samples nvidia-smi SM clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for l in f:
                if l.startswith("model name"):
                    return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return ""


def cpu_reference(cfg, sd, steps: int, warmup: int, batch: int, frames: int):
    """The reference's CPU detokenize (oracle port: same ATen/oneDNN op sequence as
    sparktts/models/bicodec.py:171-189; pinned bit-exact against the reference modules at B=1,
    oracle/validate_against_reference.py) on all host cores, on `batch` utterances of the step's shape.
    Returns (audio_s_per_s, info)."""
    import torch
    from oracle import bicodec_oracle as O                       # cpu_baseline / reference arm only
    from spark_tts_b200.synthetic import synthetic_tokens

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sem, glob = synthetic_tokens(cfg, batch, frames, seed=4321)
    for _ in range(warmup):
        O.detokenize(sd, cfg, sem, glob)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.detokenize(sd, cfg, sem, glob)
        ts.append(time.perf_counter() - t0)
    mean = sum(ts) / len(ts)
    audio_s = batch * frames / FRAME_RATE
    return audio_s / mean, dict(cores=cores, cpu=_cpu_model(), ms=mean * 1e3,
                                sample=f"{batch} of the step's utterances x {frames / FRAME_RATE:.0f} s as one CPU "
                                       f"batch, mean of {steps} runs after {warmup} warm-up")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=500, help="token frames per utterance (50 Hz)")
    ap.add_argument("--sub", default="all", help="sub-records: all | none | comma list of bf16,c1,c3,c4,c5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    subs = ({"bf16", "c1", "c3", "c4", "c5"} if args.sub == "all"
            else set(x for x in args.sub.split(",") if x and x != "none"))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner under
    # torchrun, ...) is diverted to stderr, the result line goes to the saved descriptor.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(result_fd, (json.dumps(obj) + "\n").encode())

    import torch
    from spark_tts_b200 import BiCodecConfig
    from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

    cfg = BiCodecConfig()
    workload = f"BiCodec detokenize batch {args.batch} x {args.frames / FRAME_RATE:.0f} s synthetic tokens per GPU"
    # identical in both arms (the reference arm times a bounded sample of it, described in cpu_baseline.sample)
    config = {"workload": workload, "utterances_per_gpu": args.batch, "frames": args.frames,
              "global_batch": args.batch * max(world, 1), "sharding": "by utterance, no collective",
              "precision_mode": args.precision, "weights": "synthetic random-init BiCodec (seed 0)",
              "l2": "not flushed: per-step activation working set (~16 GB) >> 126 MB L2"}
    warmup = max(args.warmup, 3)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sd = synthetic_state_dict(cfg, 0)
        # bounded sample of the step: as many of its utterances as keep the whole run near two minutes
        # (~0.7 s per 10 s utterance on 16 cores); the utterances go through the CPU path as ONE batch
        per_utt = 0.7 * args.frames / 500.0
        sample = 1
        while sample < min(args.batch, 8) and 2 * sample * per_utt * (args.steps + warmup) <= 120.0:
            sample *= 2
        v, info = cpu_reference(cfg, sd, args.steps, warmup, sample, args.frames)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": info["ms"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": info["cores"], "kind": "port",
                                 "sample": info["sample"], "cpu": info["cpu"]},
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "one CPU process on rank 0's host cores at every N (the reference has no multi-GPU path); "
                        "ms_per_step is the time of the bounded sample, value = its audio-seconds / that time"}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    from spark_tts_b200 import BiCodec, BiCodecTokenizer, sharding, _lib
    from spark_tts_b200.streaming import StreamingDetokenizer

    fp32_terms = int(_lib.load().sparkcodec_fp32_terms())
    FP32_DTYPE = {2: "f32 (fp16 main product + two e5m2 cross products on tcgen05, fp32 accumulate)",
                  3: "f32 (bf16x3 split products on tcgen05, fp32 accumulate)"}
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    sd = synthetic_state_dict(cfg, 0)
    model = BiCodec.from_state_dict(cfg, sd, device=dev, precision=args.precision)
    tok = BiCodecTokenizer(device=dev, model=model)
    B, T = args.batch, args.frames
    sem_h, glob_h = synthetic_tokens(cfg, B, T, seed=1234 + rank)
    glob_h = glob_h.squeeze(1)
    sem_d, glob_d = sem_h.to(dev), glob_h.to(dev).unsqueeze(1)
    sem_p, glob_p = sem_h.pin_memory(), glob_h.pin_memory()
    audio_s_per_step = B * T / FRAME_RATE
    peaks = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warm, sample_clocks=False):
        """W untimed calls, then exactly `steps` calls between barrier + synchronize, CUDA events on the launching
        stream, max over ranks.  -> (total ms, clocks or None)"""
        for _ in range(warm):
            fn()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        return reduce_max(e0.elapsed_time(e1)), clocks

    def profile_rows(m, sem, glob, n_prof=3):
        rows = None
        for _ in range(n_prof):                      # average launch duration over n_prof profiled passes
            m.profile(True)
            m.detokenize(sem, glob)
            cur = m.profile_read()
            m.profile(False)
            if rows is None:
                rows = cur
            else:
                for a, b in zip(rows, cur):
                    a["ms"] += b["ms"]
        for r in rows:
            r["ms"] /= n_prof
        return rows

    def roofline_of(rows, precision):
        """Roofline record of the dominant tensor kernel (by name aggregate) + the HBM-class kernels."""
        gemm = [r for r in rows if r["name"].startswith(("conv_gemm_tc", "resunit_fused", "convnext_fused"))]
        agg = {}
        for r in gemm:
            a = agg.setdefault(r["name"], dict(ms=0.0, flops=0.0, n=0))
            a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["n"] += 1
        total_ms = sum(r["ms"] for r in rows)
        gemm_ms = sum(r["ms"] for r in gemm)
        gemm_flops = sum(r["flops"] for r in gemm)
        top_name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        work = float(fp32_terms) if precision == "fp32" else 1.0   # bf16-MMA equivalents of tensor time per MAC
        ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
        traffic, traffic_src = _ncu_traffic(top_name, precision, B, T)
        roof = {
            "bound": "tensor", "kernel": top_name, "launches": top["n"],
            "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peaks["source"] + ", sustained bf16",
            "tensor_work_factor": work, "tensor_pipe_frac": ach * work / peaks["tf_sustained"],
            "ms_per_launch": top["ms"] / top["n"], "share_of_step": top["ms"] / total_ms,
            "all_gemm": {"achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12,
                         "frac": gemm_flops / (gemm_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                         "tensor_pipe_frac": gemm_flops * work / (gemm_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                         "share_of_step": gemm_ms / total_ms},
            "note": "achieved = algorithmic 2*MAC of the convolution / CUDA-event duration (per-launch events of a "
                    "separate profiled pass, mean of 3); the fp32 mode spends tensor_work_factor bf16-MMA "
                    "equivalents of tensor time per algorithmic MAC (2: one fp16 product + two e5m2 products at twice "
                    "the rate; 3: three bf16 products)"}
        fused = [r for r in rows if r["name"].startswith("resunit_fused")]
        if fused:
            f_ms, f_fl = sum(r["ms"] for r in fused), sum(r["flops"] for r in fused)
            roof["resunit_fused"] = {
                "launches": len(fused), "achieved": f_fl / (f_ms * 1e-3) / 1e12,
                "tensor_pipe_frac": f_fl * work / (f_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                "share_of_step": f_ms / total_ms}
        stream = {}
        sq = re.compile(r"conv_gemm_tc cin=(\d+) n=(\d+) .* taps=1 .* res=1$")
        for r in rows:
            key = r["name"] if r["name"] in ("head", "dwconv_ln", "ln") else None
            m = sq.match(r["name"])
            if m and m.group(1) == m.group(2):       # the square 1x1 residual convs of the wide ResidualUnits only
                key = "conv1x1_residual (conv_gemm_tc, HBM-bound class)"
            if key:
                a = stream.setdefault(key, dict(ms=0.0, bytes=0.0, n=0))
                a["ms"] += r["ms"]; a["bytes"] += r["bytes"]; a["n"] += 1
        hbm = {k: {"launches": v["n"], "achieved_gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                   "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                   "share_of_step": v["ms"] / total_ms} for k, v in stream.items()}
        return roof, hbm

    # ---- device-resident throughput ("value") ----
    model.validate_tokens = False
    launches0 = model.launch_count()
    ms, clocks = timed(lambda: model.detokenize(sem_d, glob_d), args.steps, warmup, sample_clocks=True)
    launches = (model.launch_count() - launches0) * args.steps // (args.steps + warmup)
    model.check_tokens()
    value = world * audio_s_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the reference-surface call, host buffers in and out ("e2e") ----
    # BiCodecTokenizer.detokenize(global_tokens, semantic_tokens) -> numpy, exactly the call cli/SparkTTS.py:231-234
    # makes (which moves its host tensors with .to(device) first); token validation stays on as a user has it.
    model.validate_tokens = True
    out_np = [None]

    def e2e_call():
        out_np[0] = tok.detokenize(glob_p.to(dev, non_blocking=True), sem_p.to(dev, non_blocking=True))

    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()                                   # ends with the D2H of the waveform + a stream sync
    wall_e2e = reduce_max((time.perf_counter() - t0) * 1e3)
    barrier()
    for _ in range(2):
        tok.detokenize_pinned(glob_p, sem_p)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = tok.detokenize_pinned(glob_p, sem_p)
    wall_pinned = reduce_max((time.perf_counter() - t0) * 1e3)
    barrier()
    model.validate_tokens = False
    e2e = {"value": world * audio_s_per_step * args.steps / (wall_e2e * 1e-3), "unit": "audio-s/s",
           "h2d_bytes_per_step": sem_p.numel() * sem_p.element_size() + glob_p.numel() * glob_p.element_size(),
           "d2h_bytes_per_step": int(out_np[0].size) * 4, "ms_per_step": wall_e2e / args.steps,
           "call": "BiCodecTokenizer.detokenize(global_tokens, semantic_tokens) -> numpy (the reference surface), "
                   "pinned host tokens in, host waveform out, token validation on; host wall clock, max over ranks",
           "pinned_variant": {"value": world * audio_s_per_step * args.steps / (wall_pinned * 1e-3),
                              "ms_per_step": wall_pinned / args.steps,
                              "call": "BiCodecTokenizer.detokenize_pinned (reusable pinned output, no validation sync)"}}

    line = {"metric": METRIC, "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": (FP32_DTYPE[fp32_terms] if args.precision == "fp32" else "bf16"),
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches}

    # ---- roofline of the dominant kernel: per-launch CUDA events in a separate, untimed pass ----
    if rank == 0 and not args.no_profile:
        roof, hbm = roofline_of(profile_rows(model, sem_d, glob_d), args.precision)
        line["roofline"], line["hbm_kernels"] = roof, hbm

    # ================================================================== sub-records (other BASELINE configs)
    sub_steps = max(10, args.steps)
    if "bf16" in subs and args.precision == "fp32":
        m16 = BiCodec.from_state_dict(cfg, sd, device=dev, precision="bf16")
        m16.validate_tokens = False
        ms16, clk16 = timed(lambda: m16.detokenize(sem_d, glob_d), sub_steps, warmup, sample_clocks=True)
        rec = {"config": "BASELINE config 2, bf16 mode: batch 64 x 10 s per GPU", "dtype": "bf16", "steps": sub_steps,
               "warmup": warmup, "ms_per_step": ms16 / sub_steps,
               "value": world * audio_s_per_step * sub_steps / (ms16 * 1e-3), "unit": "audio-s/s", "clocks": clk16,
               "tensor_ceiling_audio_s_per_s": world * peaks["tf_sustained"] * 1e12 / (GFLOP_PER_FRAME * 1e9 * FRAME_RATE)}
        if rank == 0 and not args.no_profile:
            rec["roofline"], rec["hbm_kernels"] = roofline_of(profile_rows(m16, sem_d, glob_d), "bf16")
        line["bf16"] = rec
        del m16
        torch.cuda.empty_cache()

    if "c3" in subs:
        B3, T3 = 128, 1500
        s3, g3 = synthetic_tokens(cfg, B3, T3, 3000 + rank)
        s3, g3 = s3.to(dev), g3.to(dev)
        ms3, _ = timed(lambda: model.detokenize(s3, g3), sub_steps, 3)
        line["c3"] = {"config": "BASELINE config 3: batch 1024 x 30 s utterance-sharded = 128 utterances per GPU at "
                                "every N (weak scaling), no data-path collective", "n_gpus": world,
                      "precision_mode": args.precision, "steps": sub_steps, "warmup": 3,
                      "ms_per_step": ms3 / sub_steps,
                      "value": world * B3 * T3 / FRAME_RATE * sub_steps / (ms3 * 1e-3), "unit": "audio-s/s",
                      "passes_per_step": "batch split into passes that fit the 24 GB workspace cap"}
        del s3, g3

    if "c5" in subs:
        B5, T5 = 32, 6000
        s5, g5 = synthetic_tokens(cfg, B5, T5, 5000)            # the same tokens on every rank (replicated)
        s5, g5 = s5.to(dev), g5.to(dev)
        ph, wh = model.halo_frames()
        if world > 1:
            fx = lambda: sharding.detokenize_time_sharded(model, s5, g5, exchange=True)
            fr = lambda: sharding.detokenize_time_sharded(model, s5, g5, exchange=False)
            how = f"each utterance time-sharded over {world} ranks, NCCL halo exchange overlapped with staging the interior rows"
            wx, (a5, b5) = fx()
            wr, _ = fr()
        else:
            fx = lambda: sharding.detokenize_time_windows(model, s5, g5, 8, exchange=True)
            fr = lambda: sharding.detokenize_time_windows(model, s5, g5, 8, exchange=False)
            how = "N = 1: the 8-window time-sharded schedule run by one process (halo rows from the neighbouring window)"
            wx, wr = fx(), fr()
            a5, b5 = 0, T5 // 8
        diff = reduce_max((wx - wr).abs().max().item())
        del wx, wr
        msx, _ = timed(fx, sub_steps, 2)
        msr, _ = timed(fr, sub_steps, 2)
        line["c5"] = {"config": "BASELINE config 5: batch 32 x 120 s, split along time; " + how, "n_gpus": world,
                      "precision_mode": args.precision, "steps": sub_steps, "warmup": 2,
                      "frames_per_shard": b5 - a5, "halo_frames": {"prenet": ph, "wavegen": wh},
                      "halo_bytes_per_neighbour": B5 * wh * cfg.d_model * 4,
                      "exchange": {"ms_per_step": msx / sub_steps,
                                   "value": B5 * T5 / FRAME_RATE * sub_steps / (msx * 1e-3), "unit": "audio-s/s"},
                      "recompute": {"ms_per_step": msr / sub_steps,
                                    "value": B5 * T5 / FRAME_RATE * sub_steps / (msr * 1e-3), "unit": "audio-s/s"},
                      "max_abs_diff_between_variants": diff, "scaling": "strong"}
        del s5, g5
        torch.cuda.empty_cache()

    if "c4" in subs and world == 1:
        S4, T4 = 256, 50
        s4, g4 = synthetic_tokens(cfg, S4, T4, 4000)
        s4d, g4d = s4.to(dev), g4.squeeze(1).to(dev)
        rec4 = {"config": "BASELINE config 4: 256 concurrent streams x 50-token chunks (1 s of audio each), one "
                          "CUDA-graph replay per round", "rounds": 200,
                "timed": "chunk tokens on device -> waveform chunk in pinned host memory (D2H + stream sync "
                         "included), host wall clock per round"}
        for prec in ("fp32", "bf16"):
            m4 = model if prec == args.precision else BiCodec.from_state_dict(cfg, sd, device=dev, precision=prec)
            st = StreamingDetokenizer(m4, use_graphs=True)
            for _ in range(5):
                st.decode_batch(s4d, g4d)
            lat = []
            for _ in range(rec4["rounds"]):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                st.decode_batch(s4d, g4d)
                lat.append((time.perf_counter() - t0) * 1e3)
            lat.sort()
            rec4[prec] = {"p50_ms": lat[len(lat) // 2], "p99_ms": lat[int(len(lat) * 0.99) - 1], "max_ms": lat[-1],
                          "audio_s_per_s_at_p50": S4 * T4 / FRAME_RATE / (lat[len(lat) // 2] * 1e-3)}
            del st
            if m4 is not model:
                del m4
        line["c4"] = rec4

    if "c1" in subs and world == 1:
        # BASELINE config 1's shape (1 utterance x 10 s): the call cli/SparkTTS.py:231-234 makes, tokens on the device
        # -> waveform on the device, one CUDA-event pair per call (tools/latency_single.py is the stand-alone version)
        try:
            rec1 = {"config": "BASELINE config 1's shape: 1 utterance x 10 s (500 semantic + 32 global tokens), batch 1",
                    "calls": 100, "timed": "BiCodec.detokenize, tokens on device -> waveform on device, CUDA events "
                                           "around each call, eager launches (programmatic dependent launch, narrow "
                                           "N tiles for few-tile launches)"}
            s1, g1 = synthetic_tokens(cfg, 1, 500, 1000)
            s1d, g1d = s1.to(dev), g1.to(dev)
            for prec in ("fp32", "bf16"):
                for _ in range(10):
                    model.detokenize(s1d, g1d, precision=prec)
                lat = []
                for _ in range(rec1["calls"]):
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record()
                    model.detokenize(s1d, g1d, precision=prec)
                    ev1.record()
                    ev1.synchronize()
                    lat.append(ev0.elapsed_time(ev1))
                lat.sort()
                p50 = lat[len(lat) // 2]
                rec1[prec] = {"p50_ms": p50, "p95_ms": lat[int(len(lat) * 0.95) - 1],
                              "audio_s_per_s_at_p50": 500 / FRAME_RATE / (p50 * 1e-3)}
            line["c1"] = rec1
        except Exception as exc:                      # a latency side record must never cost the bench line
            line["c1"] = {"error": repr(exc)}

    if rank == 0:
        # ---- CPU baseline beside it (bounded sample of the same step) ----
        if not args.no_cpu_baseline and world == 1:
            v, info = cpu_reference(cfg, sd, steps=3, warmup=1, batch=min(B, 4), frames=T)
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": info["cores"], "kind": "port",
                                    "sample": info["sample"], "cpu": info["cpu"]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
