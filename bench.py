#!/usr/bin/env python
"""bench.py -- BiCodec detokenize throughput (audio-seconds decoded per wall-second) on B200.

    python bench.py --gpus N --steps K --warmup W          # ours (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N ...          # reference CPU detokenize, rank 0 only

One "step" = one detokenize pass over one batch of synthetic token streams with the synthetic
(random-init) BiCodec checkpoint.  Per-GPU workload (every N): BASELINE config 2 = batch 64 x 10 s
(500 semantic + 32 global tokens per utterance), fp32 parity mode; the batch is sharded by utterance,
no data-path collective ("scaling": "weak").  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
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


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def _ncu_traffic(kernel_name: str, args):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same
    command (profiles/r1_traffic.json: {"<precision> b<batch> t<frames>": {"<kernel name>": bytes}}), else None."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return d[f"{args.precision} b{args.batch} t{args.frames}"].get(kernel_name)
    except (OSError, KeyError, ValueError):
        return None


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


def cpu_reference(cfg, sd, steps: int, warmup: int, batch: int = 1, frames: int = 500):
    """The reference's CPU detokenize (oracle port: same ATen/oneDNN op sequence as
    sparktts/models/bicodec.py:171-189) on all host cores.  Returns (audio_s_per_s, info)."""
    import torch
    from oracle import bicodec_oracle as O                       # cpu_baseline leg only
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
    ts.sort()
    med = ts[len(ts) // 2]
    cpu_model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for l in f:
                if l.startswith("model name"):
                    cpu_model = l.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    audio_s = batch * frames / FRAME_RATE
    return audio_s / med, dict(cores=cores, cpu=cpu_model, sample=f"{batch} utterance x {frames / FRAME_RATE:.0f} s "
                               f"(BASELINE config 1), median of {steps} runs after {warmup} warm-up", ms=med * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=500, help="token frames per utterance (50 Hz)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()

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
    config = {"workload": workload, "utterances_per_gpu": args.batch, "frames": args.frames,
              "global_batch": args.batch * max(world, 1), "sharding": "by utterance, no collective",
              "precision_mode": args.precision, "weights": "synthetic random-init BiCodec (seed 0)",
              "l2": "not flushed: per-step activation working set (~16 GB) >> 126 MB L2"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sd = synthetic_state_dict(cfg, 0)
        steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        v, info = cpu_reference(cfg, sd, steps, warmup)
        config = dict(config, workload=workload + " [timed on a bounded sample: " + info["sample"] + "]")
        line = {"impl": "reference", "metric": "audio-sec decoded/sec (BiCodec detokenize)", "value": v,
                "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": info["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": info["cores"], "kind": "port",
                                 "sample": info["sample"], "cpu": info["cpu"]},
                "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    from spark_tts_b200 import BiCodec, BiCodecTokenizer

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

    # ---- device-resident throughput ("value") ----
    model.validate_tokens = False
    for _ in range(max(args.warmup, 3)):
        model.detokenize(sem_d, glob_d)
    model.check_tokens()
    sampler = ClockSampler(local_rank)
    launches0 = model.launch_count()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wav = model.detokenize(sem_d, glob_d)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = reduce_max(e0.elapsed_time(e1))
    launches = model.launch_count() - launches0
    value = world * audio_s_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the public façade, host buffers in and out ("e2e") ----
    for _ in range(2):
        tok.detokenize_pinned(glob_p, sem_p)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = tok.detokenize_pinned(glob_p, sem_p)
    e1.record()
    barrier()
    ms_e2e = reduce_max(max(e0.elapsed_time(e1), 0.0))
    wall_e2e = reduce_max((time.perf_counter() - t0) * 1e3)
    ms_e2e = max(ms_e2e, wall_e2e)       # the D2H + sync is host-visible: take the slower clock
    e2e = {"value": world * audio_s_per_step * args.steps / (ms_e2e * 1e-3), "unit": "audio-s/s",
           "h2d_bytes_per_step": sem_p.numel() * sem_p.element_size() + glob_p.numel() * glob_p.element_size(),
           "d2h_bytes_per_step": out.numel() * 4, "ms_per_step": ms_e2e / args.steps}

    line = {"metric": "audio-sec decoded/sec (BiCodec detokenize)", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (bf16x3 split products on tcgen05, fp32 accumulate)" if args.precision == "fp32" else "bf16",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches}

    if rank == 0:
        peaks = _peaks()
        # ---- roofline of the dominant kernel: per-launch CUDA events in a separate, untimed pass ----
        if not args.no_profile:
            n_prof = 3
            rows = None
            for _ in range(n_prof):                      # average launch duration over n_prof profiled passes
                model.profile(True)
                model.detokenize(sem_d, glob_d)
                cur = model.profile_read()
                model.profile(False)
                if rows is None:
                    rows = cur
                else:
                    for a, b in zip(rows, cur):
                        a["ms"] += b["ms"]
            for r in rows:
                r["ms"] /= n_prof
            # tensor-pipe kernels: the generic tcgen05 conv kernel and the fused ResidualUnit kernel
            gemm = [r for r in rows if r["name"].startswith(("conv_gemm_tc", "resunit_fused"))]
            agg = {}
            for r in gemm:
                a = agg.setdefault(r["name"], dict(ms=0.0, flops=0.0, n=0))
                a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["n"] += 1
            total_ms = sum(r["ms"] for r in rows)
            gemm_ms = sum(r["ms"] for r in gemm)
            gemm_flops = sum(r["flops"] for r in gemm)
            top_name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
            work = 3.0 if args.precision == "fp32" else 1.0
            ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
            line["roofline"] = {
                "bound": "tensor", "kernel": top_name, "launches": top["n"],
                "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                "traffic": _ncu_traffic(top_name, args), "peak_source": peaks["source"] + ", sustained bf16",
                "tensor_work_factor": work, "tensor_pipe_frac": ach * work / peaks["tf_sustained"],
                "all_gemm": {"achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12,
                             "frac": gemm_flops / (gemm_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                             "tensor_pipe_frac": gemm_flops * work / (gemm_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                             "share_of_step": gemm_ms / total_ms},
                "note": "achieved = algorithmic 2*MAC of the convolution / CUDA-event duration; the fp32 mode issues "
                        "3 bf16 MMAs per algorithmic MAC (tensor_work_factor)"}
            fused = [r for r in rows if r["name"].startswith("resunit_fused")]
            if fused:
                f_ms, f_fl = sum(r["ms"] for r in fused), sum(r["flops"] for r in fused)
                line["roofline"]["resunit_fused"] = {
                    "launches": len(fused), "achieved": f_fl / (f_ms * 1e-3) / 1e12,
                    "tensor_pipe_frac": f_fl * work / (f_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                    "share_of_step": f_ms / total_ms}
            stream = {}
            for r in rows:
                key = r["name"] if r["name"] in ("head", "dwconv_ln", "ln") else None
                if r["name"].startswith("conv_gemm_tc") and " taps=1 " in r["name"] and r["name"].endswith("res=1"):
                    key = "conv1x1_residual (conv_gemm_tc, HBM-bound class)"
                if key:
                    a = stream.setdefault(key, dict(ms=0.0, bytes=0.0, n=0))
                    a["ms"] += r["ms"]; a["bytes"] += r["bytes"]; a["n"] += 1
            line["hbm_kernels"] = {k: {"launches": v["n"], "achieved_gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                       "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                       "share_of_step": v["ms"] / total_ms} for k, v in stream.items()}
        # ---- CPU baseline beside it (bounded sample) ----
        if not args.no_cpu_baseline and world == 1:
            v, info = cpu_reference(cfg, sd, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": info["cores"], "kind": "port",
                                    "sample": info["sample"], "cpu": info["cpu"]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
