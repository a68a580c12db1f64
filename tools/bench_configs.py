"""Measures BASELINE.json configs 3-5 (bench.py covers configs 1-2).  One JSON line per config (rank 0).

    python tools/bench_configs.py c4                                   # 1 GPU: streaming chunk latency
    torchrun --nproc-per-node N ... tools/bench_configs.py c3 c5       # N GPUs: utterance / time sharding

c3: batch 1024 x 30 s, utterance-sharded (1024/8 = 128 utterances per GPU at every N: weak scaling)
c4: 256 concurrent streams x 50-token chunks, p50/p99 of "chunk tokens on device -> waveform in pinned host"
c5: batch 32 x 120 s, each utterance time-sharded over the N GPUs with NCCL halo exchange; checked
    against the communication-free recompute variant on the same shard.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from spark_tts_b200 import BiCodec, BiCodecConfig, sharding
from spark_tts_b200.streaming import StreamingDetokenizer
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c4"]
    precision = "bf16" if "--bf16" in sys.argv else "fp32"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = BiCodecConfig()
    model = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev, precision=precision)
    model.validate_tokens = False

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def tmax(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        return tmax(e0.elapsed_time(e1)) / steps

    for cfg_name in which:
        if cfg_name == "c3":
            B, T = 128, 1500
            sem, glob = synthetic_tokens(cfg, B, T, 3000 + rank)
            sem, glob = sem.to(dev), glob.to(dev)
            ms = timed(lambda: model.detokenize(sem, glob), steps=2, warmup=1)
            line = {"config": "c3: batch 1024 x 30 s utterance-sharded (128 utterances per GPU)", "n_gpus": world,
                    "precision": precision, "ms_per_step": ms, "audio_s_per_s": world * B * T / 50.0 / (ms * 1e-3),
                    "scaling": "weak", "collectives_on_data_path": 0}
        elif cfg_name == "c4":
            S, T = 256, 50
            st = StreamingDetokenizer(model, use_graphs=True)
            sem, glob = synthetic_tokens(cfg, S, T, 4000 + rank)
            semd, globd = sem.to(dev), glob.squeeze(1).to(dev)
            for _ in range(5):
                st.decode_batch(semd, globd)
            lat = []
            for i in range(200):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                st.decode_batch(semd, globd)           # ends with the D2H copy + stream sync
                lat.append((time.perf_counter() - t0) * 1e3)
            lat.sort()
            eager = StreamingDetokenizer(model, use_graphs=False)
            for _ in range(3):
                eager.decode_batch(semd, globd)
            le = []
            for i in range(50):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                eager.decode_batch(semd, globd)
                le.append((time.perf_counter() - t0) * 1e3)
            le.sort()
            line = {"config": "c4: 256 streams x 50-token chunk (1 s audio each), CUDA-graph replay", "n_gpus": 1,
                    "precision": precision, "rounds": len(lat), "p50_ms": lat[len(lat) // 2],
                    "p99_ms": lat[int(len(lat) * 0.99) - 1], "max_ms": lat[-1],
                    "eager_p50_ms": le[len(le) // 2], "audio_s_per_s": S * T / 50.0 / (lat[len(lat) // 2] * 1e-3),
                    "timed": "chunk tokens on device -> waveform chunk in pinned host memory (sync included)"}
        elif cfg_name == "c5":
            B, T = 32, 6000
            sem, glob = synthetic_tokens(cfg, B, T, 5000)      # same tokens on every rank (replicated)
            sem, glob = sem.to(dev), glob.to(dev)
            wav, (a, b) = sharding.detokenize_time_sharded(model, sem, glob, exchange=True)
            ref, _ = sharding.detokenize_time_sharded(model, sem, glob, exchange=False)
            err = (wav - ref).abs().max().item()
            ms = timed(lambda: sharding.detokenize_time_sharded(model, sem, glob, exchange=True), steps=2, warmup=1)
            ms_nc = timed(lambda: sharding.detokenize_time_sharded(model, sem, glob, exchange=False), steps=2, warmup=1)
            ph, wh = model.halo_frames()
            line = {"config": "c5: batch 32 x 120 s, time-sharded with NCCL halo exchange", "n_gpus": world,
                    "precision": precision, "frames_per_rank": b - a, "halo_frames": {"prenet": ph, "wavegen": wh},
                    "halo_bytes_per_neighbour": B * wh * cfg.d_model * 4, "ms_per_step": ms,
                    "audio_s_per_s": B * T / 50.0 / (ms * 1e-3), "ms_per_step_recompute_variant": ms_nc,
                    "max_abs_diff_vs_recompute_variant": tmax(err), "scaling": "strong"}
        else:
            raise SystemExit(f"unknown config {cfg_name}")
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
