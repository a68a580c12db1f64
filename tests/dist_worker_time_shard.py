"""Worker of tests/test_gpu_nccl_sharding.py: launched by torchrun with one rank per GPU (NCCL).  Every rank decodes
its time shard of the same utterances with the CUDA path; rank 0 gathers the shards and checks them against the
un-sharded CUDA decode (max-abs <= 2e-6) and against the oracle (fp32 bar).  Also checks utterance sharding.
Prints one line "NCCL_SHARDING_OK ..." on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from spark_tts_b200 import BiCodec, BiCodecConfig, sharding
    from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens
    cfg = BiCodecConfig()
    sd = synthetic_state_dict(cfg, 0)
    model = BiCodec.from_state_dict(cfg, sd, device=dev)
    B, T = 2, 400
    sem, glob = synthetic_tokens(cfg, B, T, 6006)                      # the same tokens on every rank
    semd, globd = sem.to(dev), glob.to(dev)
    results = {}
    for name, exchange in (("exchange", True), ("recompute", False)):
        wav, (a, b) = sharding.detokenize_time_sharded(model, semd, globd, exchange=exchange)
        assert wav.shape == (B, 1, (b - a) * cfg.hop)
        n_max = max(sharding.shard_bounds(T, world, r)[1] - sharding.shard_bounds(T, world, r)[0] for r in range(world))
        pad = torch.zeros((B, 1, n_max * cfg.hop), device=dev)
        pad[:, :, : wav.shape[2]] = wav
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        full = torch.cat([p[:, :, : (sharding.shard_bounds(T, world, r)[1] - sharding.shard_bounds(T, world, r)[0]) * cfg.hop]
                          for r, p in enumerate(parts)], dim=2)
        results[name] = full
    whole = model.detokenize(semd, globd)
    # utterance sharding: gathered result == un-sharded decode, bit for bit (batch-invariant kernels)
    gathered, _ = sharding.detokenize_utterance_sharded(model, semd, globd, gather=True)
    assert torch.equal(gathered, whole), "utterance-sharded result differs from the un-sharded one"
    if rank == 0:
        from oracle import bicodec_oracle as O                         # checker
        torch.set_num_threads(8)
        ref = O.detokenize(sd, cfg, sem, glob)
        for name, full in results.items():
            err = (full - whole).abs().max().item()
            assert err <= 2e-6, f"{name}: sharded vs un-sharded max-abs {err}"
            e2 = (ref - full.cpu()).abs().max().item()
            snr = O.snr_db(ref, full.cpu())
            assert e2 <= 1e-3 and snr >= 60.0, f"{name}: vs oracle max_abs={e2:.3e} snr={snr:.1f}"
        print(f"NCCL_SHARDING_OK world={world} launches={model.launch_count()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
