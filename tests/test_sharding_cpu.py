"""CPU, world_size 2, gloo: the utterance / time partition and the halo exchange (spark_tts_b200.sharding)
with the oracle plugged in as the compute backend.  The native library is not needed here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spark_tts_b200.sharding import shard_bounds


class OracleBackend:
    """Same duck-type as spark_tts_b200.BiCodec (prenet / wavegen / detokenize / halo_frames / hop)."""

    def __init__(self, cfg, sd):
        self.cfg, self.sd, self.hop = cfg, sd, cfg.hop

    def halo_frames(self):
        return 57, 11

    def detokenize(self, sem, glob):
        from oracle import bicodec_oracle as O
        return O.detokenize(self.sd, self.cfg, sem, glob)

    def prenet(self, sem, glob):
        from oracle import bicodec_oracle as O
        with torch.no_grad():
            z = O.vq_detokenize(self.sd, sem)
            d = O.speaker_detokenize(self.sd, glob, self.cfg.fsq_levels)
            return (O.prenet(self.sd, z, d) + d.unsqueeze(-1)).transpose(1, 2).contiguous()

    def wavegen(self, x):
        from oracle import bicodec_oracle as O
        with torch.no_grad():
            return O.wave_generator(self.sd, x.transpose(1, 2), self.cfg.rates, self.cfg.kernel_sizes)


class StagedOracleBackend(OracleBackend):
    """Adds the staged-input surface of the native BiCodec (can_stage / wavegen_stage / wavegen_staged), so the
    interior-first ordering of spark_tts_b200.sharding is exercised on CPU too."""

    def __init__(self, cfg, sd):
        super().__init__(cfg, sd)
        self.buf, self.filled, self.log = None, None, []

    def can_stage(self, batch, frames_total):
        return True

    def wavegen_stage(self, x_rows, frames_total, row_offset):
        B, rows, D = x_rows.shape
        if self.buf is None or self.buf.shape != (B, frames_total, D):
            self.buf = torch.full((B, frames_total, D), float("nan"))
            self.filled = torch.zeros(frames_total, dtype=torch.bool)
        assert not self.filled[row_offset:row_offset + rows].any(), "rows staged twice"
        self.buf[:, row_offset:row_offset + rows] = x_rows
        self.filled[row_offset:row_offset + rows] = True
        self.log.append((row_offset, rows))

    def wavegen_staged(self, batch, frames_total):
        assert self.buf.shape[:2] == (batch, frames_total) and bool(self.filled.all()), "window not fully staged"
        x, self.buf, self.filled = self.buf, None, None
        return self.wavegen(x)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spark_tts_b200 import sharding
        from spark_tts_b200.config import BiCodecConfig
        from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens
        cfg = BiCodecConfig()
        staged = mode == "time_exchange_staged"
        model = (StagedOracleBackend if staged else OracleBackend)(cfg, synthetic_state_dict(cfg, 0))
        if mode == "utterance":
            sem, glob = synthetic_tokens(cfg, 3, 12, 21)
            full, rng = sharding.detokenize_utterance_sharded(model, sem, glob, gather=True)
            local, (lo, hi) = sharding.detokenize_utterance_sharded(model, sem, glob)
            assert (lo, hi) == shard_bounds(3, world, rank) and local.shape[0] == hi - lo
            assert torch.equal(full[lo:hi], local)
            if rank == 0:
                torch.save(dict(wav=full, sem=sem, glob=glob), os.path.join(out_dir, "out.pt"))
        else:
            sem, glob = synthetic_tokens(cfg, 1, 150, 22)
            wav, (a, b) = sharding.detokenize_time_sharded(model, sem, glob, exchange=mode.startswith("time_exchange"))
            assert wav.shape == (1, 1, (b - a) * cfg.hop)
            if staged:   # the own (interior) rows were staged BEFORE the halo that arrived over the wire
                assert model.log[0][1] == b - a and len(model.log) == 2 and model.log[1][1] == 11
            gathered = [None] * world
            dist.all_gather_object(gathered, (a, b, wav))
            if rank == 0:
                gathered.sort(key=lambda t: t[0])
                assert gathered[0][0] == 0 and gathered[-1][1] == 150
                torch.save(dict(wav=torch.cat([g[2] for g in gathered], dim=2), sem=sem, glob=glob),
                           os.path.join(out_dir, "out.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["utterance", "time_exchange", "time_exchange_staged", "time_recompute"])
def test_two_rank_sharding_matches_unsharded(mode, tmp_path, cfg, state_dict):
    from oracle import bicodec_oracle as O
    mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    out = torch.load(os.path.join(str(tmp_path), "out.pt"))
    ref = O.detokenize(state_dict, cfg, out["sem"], out["glob"])
    assert out["wav"].shape == ref.shape
    # same fp32 ATen ops on a different window: only oneDNN blocking round-off differs
    assert O.snr_db(ref, out["wav"]) > 95.0
    assert (ref - out["wav"]).abs().max().item() < 1e-4


@pytest.mark.parametrize("backend", [OracleBackend, StagedOracleBackend])
@pytest.mark.parametrize("exchange", [True, False])
def test_single_process_time_windows_match_unsharded(backend, exchange, cfg, state_dict):
    """detokenize_time_windows = the N-rank time-sharded schedule run by one process (long-form on one GPU)."""
    from oracle import bicodec_oracle as O
    from spark_tts_b200 import sharding
    from spark_tts_b200.synthetic import synthetic_tokens
    torch.set_num_threads(8)
    sem, glob = synthetic_tokens(cfg, 2, 100, 23)
    model = backend(cfg, state_dict)
    got = sharding.detokenize_time_windows(model, sem, glob, 3, exchange=exchange)
    ref = O.detokenize(state_dict, cfg, sem, glob)
    assert got.shape == ref.shape
    assert O.snr_db(ref, got) > 95.0 and (ref - got).abs().max().item() < 1e-4
    with pytest.raises(ValueError):
        sharding.detokenize_time_windows(model, sem[:, :20], glob, 3)          # windows shorter than the halo


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
