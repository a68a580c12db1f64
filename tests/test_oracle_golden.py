"""CPU: the oracle restatement reproduces the golden vectors generated from the reference's own
modules (tests/golden/make_golden.py) -- this is what pins the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from oracle import bicodec_oracle as O


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("detok_")[-1][:-4])
def test_oracle_matches_reference_golden(path, cfg, state_dict):
    g = np.load(path)
    assert int(g["weight_seed"]) == 0
    sem = torch.from_numpy(g["semantic_tokens"])
    glob = torch.from_numpy(g["global_tokens"])
    taps = {}
    wav = O.detokenize(state_dict, cfg, sem, glob, taps)
    # integer stages: bit exact
    rows = torch.nn.functional.embedding(sem.long(), state_dict["quantizer.codebook.weight"])
    assert np.array_equal(rows.numpy(), g["codebook_rows"])
    codes = O.fsq_codes(glob.long().transpose(1, 2).squeeze(-1), cfg.fsq_levels)
    assert np.array_equal(codes.numpy(), g["fsq_codes"])
    # floating-point stages: identical ATen ops; allow only oneDNN batch-shape round-off
    assert O.snr_db(torch.from_numpy(g["d_vector"]), taps["d_vector"]) > 120
    assert O.snr_db(torch.from_numpy(g["z_q_first8"]), taps["z_q"].transpose(1, 2)[:, :, :8]) > 120
    assert O.snr_db(torch.from_numpy(g["prenet_plus_d_first8"]),
                    taps["prenet_plus_d"].transpose(1, 2)[:, :, :8]) > 100
    ref = torch.from_numpy(g["output_waveform"])
    assert wav.shape == ref.shape and wav.dtype == torch.float32
    assert (wav - ref).abs().max().item() <= 1e-4
    assert O.snr_db(ref, wav) >= 90.0


def test_fsq_round_trip_invariant(cfg):
    """The reference's own stated invariant (residual_fsq.py:430-432): codes -> indices -> codes."""
    levels = cfg.fsq_levels
    n = 1
    for l in levels:
        n *= l
    idx = torch.arange(n)
    codes = O.fsq_codes(idx, levels)
    lv = torch.tensor(levels)
    basis = torch.cumprod(torch.tensor([1] + levels[:-1]), 0)
    back = ((codes * (lv // 2) + (lv // 2)) * basis).sum(-1).to(torch.int64)
    assert torch.equal(back, idx)
    assert set(codes.unique().tolist()) == {-1.0, -0.5, 0.0, 0.5}


def test_tokenizer_facade_shapes(cfg, state_dict):
    """audio_tokenizer.py:132-146: (B,320T) numpy, squeezed to (320T,) for B == 1."""
    from spark_tts_b200.synthetic import synthetic_tokens
    sem, glob = synthetic_tokens(cfg, 1, 3, 7)
    w = O.tokenizer_detokenize(state_dict, cfg, glob.squeeze(1), sem)
    assert w.shape == (3 * cfg.hop,) and w.dtype == np.float32
    sem, glob = synthetic_tokens(cfg, 2, 3, 7)
    w = O.tokenizer_detokenize(state_dict, cfg, glob.squeeze(1), sem)
    assert w.shape == (2, 3 * cfg.hop)
