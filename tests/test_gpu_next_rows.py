"""GPU tests of the rows next to the hot path: checkpoint ingestion through the reference's
directory layout, ragged-length batching, and the on-device token feed."""
import numpy as np
import pytest
import torch

from test_next_rows_cpu import _write_checkpoint

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(cfg, state_dict, dev):
    from spark_tts_b200 import BiCodec
    return BiCodec.from_state_dict(cfg, state_dict, device=dev)


def test_load_from_checkpoint_dir_equals_from_state_dict(tmp_path, cfg, state_dict, model, dev):
    """BiCodecTokenizer(model_dir) -> BiCodec.load_from_checkpoint(model_dir/BiCodec) (audio_tokenizer.py:46-56,
    bicodec.py:69-111): yaml + safetensors in, the same waveform as the in-memory state dict, bit for bit."""
    from spark_tts_b200 import BiCodec, BiCodecTokenizer
    from spark_tts_b200.synthetic import synthetic_tokens
    root = _write_checkpoint(tmp_path, cfg, state_dict)
    m2 = BiCodec.load_from_checkpoint(str(root / "BiCodec"), device=dev)
    sem, glob = synthetic_tokens(cfg, 2, 23, 71)
    a = model.detokenize(sem.to(dev), glob.to(dev))
    b = m2.detokenize(sem.to(dev), glob.to(dev))
    assert torch.equal(a, b)
    tok = BiCodecTokenizer(str(root), device=dev)
    w = tok.detokenize(glob.squeeze(1).to(dev), sem.to(dev))
    assert isinstance(w, np.ndarray) and np.array_equal(w, a.squeeze(1).cpu().numpy())


def test_missing_tensor_is_reported_by_key(cfg, state_dict, dev):
    from spark_tts_b200 import BiCodec
    sd = {k: v for k, v in state_dict.items() if k != "decoder.model.3.block.2.block.1.weight_g"}
    with pytest.raises(ValueError, match="decoder.model.3.block.2.block.1"):
        BiCodec.from_state_dict(cfg, sd, device=dev)


def test_ragged_lengths_are_bucketed_and_exact(model, cfg, dev):
    """Mixed lengths in one call (the reference's Triton vocoder can only torch.cat equal lengths,
    runtime/triton_trtllm/model_repo/vocoder/1/model.py:72-106): every waveform equals the utterance decoded alone."""
    from spark_tts_b200.synthetic import synthetic_tokens
    lens = [50, 17, 50, 0, 131, 17]
    sem, glob = synthetic_tokens(cfg, len(lens), max(lens), 81)
    semd, globd = sem.to(dev), glob.to(dev)
    rows = [semd[i, :n] for i, n in enumerate(lens)]
    n0 = model.launch_count()
    outs = model.detokenize_ragged(rows, globd)
    per_pass = None
    assert len(outs) == len(lens)
    for i, n in enumerate(lens):
        assert outs[i].shape == (n * cfg.hop,)
        if n:
            alone = model.detokenize(semd[i:i + 1, :n].contiguous(), globd[i:i + 1])
            assert torch.equal(alone[0, 0], outs[i])
    del n0, per_pass


def test_token_feed_matches_the_regex_path(model, cfg, dev):
    """ids -> codes on the device == decode-to-text + regex on the host (cli/SparkTTS.py:213-228), then the
    codes go straight into detokenize without touching the host."""
    from spark_tts_b200 import token_feed
    rng = np.random.default_rng(5)
    sem_base, glob_base, B, N = 151_700, 151_665 - 4096, 3, 700      # arbitrary contiguous id ranges
    ids = np.full((B, N), 7, dtype=np.int64)
    want_s, want_g = [], []
    for b in range(B):
        kinds = rng.choice(4, size=N, p=[0.15, 0.7, 0.1, 0.05])
        s_codes = rng.integers(0, cfg.codebook_size, size=N)
        g_codes = rng.integers(0, 4096, size=N)
        other = rng.integers(0, 100_000, size=N)
        ids[b] = np.where(kinds == 1, sem_base + s_codes, np.where(kinds == 2, glob_base + g_codes, other))
        ids[b, -1] = sem_base + cfg.codebook_size          # one past the range: not a code
        text = "".join(f"<|bicodec_semantic_{i - sem_base}|>" if sem_base <= i < sem_base + cfg.codebook_size
                       else f"<|bicodec_global_{i - glob_base}|>" if glob_base <= i < glob_base + 4096 else "tok"
                       for i in ids[b])
        s, g = token_feed.codes_from_text(text)
        want_s.append(s[0].tolist()); want_g.append(g[0, 0].tolist())
    for dt in (torch.int64, torch.int32):
        sem, sem_len, glob, glob_len = token_feed.codes_from_token_ids(torch.from_numpy(ids).to(dev, dt), sem_base, glob_base,
                                                                       cfg.codebook_size, 4096, max_global=32)
        assert sem.dtype == torch.int32 and sem_len.tolist() == [len(w) for w in want_s]
        assert glob_len.tolist() == [len(w) for w in want_g]
        for b in range(B):
            assert sem[b, :sem_len[b]].tolist() == want_s[b]
            assert glob[b].tolist() == want_g[b][:32]
    rows = token_feed.split_ragged(sem, sem_len)
    outs = model.detokenize_ragged(rows, glob)
    assert [o.numel() for o in outs] == [len(w) * cfg.hop for w in want_s]
    assert all(bool(torch.isfinite(o).all()) for o in outs)
