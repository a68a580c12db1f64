"""CPU tests of the rows next to the hot path: checkpoint ingestion (config.yaml +
model.safetensors exactly as the reference lays them out) and the host mirror of the reference's token parsing."""
import os
import re

import pytest
import torch


def _write_checkpoint(tmp_path, cfg, state_dict):
    """<root>/BiCodec/{config.yaml, model.safetensors} -- the layout BiCodecTokenizer(model_dir) expects
    (sparktts/models/audio_tokenizer.py:46-56, sparktts/models/bicodec.py:69-111)."""
    import yaml
    from safetensors.torch import save_file

    d = tmp_path / "BiCodec"
    d.mkdir()
    doc = {"audio_tokenizer": {
        "mel_params": {"sample_rate": 16000, "n_fft": 1024, "win_length": 640, "hop_length": 320, "mel_fmin": 10, "mel_fmax": None,
                       "num_mels": 128},
        "encoder": {"input_channels": 1024, "vocos_dim": 384, "vocos_intermediate_dim": 2048, "vocos_num_layers": 12,
                    "out_channels": 1024, "sample_ratios": [1, 1]},
        "decoder": {"input_channel": cfg.d_model, "channels": cfg.dec_channels, "rates": list(cfg.rates),
                    "kernel_sizes": list(cfg.kernel_sizes)},
        "quantizer": {"input_dim": cfg.d_model, "codebook_size": cfg.codebook_size, "codebook_dim": cfg.codebook_dim,
                      "commitment": 0.25, "codebook_loss_weight": 2.0, "use_l2_normlize": True, "threshold_ema_dead_code": 0.2},
        "speaker_encoder": {"input_dim": 128, "out_dim": cfg.d_model, "latent_dim": cfg.latent_dim,
                            "token_num": cfg.token_num, "fsq_levels": list(cfg.fsq_levels), "fsq_num_quantizers": 1},
        "prenet": {"input_channels": cfg.d_model, "vocos_dim": cfg.vocos_dim,
                   "vocos_intermediate_dim": cfg.vocos_intermediate_dim, "vocos_num_layers": cfg.vocos_num_layers,
                   "out_channels": cfg.d_model, "condition_dim": cfg.d_model, "sample_ratios": [1, 1],
                   "use_tanh_at_final": False},
        "postnet": {"input_channels": 1024, "vocos_dim": 384, "vocos_intermediate_dim": 2048, "vocos_num_layers": 6,
                    "out_channels": 1024, "use_tanh_at_final": False}}}
    with open(d / "config.yaml", "w") as f:
        yaml.safe_dump(doc, f)
    # a real checkpoint also carries encode-side / training-only tensors: they must be ignored
    extra = {"encoder.embed.weight": torch.zeros(4, 4, 7), "quantizer.cluster_size": torch.zeros(8),
             "postnet.linear.weight": torch.zeros(2, 2)}
    save_file({**{k: v.contiguous() for k, v in state_dict.items()}, **extra}, str(d / "model.safetensors"))
    return tmp_path


def test_config_yaml_round_trip(tmp_path, cfg, state_dict):
    from spark_tts_b200.config import load_bicodec_yaml
    root = _write_checkpoint(tmp_path, cfg, state_dict)
    got = load_bicodec_yaml(os.path.join(root, "BiCodec", "config.yaml"))
    assert got == cfg and got.hop == 320 and got.frame_rate == 50.0


def test_config_rejects_what_the_path_does_not_implement(cfg):
    from spark_tts_b200.config import BiCodecConfig
    base = {"quantizer": {"input_dim": 1024, "codebook_size": 8192, "codebook_dim": 8},
            "speaker_encoder": {"fsq_levels": [4] * 6, "token_num": 32, "latent_dim": 128},
            "prenet": {"vocos_dim": 384, "vocos_intermediate_dim": 2048, "vocos_num_layers": 12, "sample_ratios": [1, 1]},
            "decoder": {"channels": 1536, "rates": [8, 5, 4, 2], "kernel_sizes": [16, 11, 8, 4]}}
    assert BiCodecConfig.from_yaml_dict(base) == cfg
    for key, val in (("sample_ratios", [2, 2]), ("use_tanh_at_final", True)):
        bad = {**base, "prenet": {**base["prenet"], key: val}}
        with pytest.raises(ValueError):
            BiCodecConfig.from_yaml_dict(bad)


def test_safetensors_keys_match_what_the_library_consumes(tmp_path, cfg, state_dict):
    """Every tensor of the detokenize path survives the safetensors round trip bit-exactly under the
    reference's key names (weight_g / weight_v for the weight-normed convs)."""
    from safetensors.torch import load_file
    root = _write_checkpoint(tmp_path, cfg, state_dict)
    sd = load_file(os.path.join(root, "BiCodec", "model.safetensors"))
    for k, v in state_dict.items():
        assert torch.equal(sd[k], v), k
    assert "decoder.model.1.block.1.weight_g" in sd and tuple(sd["decoder.model.1.block.1.weight_g"].shape) == (1536, 1, 1)


def test_codes_from_text_is_the_reference_regex():
    """cli/SparkTTS.py:213-228: re.findall over the decoded text, semantic as (1, T) long, global as (1, 1, G)."""
    from spark_tts_b200.token_feed import codes_from_text
    text = ("<|start_global_token|><|bicodec_global_3|><|bicodec_global_4095|><|end_global_token|>"
            "<|start_semantic_token|><|bicodec_semantic_17|><|bicodec_semantic_0|>junk<|bicodec_semantic_8191|>")
    sem, glob = codes_from_text(text)
    want_s = [int(t) for t in re.findall(r"bicodec_semantic_(\d+)", text)]
    want_g = [int(t) for t in re.findall(r"bicodec_global_(\d+)", text)]
    assert sem.dtype == torch.int64 and sem.shape == (1, 3) and sem[0].tolist() == want_s == [17, 0, 8191]
    assert glob.shape == (1, 1, 2) and glob[0, 0].tolist() == want_g == [3, 4095]
    sem, glob = codes_from_text("no codes here")
    assert sem.shape == (1, 0) and glob.shape == (1, 1, 0)


def test_token_feed_refuses_cpu_tensors():
    from spark_tts_b200.token_feed import codes_from_token_ids
    with pytest.raises(RuntimeError):
        codes_from_token_ids(torch.zeros((1, 4), dtype=torch.int64), 100, 9000)


def test_example_wav_writer_round_trips(tmp_path):
    """examples/vocode_tokens.py writes what the reference CLI hands to soundfile: 16 kHz mono, here as 16-bit PCM."""
    import importlib.util
    import os
    import wave

    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("vocode_tokens", os.path.join(root, "examples", "vocode_tokens.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    x = np.sin(np.linspace(0, 40, 3200)).astype(np.float32) * 0.5
    x[:3] = [1.5, -1.5, 0.0]                       # clipped, not wrapped
    mod.write_wav(str(tmp_path / "a.wav"), x, 16000)
    with wave.open(str(tmp_path / "a.wav")) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 3200)
        pcm = np.frombuffer(w.readframes(3200), dtype="<i2")
    assert pcm[0] == 32767 and pcm[1] == -32767 and pcm[2] == 0
    assert np.abs(pcm[3:] / 32767.0 - x[3:]).max() < 1e-4
    with __import__("pytest").raises(SystemExit):
        mod.main(["--out", str(tmp_path / "b.wav")])      # neither --synthetic nor a model dir
