"""BiCodec detokenize-path configuration.

Mirrors the ``audio_tokenizer`` section of the reference's ``BiCodec/config.yaml`` that
``BiCodec.load_from_checkpoint`` consumes (/root/reference sparktts/models/bicodec.py:81-88).
Only the keys the detokenize path needs are kept; the values below are the Spark-TTS-0.5B
model-card values (SURVEY.md §8d "Config provenance") and stay configuration, not constants.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List


@dataclass
class BiCodecConfig:
    # quantizer (sparktts/modules/vq/factorized_vector_quantize.py:35-68)
    d_model: int = 1024
    codebook_size: int = 8192
    codebook_dim: int = 8
    # speaker_encoder (sparktts/modules/speaker/speaker_encoder.py:44-69)
    fsq_levels: List[int] = field(default_factory=lambda: [4, 4, 4, 4, 4, 4])
    token_num: int = 32
    latent_dim: int = 128
    # prenet (sparktts/modules/encoder_decoder/feat_decoder.py:34-76)
    vocos_dim: int = 384
    vocos_intermediate_dim: int = 2048
    vocos_num_layers: int = 12
    downsample_layers: int = 2       # VocosBackbone(num_layers=2) inside each downsample stage
    sample_ratios: List[int] = field(default_factory=lambda: [1, 1])
    # decoder / WaveGenerator (sparktts/modules/encoder_decoder/wave_generator.py:56-83)
    dec_channels: int = 1536
    rates: List[int] = field(default_factory=lambda: [8, 5, 4, 2])
    kernel_sizes: List[int] = field(default_factory=lambda: [16, 11, 8, 4])
    sample_rate: int = 16000
    # mel_params of config.yaml (BiCodec.init_mel_transformer, sparktts/models/bicodec.py:191-211): input of the
    # speaker half of tokenize; and the reference-clip rule of BiCodecTokenizer.get_ref_clip (audio_tokenizer.py:57-71)
    mel_n_fft: int = 1024
    mel_win_length: int = 640
    mel_hop_length: int = 320
    mel_fmin: float = 10.0
    mel_fmax: Any = None
    num_mels: int = 128
    ref_segment_duration: float = 6.0
    latent_hop_length: int = 320
    volume_normalize: bool = True

    @property
    def hop(self) -> int:
        h = 1
        for r in self.rates:
            h *= r
        return h

    @property
    def frame_rate(self) -> float:
        return self.sample_rate / self.hop

    @classmethod
    def from_yaml_dict(cls, cfg: Dict[str, Any]) -> "BiCodecConfig":
        """Build from the ``audio_tokenizer`` dict of the reference's BiCodec/config.yaml."""
        q, s, p, d = cfg["quantizer"], cfg["speaker_encoder"], cfg["prenet"], cfg["decoder"]
        ratios = list(p.get("sample_ratios", [1, 1]))
        if any(r != 1 for r in ratios):
            raise ValueError("only sample_ratios == [1, 1] (the released BiCodec) is supported")
        if p.get("use_tanh_at_final", False):
            raise ValueError("prenet.use_tanh_at_final=True is not supported")
        if int(s.get("fsq_num_quantizers", 1)) != 1:
            raise ValueError("only fsq_num_quantizers == 1 is supported")
        return cls(
            d_model=int(q["input_dim"]),
            codebook_size=int(q["codebook_size"]),
            codebook_dim=int(q["codebook_dim"]),
            fsq_levels=[int(v) for v in s["fsq_levels"]],
            token_num=int(s["token_num"]),
            latent_dim=int(s["latent_dim"]),
            vocos_dim=int(p["vocos_dim"]),
            vocos_intermediate_dim=int(p["vocos_intermediate_dim"]),
            vocos_num_layers=int(p["vocos_num_layers"]),
            sample_ratios=ratios,
            dec_channels=int(d["channels"]),
            rates=[int(v) for v in d["rates"]],
            kernel_sizes=[int(v) for v in d["kernel_sizes"]],
            **cls._mel_fields(cfg),
        )

    @staticmethod
    def _mel_fields(cfg: Dict[str, Any]) -> Dict[str, Any]:
        m = cfg.get("mel_params")
        if not m:
            return {}
        out = dict(mel_n_fft=int(m["n_fft"]), mel_win_length=int(m.get("win_length", m["n_fft"])),
                   mel_hop_length=int(m.get("hop_length", int(m["n_fft"]) // 4)), mel_fmin=float(m.get("mel_fmin", 0.0)),
                   mel_fmax=(None if m.get("mel_fmax") is None else float(m["mel_fmax"])), num_mels=int(m["num_mels"]))
        if "sample_rate" in m:
            out["sample_rate"] = int(m["sample_rate"])
        return out


def load_bicodec_yaml(path: str) -> BiCodecConfig:
    """PyYAML replacement for the reference's OmegaConf ``load_config`` (sparktts/utils/file.py:116-130)."""
    import yaml

    with open(path) as f:
        cfg = yaml.safe_load(f)
    if "audio_tokenizer" in cfg:
        cfg = cfg["audio_tokenizer"]
    return BiCodecConfig.from_yaml_dict(cfg)
