"""ORACLE tooling (this container only): build the reference's own detokenize modules from
/root/reference and load a checkpoint state dict into them.

/root/reference does not exist on the GPU box, so nothing under ``tests -m gpu``, ``smoke()`` or
``bench.py`` imports this file; it is used by ``oracle/validate_against_reference.py`` and
``tests/golden/make_golden.py`` and by CPU tests that skip when the reference is absent.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SPARKTTS_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sparktts"))


def _import_reference():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "omegaconf" not in sys.modules:      # not installed here; only imported for a type name
        stub = types.ModuleType("omegaconf")
        stub.DictConfig = dict
        stub.OmegaConf = object
        sys.modules["omegaconf"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class ReferenceDetokenizer(torch.nn.Module):
    """The four reference sub-modules of BiCodec.detokenize (sparktts/models/bicodec.py:171-189),
    composed exactly as BiCodecVocoderWrapper does (export_sparktts_onnx.py:267-312), i.e. with
    ``onnx_export_mode=True`` for the FSQ gather because ``einx`` is not installed here."""

    def __init__(self, cfg):
        super().__init__()
        _import_reference()
        from sparktts.modules.vq.factorized_vector_quantize import FactorizedVectorQuantize
        from sparktts.modules.speaker.speaker_encoder import SpeakerEncoder
        from sparktts.modules.encoder_decoder.feat_decoder import Decoder
        from sparktts.modules.encoder_decoder.wave_generator import WaveGenerator

        self.quantizer = FactorizedVectorQuantize(
            input_dim=cfg.d_model, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim,
            commitment=0.25)
        self.speaker_encoder = SpeakerEncoder(
            input_dim=128, out_dim=cfg.d_model, latent_dim=cfg.latent_dim, token_num=cfg.token_num,
            fsq_levels=list(cfg.fsq_levels), fsq_num_quantizers=1)
        self.prenet = Decoder(
            input_channels=cfg.d_model, vocos_dim=cfg.vocos_dim,
            vocos_intermediate_dim=cfg.vocos_intermediate_dim, vocos_num_layers=cfg.vocos_num_layers,
            out_channels=cfg.d_model, condition_dim=cfg.d_model, sample_ratios=list(cfg.sample_ratios),
            use_tanh_at_final=False)
        self.decoder = WaveGenerator(
            input_channel=cfg.d_model, channels=cfg.dec_channels, rates=list(cfg.rates),
            kernel_sizes=list(cfg.kernel_sizes))
        self.eval()

    def load_checkpoint(self, sd):
        missing, unexpected = self.load_state_dict(sd, strict=False)
        # everything in the synthetic checkpoint must be consumed; the only keys the checkpoint may
        # lack are encode-side ones (ECAPA/perceiver, in_project, cluster_size, project_in)
        assert not unexpected, unexpected
        on_path = [k for k in missing if not (
            k.startswith("speaker_encoder.speaker_encoder.") or k.startswith("speaker_encoder.perceiver_sampler.")
            or k.startswith("quantizer.in_project") or k == "quantizer.cluster_size"
            or k.startswith("speaker_encoder.quantizer.project_in"))]
        assert not on_path, on_path

        def _rm(m):  # bicodec.py:213-221
            try:
                torch.nn.utils.remove_weight_norm(m)
            except ValueError:
                pass
        self.apply(_rm)
        return self

    @torch.no_grad()
    def detokenize(self, semantic_tokens, global_tokens):
        z_q = self.quantizer.detokenize(semantic_tokens)
        d_vector = self.speaker_encoder.detokenize(global_tokens, onnx_export_mode=True)
        x = self.prenet(z_q, d_vector)
        x = x + d_vector.unsqueeze(-1)
        return self.decoder(x)


class ReferenceSemanticTokenizer(torch.nn.Module):
    """The reference modules behind the semantic half of BiCodec.tokenize (sparktts/models/bicodec.py:151-169):
    ``Encoder`` (feat_encoder.py) and ``FactorizedVectorQuantize`` (its in_project + codebook search)."""

    def __init__(self, cfg):
        super().__init__()
        _import_reference()
        from sparktts.modules.vq.factorized_vector_quantize import FactorizedVectorQuantize
        from sparktts.modules.encoder_decoder.feat_encoder import Encoder

        self.encoder = Encoder(
            input_channels=cfg.d_model, vocos_dim=cfg.vocos_dim, vocos_intermediate_dim=cfg.vocos_intermediate_dim,
            vocos_num_layers=cfg.vocos_num_layers, out_channels=cfg.d_model, sample_ratios=list(cfg.sample_ratios))
        self.quantizer = FactorizedVectorQuantize(
            input_dim=cfg.d_model, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim,
            commitment=0.25)
        self.eval()

    def load_checkpoint(self, sd):
        own = {k: v for k, v in sd.items() if k.startswith("encoder.") or k.startswith("quantizer.")}
        missing, unexpected = self.load_state_dict(own, strict=False)
        assert not unexpected, unexpected
        assert all(k == "quantizer.cluster_size" for k in missing), missing
        return self

    @torch.no_grad()
    def tokenize(self, feat):
        """feat (B,T,D) -> semantic tokens (B,T), exactly the two lines of bicodec.py:165-166."""
        z = self.encoder(feat.transpose(1, 2))
        return self.quantizer.tokenize(z)


class ReferenceSpeakerTokenizer(torch.nn.Module):
    """The reference modules behind the speaker half of BiCodec.tokenize (sparktts/models/bicodec.py:162-167):
    the torchaudio MelSpectrogram that ``init_mel_transformer`` builds (:191-211) and ``SpeakerEncoder.tokenize``
    (speaker_encoder.py:100-105: ECAPA-TDNN latent -> perceiver resampler -> ResidualFSQ indices)."""

    def __init__(self, cfg):
        super().__init__()
        _import_reference()
        import torchaudio.transforms as TT
        from sparktts.modules.speaker.speaker_encoder import SpeakerEncoder

        self.mel_transformer = TT.MelSpectrogram(
            cfg.sample_rate, cfg.mel_n_fft, cfg.mel_win_length, cfg.mel_hop_length, cfg.mel_fmin, cfg.mel_fmax,
            n_mels=cfg.num_mels, power=1, norm="slaney", mel_scale="slaney")
        self.speaker_encoder = SpeakerEncoder(
            input_dim=cfg.num_mels, out_dim=cfg.d_model, latent_dim=cfg.latent_dim, token_num=cfg.token_num,
            fsq_levels=list(cfg.fsq_levels), fsq_num_quantizers=1)
        self.eval()

    def load_checkpoint(self, sd):
        own = {k: v for k, v in sd.items() if k.startswith("speaker_encoder.")}
        missing, unexpected = self.load_state_dict(own, strict=False)
        assert not unexpected, unexpected
        # only the x-vector head (attentive pooling, bn, linear), which tokenize computes and discards, may be absent
        head = ("speaker_encoder.speaker_encoder.pool.", "speaker_encoder.speaker_encoder.bn.",
                "speaker_encoder.speaker_encoder.linear.", "mel_transformer.")   # (the mel buffers are derived, not loaded)
        assert all(k.startswith(head) for k in missing), missing
        return self

    @torch.no_grad()
    def tokenize(self, ref_wav):
        """ref_wav (B, n) -> global tokens (B, 1, N) int32, exactly bicodec.py:163 + :167 (ref_wav as (B, 1, n))."""
        mel = self.mel_transformer(ref_wav.unsqueeze(1)).squeeze(1)
        return self.speaker_encoder.tokenize(mel.transpose(1, 2))
