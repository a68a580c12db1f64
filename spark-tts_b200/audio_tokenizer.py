"""``BiCodecTokenizer`` -- drop-in for the detokenize half of the reference façade
(/root/reference sparktts/models/audio_tokenizer.py:29-146).

``cli/SparkTTS.py:231-234`` calls ``self.audio_tokenizer.detokenize(global_tokens (1,32), semantic (1,T))``
on ``self.device`` and writes the returned numpy array with soundfile; that call keeps working unchanged.
The tokenize side (``cli/SparkTTS.py:90-92`` for voice cloning) mirrors the reference too: audio loading and the
reference clip on the host, the wav2vec2 feature mix through HuggingFace transformers exactly as the reference
does (it is a third-party model, not part of BiCodec), and ``BiCodec.tokenize`` -- feature encoder + code search,
mel -> ECAPA-TDNN -> perceiver -> FSQ -- on the B200 kernels.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .bicodec import BiCodec


def audio_volume_normalize(audio: np.ndarray, coeff: float = 0.2) -> np.ndarray:
    """sparktts/utils/audio.py:33-74: scale so that the mean of the loudest 10 % of the non-silent samples is ``coeff``
    (scale clamped to [0.1, 10]), then keep the peak at or below 1."""
    temp = np.sort(np.abs(audio))
    if temp[-1] < 0.1:
        scaling_factor = max(temp[-1], 1e-3)
        audio = audio / scaling_factor * 0.1
    temp = temp[temp > 0.01]
    L = temp.shape[0]
    if L <= 10:
        return audio
    volume = np.mean(temp[int(0.9 * L):int(0.99 * L)])
    audio = audio * np.clip(coeff / volume, a_min=0.1, a_max=10)
    max_value = np.max(np.abs(audio))
    if max_value > 1:
        audio = audio / max_value
    return audio


class BiCodecTokenizer:
    def __init__(self, model_dir=None, device: Optional[torch.device] = None, model: Optional[BiCodec] = None,
                 **kwargs):
        """``model_dir`` is the Spark-TTS-0.5B directory (holds ``BiCodec/``), as in the reference; tests and
        the benchmark pass an already-built ``model`` because no checkpoint exists offline."""
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.model_dir = model_dir
        if model is not None:
            self.model = model.to(self.device)
        else:
            self.model = BiCodec.load_from_checkpoint(os.path.join(str(model_dir), "BiCodec"), **kwargs).to(self.device)
        self._pinned: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ tokenize side (audio_tokenizer.py:57-130)
    def _config(self):
        return self.model.cfg

    def get_ref_clip(self, wav: np.ndarray) -> np.ndarray:
        """Reference clip for the speaker tokens (audio_tokenizer.py:57-71): ref_segment_duration seconds rounded
        down to a multiple of latent_hop_length, the audio tiled when it is shorter."""
        c = self._config()
        n = int(c.sample_rate * c.ref_segment_duration) // c.latent_hop_length * c.latent_hop_length
        if n > len(wav):
            wav = np.tile(wav, n // len(wav) + 1)
        return wav[:n]

    def load_audio(self, wav_path) -> np.ndarray:
        """sparktts/utils/audio.py:77-118 load_audio: mono, resampled to the model rate, optional volume normalisation
        (:33-74).  soundfile / soxr are not dependencies of this package: WAV files are read with the standard
        library, other formats through torchaudio."""
        c = self._config()
        try:
            from scipy.io import wavfile
            sr, audio = wavfile.read(str(wav_path))
            if audio.dtype.kind == "i":
                audio = audio.astype(np.float32) / float(np.iinfo(audio.dtype).max + 1)
            elif audio.dtype.kind == "u":
                audio = (audio.astype(np.float32) - 128.0) / 128.0
            audio = audio.astype(np.float32)
        except Exception:
            import torchaudio
            t, sr = torchaudio.load(str(wav_path))
            audio = t.transpose(0, 1).numpy()
        if audio.ndim > 1:
            audio = audio[:, 0]
        if sr != c.sample_rate:
            import torchaudio.functional as AF
            audio = AF.resample(torch.from_numpy(np.ascontiguousarray(audio)), sr, c.sample_rate).numpy()
        if c.volume_normalize:
            audio = audio_volume_normalize(audio)
        return audio.astype(np.float32)

    def process_audio(self, wav_path):
        """-> (wav (n,), ref_wav (1, n_ref) float tensor), audio_tokenizer.py:73-86."""
        wav = self.load_audio(wav_path)
        return wav, torch.from_numpy(self.get_ref_clip(wav)).unsqueeze(0).float()

    def extract_wav2vec2_features(self, wavs) -> torch.Tensor:
        """The wav2vec2-large-xlsr-53 feature mix the reference feeds BiCodec (audio_tokenizer.py:88-102): mean of hidden
        states 11, 14 and 16.  Third-party model through HuggingFace transformers, loaded from
        ``<model_dir>/wav2vec2-large-xlsr-53`` on first use, exactly like the reference."""
        if getattr(self, "feature_extractor", None) is None:
            if self.model_dir is None:
                raise RuntimeError("wav2vec2 features need model_dir/wav2vec2-large-xlsr-53 (pass `feat` to "
                                   "tokenize_batch / BiCodec.tokenize to supply them yourself)")
            from transformers import Wav2Vec2FeatureExtractor, Wav2Vec2Model
            path = os.path.join(str(self.model_dir), "wav2vec2-large-xlsr-53")
            self.processor = Wav2Vec2FeatureExtractor.from_pretrained(path)
            self.feature_extractor = Wav2Vec2Model.from_pretrained(path).to(self.device)
            self.feature_extractor.config.output_hidden_states = True
        inputs = self.processor(wavs, sampling_rate=16000, return_tensors="pt", padding=True,
                                output_hidden_states=True).input_values
        with torch.no_grad():
            feat = self.feature_extractor(inputs.to(self.feature_extractor.device))
        return (feat.hidden_states[11] + feat.hidden_states[14] + feat.hidden_states[16]) / 3

    def tokenize_batch(self, batch):
        """batch: ``wav`` (list of arrays) and ``ref_wav`` (B, n); ``feat`` may be supplied, else it is extracted.
        -> (global_tokens, semantic_tokens), audio_tokenizer.py:104-117."""
        if "feat" not in batch:
            batch["feat"] = self.extract_wav2vec2_features(batch["wav"])
        semantic_tokens, global_tokens = self.model.tokenize(batch)
        return global_tokens, semantic_tokens

    def tokenize(self, audio_path: str):
        """audio file -> (global_tokens (1, 1, 32) int32, semantic_tokens (1, T) int64) on the device
        (audio_tokenizer.py:119-130)."""
        wav, ref_wav = self.process_audio(audio_path)
        feat = self.extract_wav2vec2_features(wav)
        batch = {"wav": torch.from_numpy(wav).unsqueeze(0).float().to(self.device),
                 "ref_wav": ref_wav.to(self.device), "feat": feat.to(self.device)}
        semantic_tokens, global_tokens = self.model.tokenize(batch)
        return global_tokens, semantic_tokens

    def detokenize(self, global_tokens: torch.Tensor, semantic_tokens: torch.Tensor) -> np.ndarray:
        """global (B,32), semantic (B,T) -> float32 numpy (B, hop*T), squeezed like the reference's
        ``wav_rec.detach().squeeze().cpu().numpy()`` (audio_tokenizer.py:144-146)."""
        global_tokens = global_tokens.unsqueeze(1)
        m = self.model
        check, m.validate_tokens = m.validate_tokens, False
        try:
            wav_rec = m.detokenize(semantic_tokens, global_tokens)
        finally:
            m.validate_tokens = check
        # D2H through a pinned block of torch's caching host allocator (a pageable .cpu() of a 64 x 10 s batch is a
        # staged, blocking copy); the caller gets a fresh array that owns the block, like the reference's result
        host = torch.empty(wav_rec.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(wav_rec, non_blocking=True)
        if check:
            m.check_tokens()                        # one stream sync serves the copy and the id check
        else:
            torch.cuda.current_stream(wav_rec.device).synchronize()
        return host.squeeze().numpy()

    def detokenize_pinned(self, global_tokens: torch.Tensor, semantic_tokens: torch.Tensor) -> torch.Tensor:
        """Same call, but host tokens (pinned or pageable) in, reusable pinned host waveform (B, hop*T) out:
        H2D of the tokens, the kernels and the D2H of the waveform are queued on one stream, one sync at the
        end.  This is the end-to-end path bench.py times as ``e2e``."""
        dev = self.device
        sem = semantic_tokens.to(dev, non_blocking=True)
        glob = global_tokens.to(dev, non_blocking=True).unsqueeze(1)
        check, self.model.validate_tokens = self.model.validate_tokens, False
        try:
            wav = self.model.detokenize(sem, glob)
        finally:
            self.model.validate_tokens = check
        n = wav.numel()
        if self._pinned is None or self._pinned.numel() < n:
            self._pinned = torch.empty(n, dtype=torch.float32, pin_memory=True)
        out = self._pinned[:n].view(wav.shape[0], wav.shape[2])
        out.copy_(wav.view(wav.shape[0], wav.shape[2]), non_blocking=True)
        if check:
            self.model.check_tokens()          # synchronises
        else:
            torch.cuda.current_stream(dev).synchronize()
        return out
