"""``BiCodecTokenizer`` -- drop-in for the detokenize half of the reference façade
(/root/reference sparktts/models/audio_tokenizer.py:29-146).

``cli/SparkTTS.py:231-234`` calls ``self.audio_tokenizer.detokenize(global_tokens (1,32), semantic (1,T))``
on ``self.device`` and writes the returned numpy array with soundfile; that call keeps working unchanged.
The tokenize side (wav2vec2 + encoder, run once per prompt) is out of scope of this package and raises.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .bicodec import BiCodec


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

    def tokenize(self, audio_path: str):
        raise NotImplementedError(
            "tokenize (wav2vec2 + BiCodec encoder) is outside the B200 detokenize path; use the reference's "
            "BiCodecTokenizer.tokenize for prompt audio")

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
