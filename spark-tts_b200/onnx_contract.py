"""ONNX ``bicodec_vocoder`` I/O contract (reference export_sparktts_onnx.py:767-867) on top of the native path:
inputs ``semantic_tokens`` (B,T) int64 and ``global_tokens`` (B,1,32) int32, output ``output_waveform``
(B,1,320*T) float32.  ``VocoderSession.run`` follows onnxruntime's ``InferenceSession.run`` calling
convention so code written against the exported model can switch without edits."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .bicodec import BiCodec

INPUT_NAMES = ["semantic_tokens", "global_tokens"]       # export_sparktts_onnx.py:819
OUTPUT_NAMES = ["output_waveform"]                       # export_sparktts_onnx.py:820


class _IO:
    def __init__(self, name, shape, type_):
        self.name, self.shape, self.type = name, shape, type_


class VocoderSession:
    def __init__(self, model: BiCodec, device=None):
        self.model = model.to(device or "cuda")

    def get_inputs(self) -> List[_IO]:
        n = self.model.cfg.token_num
        return [_IO("semantic_tokens", ["batch_size", "sequence_length"], "tensor(int64)"),
                _IO("global_tokens", ["batch_size", 1, n], "tensor(int32)")]

    def get_outputs(self) -> List[_IO]:
        return [_IO("output_waveform", ["batch_size", 1, "audio_length"], "tensor(float)")]

    def run(self, output_names: Optional[Sequence[str]], input_feed: Dict[str, np.ndarray]) -> List[np.ndarray]:
        missing = [k for k in INPUT_NAMES if k not in input_feed]
        if missing:
            raise ValueError(f"missing inputs {missing}")
        if output_names is not None and list(output_names) != OUTPUT_NAMES:
            raise ValueError(f"unknown outputs {list(output_names)}; the graph has {OUTPUT_NAMES}")
        sem = torch.as_tensor(np.ascontiguousarray(input_feed["semantic_tokens"]))
        glob = torch.as_tensor(np.ascontiguousarray(input_feed["global_tokens"]))
        if glob.dim() != 3:
            raise ValueError("global_tokens must have shape (batch, 1, token_num)")
        dev = self.model.device
        wav = self.model.detokenize(sem.to(dev), glob.to(dev))
        return [wav.cpu().numpy()]
