"""Chunked streaming detokenize (BASELINE config 4).

Mirrors the chunk semantics of the reference's Triton orchestrator
(/root/reference runtime/triton_trtllm/model_repo/spark_tts/1/model.py:347-385, values from run.sh:53-56):
tokens accumulate per stream; when ``chunk_size`` tokens are available the chunk is decoded
INDEPENDENTLY from scratch (zero-padded edges, same global tokens), the last ``overlap`` tokens are
kept for the next chunk and ``chunk_size`` grows by ``scale`` up to ``max_chunk``; the remainder is
flushed at end of stream.  ``cross_fade`` is the client's reconstruction
(runtime/triton_trtllm/client_grpc.py:390-415): linear fade over ``overlap`` tokens' worth of samples.

Many streams are served together: chunks of equal length that are ready at the same time are decoded as
one batch (the reference's dynamic batcher can only ``torch.cat`` equal-length requests,
model_repo/vocoder/1/model.py:91-92).  Fixed shapes -- the common (n_streams, 50) first-chunk round --
are replayed from a CUDA graph: the ~90 kernel launches of a pass cost one graph launch.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


@dataclass
class ChunkPolicy:
    frame_rate: float = 50.0
    chunk_duration: float = 1.0         # run.sh: audio_chunk_duration
    max_chunk_duration: float = 30.0    # max_audio_chunk_duration
    scale: float = 8.0                  # audio_chunk_size_scale_factor
    overlap_duration: float = 0.1       # audio_chunk_overlap_duration

    @property
    def first_chunk(self) -> int:
        return math.ceil(self.chunk_duration * self.frame_rate)

    @property
    def max_chunk(self) -> int:
        return math.ceil(self.max_chunk_duration * self.frame_rate)

    @property
    def overlap(self) -> int:
        return math.ceil(self.overlap_duration * self.frame_rate)


@dataclass
class _Stream:
    global_tokens: torch.Tensor                 # (token_num,)
    pending: List[int] = field(default_factory=list)
    chunk_size: int = 0
    closed: bool = False


def chunk_schedule(n_tokens: int, policy: ChunkPolicy) -> List[Tuple[int, int]]:
    """[begin, end) token ranges the reference decodes for a stream of ``n_tokens`` (spark_tts/1/model.py:351-385)."""
    out, start, size = [], 0, policy.first_chunk
    while n_tokens - start >= size:
        out.append((start, start + size))
        start += size - policy.overlap
        size = min(policy.max_chunk, int(size * policy.scale))
    if n_tokens - start > 0:
        out.append((start, n_tokens))
    return out


def cross_fade(chunks: List[np.ndarray], overlap_samples: int) -> np.ndarray:
    """The client's reconstruction (client_grpc.py:390-415)."""
    if not chunks:
        return np.zeros(0, dtype=np.float32)
    if len(chunks) == 1:
        return chunks[0]
    fade_out = np.linspace(1, 0, overlap_samples)
    fade_in = np.linspace(0, 1, overlap_samples)
    out = chunks[0][:-overlap_samples]
    for i in range(1, len(chunks)):
        mix = chunks[i][:overlap_samples] * fade_in + chunks[i - 1][-overlap_samples:] * fade_out
        out = np.concatenate([out, mix, chunks[i][overlap_samples:-overlap_samples]])
    return np.concatenate([out, chunks[-1][-overlap_samples:]]).astype(np.float32)


class _GraphedShape:
    """One CUDA graph per (batch, frames): static token buffers in, static waveform + pinned copy out."""

    def __init__(self, model, batch: int, frames: int):
        dev = model.device
        cfg = model.cfg
        self.sem = torch.zeros((batch, frames), dtype=torch.int64, device=dev)
        self.glob = torch.zeros((batch, 1, cfg.token_num), dtype=torch.int32, device=dev)
        self.host = torch.empty((batch, frames * cfg.hop), dtype=torch.float32, pin_memory=True)
        check, model.validate_tokens = model.validate_tokens, False
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):      # warm-up outside capture (workspace, function attributes)
                for _ in range(2):
                    model.detokenize(self.sem, self.glob)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.wav = model.detokenize(self.sem, self.glob)
                self.host.copy_(self.wav.view(batch, -1), non_blocking=True)
        finally:
            model.validate_tokens = check

    def run(self, sem: torch.Tensor, glob: torch.Tensor) -> torch.Tensor:
        self.sem.copy_(sem, non_blocking=True)
        self.glob.copy_(glob.view(self.glob.shape), non_blocking=True)
        self.graph.replay()
        return self.host


class StreamingDetokenizer:
    def __init__(self, model, policy: Optional[ChunkPolicy] = None, use_graphs: bool = True,
                 graph_min_batch: int = 8):
        self.model = model
        self.policy = policy or ChunkPolicy(frame_rate=model.cfg.frame_rate)
        self.use_graphs = use_graphs
        self.graph_min_batch = graph_min_batch
        self.streams: Dict[int, _Stream] = {}
        self._graphs: Dict[Tuple[int, int], _GraphedShape] = {}

    # ---- stream management ----
    def open(self, stream_id: int, global_tokens: torch.Tensor) -> None:
        g = global_tokens.reshape(-1).to(torch.int32).cpu()
        if g.numel() != self.model.cfg.token_num:
            raise ValueError("global_tokens must hold token_num ids")
        self.streams[stream_id] = _Stream(global_tokens=g, chunk_size=self.policy.first_chunk)

    def push(self, stream_id: int, semantic_tokens) -> None:
        self.streams[stream_id].pending.extend(int(t) for t in semantic_tokens)

    def close(self, stream_id: int) -> None:
        self.streams[stream_id].closed = True

    # ---- decode everything that is ready; returns {stream_id: [waveform chunk np.float32, ...]} ----
    def poll(self) -> Dict[int, List[np.ndarray]]:
        out: Dict[int, List[np.ndarray]] = {}
        while True:
            ready: Dict[int, List[Tuple[int, List[int]]]] = {}
            for sid, st in self.streams.items():
                if len(st.pending) >= st.chunk_size:
                    toks = st.pending[:st.chunk_size]
                    st.pending = st.pending[st.chunk_size - self.policy.overlap:]
                    st.chunk_size = min(self.policy.max_chunk, int(st.chunk_size * self.policy.scale))
                    ready.setdefault(len(toks), []).append((sid, toks))
                elif st.closed and st.pending:
                    toks, st.pending = st.pending, []
                    ready.setdefault(len(toks), []).append((sid, toks))
            if not ready:
                break
            for frames, items in sorted(ready.items()):
                sem = torch.tensor([t for _, t in items], dtype=torch.int64)
                glob = torch.stack([self.streams[sid].global_tokens for sid, _ in items])
                wav = self.decode_batch(sem, glob)
                for i, (sid, _) in enumerate(items):
                    out.setdefault(sid, []).append(wav[i].numpy().copy())
        for sid in [s for s, st in self.streams.items() if st.closed and not st.pending]:
            del self.streams[sid]
        return out

    def decode_batch(self, sem: torch.Tensor, glob: torch.Tensor) -> torch.Tensor:
        """(B,T) host/device tokens + (B,N) globals -> pinned host waveform (B, hop*T); one sync."""
        model, dev = self.model, self.model.device
        B, T = sem.shape
        key = (B, T)
        if self.use_graphs and B >= self.graph_min_batch:
            if key not in self._graphs:
                self._graphs[key] = _GraphedShape(model, B, T)
            host = self._graphs[key].run(sem.to(dev, non_blocking=True), glob.to(dev, non_blocking=True))
            torch.cuda.current_stream(dev).synchronize()
            return host
        check, model.validate_tokens = model.validate_tokens, False
        try:
            wav = model.detokenize(sem.to(dev, non_blocking=True), glob.to(dev, non_blocking=True).unsqueeze(1))
        finally:
            model.validate_tokens = check
        host = torch.empty((B, T * model.hop), dtype=torch.float32, pin_memory=True)
        host.copy_(wav.view(B, -1), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return host
