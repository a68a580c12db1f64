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
model_repo/vocoder/1/model.py:91-92).  Recurring shapes -- the (n_streams, 50) first-chunk round and the other
sizes of the deterministic chunk schedule -- are replayed from a CUDA graph (one graph launch instead of ~80
kernel launches).  Every graph owns its workspace and output buffers (``bicodec._GraphedCall``), the cache is a
bounded LRU, and one-off shapes (end-of-stream flushes of arbitrary length) run eagerly.
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


def schedule_sizes(policy: ChunkPolicy) -> List[int]:
    """The chunk lengths the growth rule can produce (50, 400, 1500 with the reference's run.sh values)."""
    out, size = [], policy.first_chunk
    while size not in out:
        out.append(size)
        size = min(policy.max_chunk, int(size * policy.scale))
    return out


class StreamingDetokenizer:
    def __init__(self, model, policy: Optional[ChunkPolicy] = None, use_graphs: bool = True,
                 graph_min_batch: int = 8, graph_cache_size: int = 6, graph_after: int = 3,
                 graph_max_frames: int = 32768):
        """``use_graphs``: shapes (B >= graph_min_batch, B*T <= graph_max_frames) are captured when T is one of the
        schedule's chunk sizes, or once the same (B, T) has been seen ``graph_after`` times; at most
        ``graph_cache_size`` graphs (each with a private workspace + pinned output) are kept, least recently used
        evicted first."""
        self.model = model
        self.policy = policy or ChunkPolicy(frame_rate=model.cfg.frame_rate)
        self.use_graphs = use_graphs
        self.graph_min_batch = graph_min_batch
        self.graph_cache_size = graph_cache_size
        self.graph_after = graph_after
        self.graph_max_frames = graph_max_frames
        self.streams: Dict[int, _Stream] = {}
        self._graphs: "Dict[Tuple[int, int, int], object]" = {}     # insertion order = recency (LRU at the front)
        self._seen: Dict[Tuple[int, int], int] = {}
        self._sched = set(schedule_sizes(self.policy))

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

    def _graph_for(self, B: int, T: int):
        """The captured graph of shape (B, T) if the policy wants one, else None (eager)."""
        from .bicodec import _GraphedCall
        model = self.model
        if not self.use_graphs or B < self.graph_min_batch or B * T > self.graph_max_frames:
            return None
        key = (B, T, model._prec(None))
        g = self._graphs.pop(key, None)
        if g is not None and g.generation != model._generation:
            g = None                                   # captured against a handle that no longer exists
            self._graphs = {k: v for k, v in self._graphs.items() if v.generation == model._generation}
        if g is None:
            n = self._seen[(B, T)] = self._seen.get((B, T), 0) + 1
            if len(self._seen) > 4096:
                self._seen.clear()
            if T not in self._sched and n < self.graph_after:
                return None
            while len(self._graphs) >= max(self.graph_cache_size, 1):
                self._graphs.pop(next(iter(self._graphs)))          # least recently used
            g = _GraphedCall(model, B, T, key[2], pinned_out=True)
        self._graphs[key] = g                                        # most recently used at the back
        return g

    def _finish(self, validate: bool) -> None:
        """The round's one stream synchronisation.  With token validation on (``model.validate_tokens``, the default)
        the 16-byte error-flag read rides on it, so an out-of-range id raises IndexError here as it does in
        ``BiCodec.detokenize`` and in the reference's CPU path."""
        if validate:
            self.model.check_tokens()
        else:
            torch.cuda.current_stream(self.model.device).synchronize()

    def decode_batch(self, sem: torch.Tensor, glob: torch.Tensor) -> torch.Tensor:
        """(B,T) host/device tokens + (B,N) globals -> pinned host waveform (B, hop*T); one sync.  When the shape is
        replayed from a graph the result is the graph's own pinned buffer: it is overwritten by the next round of the
        same shape, so callers that keep it must copy it (``poll`` does)."""
        model, dev = self.model, self.model.device
        B, T = sem.shape
        check = bool(model.validate_tokens)
        g = self._graph_for(B, T)
        if g is not None:
            g.replay(sem.to(dev, non_blocking=True), glob.to(dev, non_blocking=True))
            self._finish(check)
            return g.host
        model.validate_tokens = False          # (no separate sync inside detokenize: _finish does both)
        try:
            wav = model.detokenize(sem.to(dev, non_blocking=True), glob.to(dev, non_blocking=True).unsqueeze(1))
        finally:
            model.validate_tokens = check
        host = torch.empty((B, T * model.hop), dtype=torch.float32, pin_memory=True)
        host.copy_(wav.view(B, -1), non_blocking=True)
        self._finish(check)
        return host
