"""``BiCodec`` -- host-side mirror of the reference class for the detokenize path.

Same surface as /root/reference sparktts/models/bicodec.py (``load_from_checkpoint`` :69-111,
``detokenize`` :171-189, ``remove_weight_norm`` :213-221, plus the harmless ``.to()``/``.eval()``
the callers use), but every stage runs in libsparkcodec.so (hand-written sm_100a kernels).
PyTorch only provides device memory, the current stream and the output tensor.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .config import BiCodecConfig, load_bicodec_yaml

_TOKEN_DTYPES = {torch.int32: _lib.I32, torch.int64: _lib.I64}


class _GraphedCall:
    """One captured CUDA graph of ``sparkcodec_detokenize`` for a fixed (batch, frames, precision): static token
    buffers in, static waveform out, its OWN workspace (a captured pointer must never be reallocated; the model's
    shared ``_ws`` is dropped and re-allocated whenever a later call needs more space).  With ``pinned_out`` the
    device->host copy of the waveform into a pinned buffer is part of the graph (the streaming server's chunk
    round).  ``generation`` is the model's handle generation at capture time: a graph captured against a handle
    that has since been destroyed (``.to(other_device)``) points at freed weights and must not be replayed."""

    def __init__(self, model: "BiCodec", batch: int, frames: int, prec: int, pinned_out: bool = False):
        dev = model._device
        lib = _lib.load()
        self.generation = model._generation
        self.sem = torch.zeros((batch, frames), dtype=torch.int64, device=dev)
        self.glob = torch.zeros((batch, model.cfg.token_num), dtype=torch.int32, device=dev)
        self.wav = torch.empty((batch, 1, model.hop * frames), dtype=torch.float32, device=dev)
        self.host = (torch.empty((batch, frames * model.hop), dtype=torch.float32, pin_memory=True)
                     if pinned_out else None)
        need = C.c_size_t()
        _lib.check(lib.sparkcodec_workspace_bytes(model._handle, batch, frames, C.byref(need)))
        self.ws = torch.empty(need.value, dtype=torch.uint8, device=dev)

        def call():
            _lib.check(lib.sparkcodec_detokenize(
                model._handle, C.c_void_p(self.sem.data_ptr()), _lib.I64, C.c_void_p(self.glob.data_ptr()), _lib.I32,
                batch, frames, prec, C.c_void_p(self.ws.data_ptr()), self.ws.numel(),
                C.c_void_p(self.wav.data_ptr()), model._stream()))

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up outside capture (one-time function attributes)
            call()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            call()
            if self.host is not None:
                self.host.copy_(self.wav.view(batch, -1), non_blocking=True)

    def replay(self, sem: torch.Tensor, glob: torch.Tensor) -> None:
        self.sem.copy_(sem, non_blocking=True)
        self.glob.copy_(glob.reshape(self.glob.shape), non_blocking=True)
        self.graph.replay()

    def run(self, sem: torch.Tensor, glob: torch.Tensor) -> torch.Tensor:
        self.replay(sem, glob)
        return self.wav.clone()                # the caller owns its waveform, like in the reference


class BiCodec:
    """B200-native BiCodec vocoder (semantic + global tokens -> waveform)."""

    def __init__(self, cfg: BiCodecConfig, state_dict: Dict[str, torch.Tensor],
                 device: Optional[torch.device] = None, precision: str = "fp32",
                 workspace_limit_bytes: int = 24 << 30):
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.cfg = cfg
        self.precision = precision
        self.validate_tokens = True          # sync + IndexError on out-of-range ids, like the CPU reference
        self.workspace_limit_bytes = int(workspace_limit_bytes)
        self._sd = state_dict                # kept on the host so .to(other_device) can rebuild
        self._handle: Optional[C.c_void_p] = None
        self._device: Optional[torch.device] = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_stream = None               # the stream whose work last used the shared workspace
        self._impl = "tc"
        # Opt-in CUDA-graph replay for small calls (the CLI's one utterance per call is launch-latency bound: ~80
        # kernels for a few ms of work).  Calls of at most graph_max_frames token frames are captured once per
        # (batch, frames, precision) and replayed; at most graph_cache_size shapes are kept.
        self.use_graphs = False
        self.graph_max_frames = 4096
        self.graph_cache_size = 4
        self._graphs: "Dict[Tuple[int, int, int], _GraphedCall]" = {}
        # bumped whenever the native handle is (re)built or destroyed: holders of captured graphs
        # (StreamingDetokenizer) compare it with _GraphedCall.generation before every replay
        self._generation = 0
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------ construction (bicodec.py:69-111)
    @classmethod
    def load_from_checkpoint(cls, model_dir, device=None, **kwargs) -> "BiCodec":
        """``model_dir`` holds ``config.yaml`` + ``model.safetensors`` exactly as the reference expects."""
        from safetensors.torch import load_file

        cfg = load_bicodec_yaml(os.path.join(str(model_dir), "config.yaml"))
        sd = load_file(os.path.join(str(model_dir), "model.safetensors"))
        return cls(cfg, sd, device=device, **kwargs)

    @classmethod
    def from_state_dict(cls, cfg: BiCodecConfig, state_dict, device=None, **kwargs) -> "BiCodec":
        return cls(cfg, state_dict, device=device, **kwargs)

    def _build(self, device: torch.device) -> None:
        if not torch.cuda.is_available():
            raise _lib.SparkCodecError(
                "BiCodec detokenize runs only on a CUDA device (sm_100a); there is no CPU fallback")
        lib = _lib.load()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        device = torch.device("cuda", idx)
        h = C.c_void_p()
        ccfg = _lib.make_config(self.cfg)
        _lib.check(lib.sparkcodec_create(C.byref(ccfg), idx, C.byref(h)))
        try:
            for key, t in self._sd.items():
                t = t.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(lib.sparkcodec_set_tensor(h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
            c = self.cfg
            _lib.check(lib.sparkcodec_set_mel_params(
                h, c.sample_rate, c.mel_n_fft, c.mel_win_length, c.mel_hop_length, c.num_mels, float(c.mel_fmin),
                -1.0 if c.mel_fmax is None else float(c.mel_fmax)))
            _lib.check(lib.sparkcodec_finalize(h))
        except Exception:
            lib.sparkcodec_destroy(h)
            raise
        self._free()
        self._graphs = {}
        self._generation += 1
        self._handle, self._device, self._ws, self._ws_stream = h, device, None, None
        self.set_impl(self._impl)

    def _free(self) -> None:
        self._graphs = {}
        self._generation += 1
        if self._handle is not None:
            _lib.load().sparkcodec_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # ------------------------------------------------------------------ nn.Module-like no-ops
    def to(self, device) -> "BiCodec":
        device = torch.device(device) if device is not None else None
        if device is None:
            return self
        if device.type != "cuda":
            raise _lib.SparkCodecError(f"BiCodec runs only on CUDA devices, got {device}")
        if self._device is None or (device.index is not None and device.index != self._device.index):
            self._build(device)
        return self

    def eval(self) -> "BiCodec":
        return self

    def remove_weight_norm(self) -> None:
        """weight-norm is folded once inside sparkcodec_finalize (bicodec.py:213-221)."""

    @property
    def device(self) -> Optional[torch.device]:
        return self._device

    @property
    def hop(self) -> int:
        return self.cfg.hop

    # ------------------------------------------------------------------ helpers
    def _ensure(self, *tensors: torch.Tensor) -> None:
        dev = None
        for t in tensors:
            if not isinstance(t, torch.Tensor):
                raise TypeError("tokens must be torch tensors")
            if t.device.type != "cuda":
                raise RuntimeError(
                    f"expected tokens on a CUDA device, got {t.device} (the detokenize path has no CPU fallback)")
            dev = dev or t.device
            if t.device != dev:
                raise RuntimeError("semantic_tokens and global_tokens are on different devices")
        if self._device is None:
            self._build(dev)
        elif dev != self._device:
            raise RuntimeError(f"model is on {self._device} but tokens are on {dev}")

    def _workspace(self, batch: int, frames: int) -> Tuple[int, int]:
        """The shared scratch of this model, grown on demand.  Every pass of one model uses it, so passes issued on
        DIFFERENT streams must not overlap: when the current stream is not the one that used it last, it first waits
        for the work already queued on that stream (nothing happens on the usual single-stream path)."""
        cur = torch.cuda.current_stream(self._device)
        prev = self._ws_stream
        if prev is not None and prev != cur and not torch.cuda.is_current_stream_capturing():
            cur.wait_stream(prev)
        self._ws_stream = cur
        need = C.c_size_t()
        _lib.check(_lib.load().sparkcodec_workspace_bytes(self._handle, batch, frames, C.byref(need)))
        want = min(need.value, max(self.workspace_limit_bytes, 1))
        if want < need.value:   # the library splits the batch; make sure one utterance fits
            one = C.c_size_t()
            _lib.check(_lib.load().sparkcodec_workspace_bytes(self._handle, 1, frames, C.byref(one)))
            want = max(want, one.value)
        if self._ws is None or self._ws.numel() < want:
            self._ws = None
            self._ws = torch.empty(want, dtype=torch.uint8, device=self._device)
        return self._ws.data_ptr(), self._ws.numel()

    def _tokens(self, semantic_tokens, global_tokens):
        if semantic_tokens.dim() != 2:
            raise ValueError(f"semantic_tokens must be (B, T), got {tuple(semantic_tokens.shape)}")
        B, T = semantic_tokens.shape
        N = self.cfg.token_num
        if global_tokens.dim() == 3 and tuple(global_tokens.shape) == (B, 1, N):
            pass
        elif global_tokens.dim() == 2 and tuple(global_tokens.shape) == (B, N):
            pass
        else:
            raise ValueError(f"global_tokens must be ({B}, 1, {N}) or ({B}, {N}), got {tuple(global_tokens.shape)}")
        for name, t in (("semantic_tokens", semantic_tokens), ("global_tokens", global_tokens)):
            if t.dtype not in _TOKEN_DTYPES:
                raise ValueError(f"{name} must be int32 or int64, got {t.dtype}")
        return semantic_tokens.contiguous(), global_tokens.contiguous(), B, T

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _prec(self, precision: Optional[str]) -> int:
        return _lib.PRECISIONS[precision or self.precision]

    def check_tokens(self) -> None:
        """Raises IndexError if any token id since the last check was out of range (synchronises)."""
        _lib.check(_lib.load().sparkcodec_check_tokens(self._handle, self._stream()))

    # ------------------------------------------------------------------ the hot path (bicodec.py:171-189)
    @torch.no_grad()
    def detokenize(self, semantic_tokens: torch.Tensor, global_tokens: torch.Tensor,
                   precision: Optional[str] = None) -> torch.Tensor:
        """semantic (B,T) int, global (B,1,32) int -> waveform (B,1,hop*T) float32 on the input device."""
        self._ensure(semantic_tokens, global_tokens)
        sem, glob, B, T = self._tokens(semantic_tokens, global_tokens)
        wav = torch.empty((B, 1, self.hop * T), dtype=torch.float32, device=self._device)
        if B == 0 or T == 0:
            return wav
        if self.use_graphs and B * T <= self.graph_max_frames and self._impl == "tc":
            key = (B, T, self._prec(precision))
            g = self._graphs.pop(key, None)
            if g is None:
                while len(self._graphs) >= self.graph_cache_size:
                    self._graphs.pop(next(iter(self._graphs)))      # least recently used
                g = _GraphedCall(self, B, T, key[2])
            self._graphs[key] = g
            wav = g.run(sem, glob)
            if self.validate_tokens:
                self.check_tokens()
            return wav
        ws, ws_bytes = self._workspace(B, T)
        _lib.check(_lib.load().sparkcodec_detokenize(
            self._handle, C.c_void_p(sem.data_ptr()), _TOKEN_DTYPES[sem.dtype], C.c_void_p(glob.data_ptr()),
            _TOKEN_DTYPES[glob.dtype], B, T, self._prec(precision), C.c_void_p(ws), ws_bytes,
            C.c_void_p(wav.data_ptr()), self._stream()))
        if self.validate_tokens:
            self.check_tokens()
        return wav

    __call__ = detokenize

    @torch.no_grad()
    def detokenize_ragged(self, semantic_list, global_tokens: torch.Tensor, precision: Optional[str] = None):
        """Utterances of DIFFERENT lengths in one call: ``semantic_list[b]`` is (T_b,) or (1, T_b), ``global_tokens``
        (B, N) or (B, 1, N).  Returns a list of (hop * T_b,) float32 tensors.

        The reference's Triton vocoder can only batch equal lengths (``torch.cat`` of the requests,
        runtime/triton_trtllm/model_repo/vocoder/1/model.py:72-106).  Here utterances are bucketed by length
        and every bucket is one batched pass; nothing is padded, because padding frames would leak into the
        last 67 real frames through the receptive field.  Every waveform is bit-identical to decoding that
        utterance alone (the kernels are batch-invariant)."""
        B = len(semantic_list)
        if global_tokens.shape[0] != B:
            raise ValueError("one row of global tokens per utterance")
        glob = global_tokens.reshape(B, -1)
        rows = [t.reshape(-1) for t in semantic_list]
        buckets: Dict[int, list] = {}
        for i, t in enumerate(rows):
            buckets.setdefault(int(t.numel()), []).append(i)
        out: list = [None] * B
        for T, idx in sorted(buckets.items()):
            if T == 0:
                for i in idx:
                    out[i] = torch.empty((0,), dtype=torch.float32, device=glob.device)
                continue
            sem = torch.stack([rows[i] for i in idx], dim=0)
            sel = torch.as_tensor(idx, device=glob.device)
            wav = self.detokenize(sem, glob.index_select(0, sel).unsqueeze(1).contiguous(), precision)
            for j, i in enumerate(idx):
                out[i] = wav[j, 0]
        return out

    # ------------------------------------------------------------------ halves, for time-sharding
    @torch.no_grad()
    def prenet(self, semantic_tokens, global_tokens, precision: Optional[str] = None) -> torch.Tensor:
        """-> x = prenet(z_q, d) + d as (B, T, d_model) fp32 channels-last."""
        self._ensure(semantic_tokens, global_tokens)
        sem, glob, B, T = self._tokens(semantic_tokens, global_tokens)
        x = torch.empty((B, T, self.cfg.d_model), dtype=torch.float32, device=self._device)
        if B == 0 or T == 0:
            return x
        ws, ws_bytes = self._workspace(B, T)
        _lib.check(_lib.load().sparkcodec_prenet(
            self._handle, C.c_void_p(sem.data_ptr()), _TOKEN_DTYPES[sem.dtype], C.c_void_p(glob.data_ptr()),
            _TOKEN_DTYPES[glob.dtype], B, T, self._prec(precision), C.c_void_p(ws), ws_bytes,
            C.c_void_p(x.data_ptr()), self._stream()))
        if self.validate_tokens:
            self.check_tokens()
        return x

    @torch.no_grad()
    def wavegen(self, x: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
        """x (B, T, d_model) fp32 channels-last -> waveform (B, 1, hop*T)."""
        self._ensure(x)
        if x.dim() != 3 or x.shape[2] != self.cfg.d_model or x.dtype != torch.float32:
            raise ValueError(f"x must be float32 (B, T, {self.cfg.d_model})")
        x = x.contiguous()
        B, T, _ = x.shape
        wav = torch.empty((B, 1, self.hop * T), dtype=torch.float32, device=self._device)
        if B == 0 or T == 0:
            return wav
        ws, ws_bytes = self._workspace(B, T)
        _lib.check(_lib.load().sparkcodec_wavegen(
            self._handle, C.c_void_p(x.data_ptr()), B, T, self._prec(precision), C.c_void_p(ws), ws_bytes,
            C.c_void_p(wav.data_ptr()), self._stream()))
        return wav

    # staged WaveGenerator input: lets a time-sharded rank overlap its halo exchange with the interior rows
    def can_stage(self, batch: int, frames_total: int) -> bool:
        """True when one workspace (under ``workspace_limit_bytes``) holds the whole (batch, frames_total) window."""
        if self._handle is None or batch <= 0 or frames_total <= 0:
            return False
        need = C.c_size_t()
        _lib.check(_lib.load().sparkcodec_workspace_bytes(self._handle, batch, frames_total, C.byref(need)))
        return need.value <= max(self.workspace_limit_bytes, 1)

    @torch.no_grad()
    def wavegen_stage(self, x_rows: torch.Tensor, frames_total: int, row_offset: int,
                      precision: Optional[str] = None) -> None:
        """Stage rows [row_offset, row_offset + rows) of a (B, frames_total, d_model) WaveGenerator input.  ``x_rows``
        is (B, rows, d_model) fp32 and may be a time slice of a wider tensor (no copy is made)."""
        self._ensure(x_rows)
        D = self.cfg.d_model
        if x_rows.dim() != 3 or x_rows.shape[2] != D or x_rows.dtype != torch.float32:
            raise ValueError(f"x_rows must be float32 (B, rows, {D})")
        B, rows, _ = x_rows.shape
        if rows == 0:
            return
        if x_rows.stride(2) != 1 or x_rows.stride(1) != D or (B > 1 and x_rows.stride(0) < rows * D):
            x_rows = x_rows.contiguous()
        ws, ws_bytes = self._workspace(B, frames_total)
        _lib.check(_lib.load().sparkcodec_wavegen_stage(
            self._handle, C.c_void_p(x_rows.data_ptr()), B, rows, x_rows.stride(0) if B > 1 else rows * D,
            frames_total, row_offset, self._prec(precision), C.c_void_p(ws), ws_bytes, self._stream()))

    @torch.no_grad()
    def wavegen_staged(self, batch: int, frames_total: int, precision: Optional[str] = None) -> torch.Tensor:
        """Runs the WaveGenerator over the rows staged by ``wavegen_stage`` -> waveform (B, 1, hop*frames_total)."""
        wav = torch.empty((batch, 1, self.hop * frames_total), dtype=torch.float32, device=self._device)
        ws, ws_bytes = self._workspace(batch, frames_total)
        _lib.check(_lib.load().sparkcodec_wavegen_staged(
            self._handle, batch, frames_total, self._prec(precision), C.c_void_p(ws), ws_bytes,
            C.c_void_p(wav.data_ptr()), self._stream()))
        return wav

    # ------------------------------------------------------------------ encode side, semantic half
    @torch.no_grad()
    def tokenize_semantic(self, feat: torch.Tensor, precision: Optional[str] = None, return_margin: bool = False):
        """The semantic half of the reference's ``BiCodec.tokenize`` (bicodec.py:151-169):
        ``quantizer.tokenize(encoder(feat.transpose(1, 2)))``.  feat (B, T, d_model) fp32 (the wav2vec2 feature
        mix) -> semantic tokens (B, T) int64.  Needs a checkpoint that carries ``encoder.*`` and
        ``quantizer.in_project.*``.  With ``return_margin`` also returns the (B, T) fp32 gap between the best and
        the second-best code distance (a near-zero gap marks a numerical tie).  (Speaker half: ``tokenize_speaker``.)
        A model in "fp32" mode runs this side as "fp32x3" (three bf16 products per MAC, ~2^-17 relative) unless the
        caller names a precision: the output is a discrete index, and the encoder runs once per prompt."""
        if precision is None and self.precision == "fp32":
            precision = "fp32x3"
        self._ensure(feat)
        if feat.dim() != 3 or feat.shape[2] != self.cfg.d_model or feat.dtype != torch.float32:
            raise ValueError(f"feat must be float32 (B, T, {self.cfg.d_model})")
        feat = feat.contiguous()
        B, T, _ = feat.shape
        tokens = torch.empty((B, T), dtype=torch.int64, device=self._device)
        margin = torch.empty((B, T), dtype=torch.float32, device=self._device) if return_margin else None
        if B and T:
            ws, ws_bytes = self._workspace(B, T)
            _lib.check(_lib.load().sparkcodec_tokenize_semantic(
                self._handle, C.c_void_p(feat.data_ptr()), B, T, self._prec(precision), C.c_void_p(ws), ws_bytes,
                C.c_void_p(tokens.data_ptr()), C.c_void_p(margin.data_ptr()) if return_margin else None,
                self._stream()))
        return (tokens, margin) if return_margin else tokens

    # ------------------------------------------------------------------ encode side, speaker half
    @torch.no_grad()
    def tokenize_speaker(self, ref_wav: torch.Tensor, return_margin: bool = False, tap: Optional[str] = None):
        """The speaker half of the reference's ``BiCodec.tokenize`` (bicodec.py:162-167):
        ``speaker_encoder.tokenize(mel_transformer(ref_wav).squeeze(1).transpose(1, 2))``.  ref_wav (B, n) or (B, 1, n)
        fp32 on the model's device -> global tokens (B, 1, token_num) int32.  Needs a checkpoint that carries the
        ECAPA-TDNN / perceiver / FSQ ``project_in`` tensors.  ``return_margin``: also the (B, token_num) distance of the
        closest FSQ coordinate to a rounding boundary.  ``tap`` (tests): also an intermediate as (B, rows, channels)."""
        self._ensure(ref_wav)
        if ref_wav.dim() == 3 and ref_wav.shape[1] == 1:
            ref_wav = ref_wav[:, 0]
        if ref_wav.dim() != 2 or ref_wav.dtype != torch.float32:
            raise ValueError("ref_wav must be float32 (B, n) or (B, 1, n)")
        ref_wav = ref_wav.contiguous()
        B, n = ref_wav.shape
        N = self.cfg.token_num
        tokens = torch.empty((B, 1, N), dtype=torch.int32, device=self._device)
        margin = torch.empty((B, N), dtype=torch.float32, device=self._device) if return_margin else None
        lib = _lib.load()
        tapped = None
        if B:
            need = C.c_size_t()
            _lib.check(lib.sparkcodec_speaker_workspace_bytes(self._handle, B, n, C.byref(need)))
            ws = torch.empty(need.value, dtype=torch.uint8, device=self._device)
            if tap is None:
                _lib.check(lib.sparkcodec_tokenize_speaker(
                    self._handle, C.c_void_p(ref_wav.data_ptr()), B, n, C.c_void_p(ws.data_ptr()), ws.numel(),
                    C.c_void_p(tokens.data_ptr()), C.c_void_p(margin.data_ptr()) if return_margin else None,
                    self._stream()))
            else:
                cap = B * (n // self.cfg.mel_hop_length + 1) * 1536 + 4096
                out = torch.empty(cap, dtype=torch.float32, device=self._device)
                shape = (C.c_int64 * 2)()
                _lib.check(lib.sparkcodec_tokenize_speaker_tap(
                    self._handle, C.c_void_p(ref_wav.data_ptr()), B, n, C.c_void_p(ws.data_ptr()), ws.numel(),
                    C.c_void_p(tokens.data_ptr()), tap.encode(), C.c_void_p(out.data_ptr()), cap, shape, self._stream()))
                rows, ch = int(shape[0]), int(shape[1])
                tapped = out[: B * rows * ch].view(B, rows, ch).clone()
            torch.cuda.current_stream(self._device).synchronize()      # `ws` is freed on return
        res = (tokens,)
        if return_margin:
            res += (margin,)
        if tap is not None:
            res += (tapped,)
        return res[0] if len(res) == 1 else res

    @torch.no_grad()
    def tokenize(self, batch: Dict[str, torch.Tensor], precision: Optional[str] = None):
        """``BiCodec.tokenize`` of the reference (bicodec.py:151-169): ``batch["feat"]`` (B, T, d_model) wav2vec2
        feature mix and ``batch["ref_wav"]`` (B, n) reference clip -> (semantic_tokens (B, T) int64,
        global_tokens (B, 1, token_num) int32), both on the model's device."""
        dev = self._device if self._device is not None else torch.device("cuda")
        feat = batch["feat"].to(dev, torch.float32)
        ref_wav = batch["ref_wav"].to(dev, torch.float32)
        semantic_tokens = self.tokenize_semantic(feat, precision)
        global_tokens = self.tokenize_speaker(ref_wav)
        return semantic_tokens, global_tokens

    def halo_frames(self) -> Tuple[int, int]:
        if self._handle is None:
            raise _lib.SparkCodecError("model is not on a device yet")
        a, b = C.c_int(), C.c_int()
        _lib.check(_lib.load().sparkcodec_halo_frames(self._handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ------------------------------------------------------------------ test / profiling hooks
    def set_impl(self, impl: str) -> None:
        """'tc' (tcgen05, the product path) or 'simt' (CUDA-core verification kernels, tests only)."""
        if impl not in ("tc", "simt", "tc_unfused"):
            raise ValueError("impl must be 'tc', 'simt' or 'tc_unfused'")
        self._impl = impl
        if self._handle is not None:
            _lib.check(_lib.load().sparkcodec_set_impl(
                self._handle, {"tc": _lib.IMPL_TC, "simt": _lib.IMPL_SIMT, "tc_unfused": _lib.IMPL_TC_UNFUSED}[impl]))

    @torch.no_grad()
    def detokenize_tap(self, semantic_tokens, global_tokens, tap: str, precision: Optional[str] = None):
        """Runs detokenize and also returns the named intermediate activation as (B, rows, channels) fp32."""
        self._ensure(semantic_tokens, global_tokens)
        sem, glob, B, T = self._tokens(semantic_tokens, global_tokens)
        wav = torch.empty((B, 1, self.hop * T), dtype=torch.float32, device=self._device)
        ws, ws_bytes = self._workspace(B, T)
        # largest activation a tap can name: a WaveGenerator stage (rows per frame x channels halve / multiply per
        # up-block: (B, 320T, 96) and (B, 160T, 192) with the model-card config), conv-in, or the MLP's hidden layer
        per_frame, rows, ch = 0, 1, self.cfg.dec_channels
        for r in self.cfg.rates:
            rows, ch = rows * r, ch // 2
            per_frame = max(per_frame, rows * ch)
        cap = B * T * max(per_frame, self.cfg.dec_channels, self.cfg.vocos_intermediate_dim, self.cfg.d_model) + 4096
        out = torch.empty(cap, dtype=torch.float32, device=self._device)
        shape = (C.c_int64 * 2)()
        _lib.check(_lib.load().sparkcodec_detokenize_tap(
            self._handle, C.c_void_p(sem.data_ptr()), _TOKEN_DTYPES[sem.dtype], C.c_void_p(glob.data_ptr()),
            _TOKEN_DTYPES[glob.dtype], B, T, self._prec(precision), C.c_void_p(ws), ws_bytes,
            C.c_void_p(wav.data_ptr()), tap.encode(), C.c_void_p(out.data_ptr()), cap, shape, self._stream()))
        rows, ch = int(shape[0]), int(shape[1])
        return wav, out[: B * rows * ch].view(B, rows, ch).clone()

    def launch_count(self) -> int:
        n = C.c_int64()
        _lib.check(_lib.load().sparkcodec_launch_count(self._handle, C.byref(n)))
        return n.value

    def profile(self, enable: bool) -> None:
        """Per-launch CUDA-event timing for the roofline pass of bench.py (never on during a timed run)."""
        _lib.check(_lib.load().sparkcodec_profile(self._handle, 1 if enable else 0))

    def profile_read(self):
        """-> list of dict(name, ms, flops, bytes), one per kernel launch since profile(True)."""
        need = C.c_size_t()
        _lib.check(_lib.load().sparkcodec_profile_read(self._handle, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _lib.check(_lib.load().sparkcodec_profile_read(self._handle, buf, need.value, C.byref(need)))
        rows = []
        for line in buf.value.decode().splitlines():
            name, ms, fl, by = line.split("\t")
            rows.append(dict(name=name, ms=float(ms), flops=float(fl), bytes=float(by)))
        return rows
