"""Token feed: the step immediately BEFORE the detokenize path.

The reference turns the LLM output into BiCodec codes on the host: ``tokenizer.batch_decode`` to text, then
``re.findall(r"bicodec_semantic_(\\d+)", text)`` (and ``bicodec_global_`` in voice-creation mode) --
/root/reference cli/SparkTTS.py:213-228, runtime/triton_trtllm/model_repo/spark_tts/1/model.py:283-295.

* :func:`codes_from_text` is the host mirror of exactly that (same regexes, same dtypes).
* :func:`codes_from_token_ids` is the B200 path: ``<|bicodec_semantic_N|>`` / ``<|bicodec_global_N|>`` are
  contiguous added-token id ranges of the Spark-TTS tokenizer, so the selection is an order-preserving
  compaction of the generated ids on the device (``sparkcodec_extract_codes``), with no decode, no regex
  and no host round trip; detokenize can start as soon as the ids exist.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import List, Tuple

import torch

from . import _lib

_SEM = re.compile(r"bicodec_semantic_(\d+)")
_GLOB = re.compile(r"bicodec_global_(\d+)")


def codes_from_text(text: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """(semantic (1, T) int64, global (1, 1, G) int64) from decoded LLM text; cli/SparkTTS.py:213-228."""
    sem = torch.tensor([int(t) for t in _SEM.findall(text)], dtype=torch.int64).unsqueeze(0)
    glob = torch.tensor([int(t) for t in _GLOB.findall(text)], dtype=torch.int64).unsqueeze(0).unsqueeze(0)
    return sem, glob


def token_bases(tokenizer) -> Tuple[int, int]:
    """Ids of ``<|bicodec_semantic_0|>`` and ``<|bicodec_global_0|>`` in a Spark-TTS (Qwen2.5) tokenizer."""
    return (int(tokenizer.convert_tokens_to_ids("<|bicodec_semantic_0|>")),
            int(tokenizer.convert_tokens_to_ids("<|bicodec_global_0|>")))


def codes_from_token_ids(token_ids: torch.Tensor, semantic_base: int, global_base: int, codebook_size: int = 8192,
                         global_size: int = 4096, max_global: int = 32):
    """Generated ids (B, N) int32/int64 on a CUDA device -> ``(semantic (B, N) int32, semantic_len (B,) int32,
    global (B, max_global) int32, global_len (B,) int32)``, all on the device, asynchronous on the current
    stream.  Row b of ``semantic`` holds ``semantic_len[b]`` codes in generation order (the rest is untouched)."""
    if token_ids.dim() != 2 or token_ids.dtype not in (torch.int32, torch.int64):
        raise ValueError("token_ids must be (B, N) int32 or int64")
    if token_ids.device.type != "cuda":
        raise RuntimeError("token_ids must live on a CUDA device (there is no CPU fallback; use codes_from_text)")
    ids = token_ids.contiguous()
    B, N = ids.shape
    dev = ids.device
    sem = torch.zeros((B, N), dtype=torch.int32, device=dev)
    glob = torch.zeros((B, max_global), dtype=torch.int32, device=dev)
    sem_len = torch.zeros((B,), dtype=torch.int32, device=dev)
    glob_len = torch.zeros((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sparkcodec_extract_codes(
            C.c_void_p(ids.data_ptr()), _lib.I64 if ids.dtype == torch.int64 else _lib.I32, B, N, int(semantic_base),
            int(codebook_size), int(global_base), int(global_size), C.c_void_p(sem.data_ptr()),
            C.c_void_p(sem_len.data_ptr()), C.c_void_p(glob.data_ptr()), int(max_global),
            C.c_void_p(glob_len.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return sem, sem_len, glob, glob_len


def split_ragged(semantic: torch.Tensor, semantic_len: torch.Tensor) -> List[torch.Tensor]:
    """Padded (B, N) + lengths -> list of (T_b,) tensors (one host sync for the lengths)."""
    lens = semantic_len.tolist()
    return [semantic[b, :n] for b, n in enumerate(lens)]
