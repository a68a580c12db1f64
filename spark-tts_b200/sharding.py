"""Multi-GPU partitioning of the detokenize path (one process per GPU, torch.distributed).

Two modes:

* **utterance sharding** (BASELINE config 3): utterances are independent units -- contiguous blocks of
  B/N utterances per rank, weights replicated, NO collective on the data path.
* **time sharding** (BASELINE config 5): every utterance is cut into N contiguous frame windows.  The
  network is a finite-receptive-field stack, so rank g
    1. runs the prenet on its window widened by ``prenet_halo`` frames (tokens are replicated, so this
       is recompute of 57 cheap 50 Hz frames, no communication) and keeps x for its own frames,
    2. posts the exchange of ``wavegen_halo`` boundary rows of x = prenet(z_q, d) + d with its left / right
       neighbour (NCCL send/recv over NVLink: B x halo x 1024 fp32 per direction) and, while the messages are
       in flight, converts its own (interior) rows into the WaveGenerator's operand format,
    3. stages the received halo rows and runs the WaveGenerator on the widened window, keeping the samples of
       its own frames.
  Zero padding applies at true utterance edges only; rows inside the halo of a shard edge are
  discarded, so the result equals the un-sharded one.

The compute backend is any object with ``prenet(sem, glob) -> (B,T,D)``, ``wavegen(x) -> (B,1,hop*T)``,
``detokenize(sem, glob)`` and ``halo_frames()`` (optionally ``can_stage`` / ``wavegen_stage`` / ``wavegen_staged``)
-- the native ``BiCodec`` on GPUs; the CPU gloo tests plug in the oracle to exercise exactly this partition /
exchange logic without a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` units owned by ``rank`` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# --------------------------------------------------------------------------------------- by utterance
def detokenize_utterance_sharded(model, semantic_tokens: torch.Tensor, global_tokens: torch.Tensor,
                                 group: Optional[dist.ProcessGroup] = None, gather: bool = False):
    """Each rank decodes its block of utterances; returns (local_waveform (b,1,hop*T), (lo, hi)).
    With ``gather=True`` the waveforms are all-gathered (the only collective, off the data path) and the
    full (B,1,hop*T) tensor is returned instead of the local block."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = semantic_tokens.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    wav = model.detokenize(semantic_tokens[lo:hi].contiguous(), global_tokens[lo:hi].contiguous())
    if not gather or world == 1:
        return wav, (lo, hi)
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    mx = max(h - l for l, h in sizes)
    pad = torch.zeros((mx,) + tuple(wav.shape[1:]), dtype=wav.dtype, device=wav.device)
    pad[: hi - lo] = wav
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)], dim=0), (0, B)


# --------------------------------------------------------------------------------------- by time
def _post_halo_exchange(x_own: torch.Tensor, halo: int, rank: int, has_left: bool, has_right: bool, group=None):
    """Posts (does not wait for) the exchange of my first / last ``halo`` rows with the left / right neighbour:
    one batched isend/irecv group (ncclGroupStart/End under NCCL, on NCCL's own stream).  Returns
    (requests, left_in, right_in, keepalive)."""
    ops: List[dist.P2POp] = []
    left_in = right_in = None
    keep = []
    if has_left:
        left_in = torch.empty_like(x_own[:, :halo])
        send_l = x_own[:, :halo].contiguous()
        keep.append(send_l)
        ops += [dist.P2POp(dist.isend, send_l, rank - 1, group), dist.P2POp(dist.irecv, left_in, rank - 1, group)]
    if has_right:
        right_in = torch.empty_like(x_own[:, -halo:])
        send_r = x_own[:, -halo:].contiguous()
        keep.append(send_r)
        ops += [dist.P2POp(dist.isend, send_r, rank + 1, group), dist.P2POp(dist.irecv, right_in, rank + 1, group)]
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return reqs, left_in, right_in, keep


def _wavegen_window(model, x_own, left_in, right_in, wait=None):
    """WaveGenerator over [left halo | own rows | right halo] -> the samples of the own rows.

    With a backend that can stage its input (the native ``BiCodec``: ``wavegen_stage`` / ``wavegen_staged``) the own
    (interior) rows are converted first -- a slice of the prenet output, no copy, no ``torch.cat`` -- ``wait()`` then
    orders the stream behind the neighbours' halos, and only the 2 x halo edge rows are staged after it: the
    exchange runs concurrently with the interior rows.  Other backends get the concatenated window."""
    B, n_own = x_own.shape[0], x_own.shape[1]
    n_left = left_in.shape[1] if left_in is not None else 0
    n_right = right_in.shape[1] if right_in is not None else 0
    total = n_left + n_own + n_right
    hop = model.hop
    if getattr(model, "can_stage", None) is not None and model.can_stage(B, total):
        model.wavegen_stage(x_own, total, n_left)                       # interior first
        if wait is not None:
            wait()                                                      # halos have landed (stream-ordered)
        if n_left:
            model.wavegen_stage(left_in, total, 0)
        if n_right:
            model.wavegen_stage(right_in, total, n_left + n_own)
        wav = model.wavegen_staged(B, total)
    else:
        if wait is not None:
            wait()
        parts = [p for p in (left_in, x_own, right_in) if p is not None]
        wav = model.wavegen(torch.cat(parts, dim=1) if len(parts) > 1 else x_own.contiguous())
    return wav[:, :, n_left * hop:(n_left + n_own) * hop].contiguous()


def detokenize_time_sharded(model, semantic_tokens: torch.Tensor, global_tokens: torch.Tensor,
                            group: Optional[dist.ProcessGroup] = None, exchange: bool = True):
    """Rank g decodes frames [a, b) of EVERY utterance; returns (waveform (B,1,hop*(b-a)), (a, b)).

    ``exchange=True`` : NCCL halo exchange of the prenet output (the north-star scheme), posted right after the
                        prenet and overlapped with the staging of the interior rows.
    ``exchange=False``: communication-free variant, the WaveGenerator halo is recomputed too.
    Every shard must be at least ``wavegen_halo`` frames long (checked)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, T = semantic_tokens.shape
    ph, wh = model.halo_frames()
    a, b = shard_bounds(T, world, rank)
    if world == 1:
        return model.detokenize(semantic_tokens, global_tokens), (0, T)
    if min(shard_bounds(T, world, r)[1] - shard_bounds(T, world, r)[0] for r in range(world)) < wh:
        raise ValueError(f"time shards of {T} frames over {world} ranks are shorter than the halo ({wh})")
    hop = model.hop
    if not exchange:
        lo, hi = max(0, a - ph - wh), min(T, b + ph + wh)
        wav = model.detokenize(semantic_tokens[:, lo:hi].contiguous(), global_tokens)
        return wav[:, :, (a - lo) * hop:(b - lo) * hop].contiguous(), (a, b)
    # 1. prenet on the window widened by its own receptive field; rows [a, b) are mine (a view, no copy)
    lo, hi = max(0, a - ph), min(T, b + ph)
    x = model.prenet(semantic_tokens[:, lo:hi].contiguous(), global_tokens)
    x_own = x[:, a - lo:b - lo]
    # 2. post the halo exchange with the neighbours (asynchronous)
    reqs, left_in, right_in, keep = _post_halo_exchange(x_own, wh, rank, rank > 0, rank < world - 1, group)

    def wait():
        for r in reqs:
            r.wait()

    # 3. wave generator on the widened window: interior rows first, halos when they have landed
    wav = _wavegen_window(model, x_own, left_in, right_in, wait)
    del keep
    return wav, (a, b)


def detokenize_time_windows(model, semantic_tokens: torch.Tensor, global_tokens: torch.Tensor, n_windows: int,
                            exchange: bool = True) -> torch.Tensor:
    """The time-sharded schedule run by ONE process: every utterance is cut into ``n_windows`` frame windows that
    are decoded one after the other exactly as ``n_windows`` ranks would (same bounds, same widened prenet window,
    same halo rows -- taken from the neighbouring window's prenet output instead of an NCCL message), and the
    waveforms are concatenated.  Long-form audio on one GPU with a workspace sized for one window, and the
    single-GPU check of the sharding arithmetic.  Returns (B, 1, hop*T)."""
    B, T = semantic_tokens.shape
    ph, wh = model.halo_frames()
    bounds = [shard_bounds(T, n_windows, g) for g in range(n_windows)]
    if n_windows == 1:
        return model.detokenize(semantic_tokens, global_tokens)
    if min(b - a for a, b in bounds) < wh:
        raise ValueError(f"time windows of {T} frames / {n_windows} are shorter than the halo ({wh})")
    hop = model.hop
    out = []
    if not exchange:
        for a, b in bounds:
            lo, hi = max(0, a - ph - wh), min(T, b + ph + wh)
            wav = model.detokenize(semantic_tokens[:, lo:hi].contiguous(), global_tokens)
            out.append(wav[:, :, (a - lo) * hop:(b - lo) * hop])
        return torch.cat(out, dim=2)
    own = []
    for a, b in bounds:
        lo, hi = max(0, a - ph), min(T, b + ph)
        x = model.prenet(semantic_tokens[:, lo:hi].contiguous(), global_tokens)
        own.append(x[:, a - lo:b - lo])
    for g in range(n_windows):
        left_in = own[g - 1][:, -wh:].contiguous() if g > 0 else None
        right_in = own[g + 1][:, :wh].contiguous() if g + 1 < n_windows else None
        out.append(_wavegen_window(model, own[g], left_in, right_in))
    return torch.cat(out, dim=2)
