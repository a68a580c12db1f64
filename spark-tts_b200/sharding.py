"""Multi-GPU partitioning of the detokenize path (one process per GPU, torch.distributed).

Two modes:

* **utterance sharding** (BASELINE config 3): utterances are independent units -- contiguous blocks of
  B/N utterances per rank, weights replicated, NO collective on the data path.
* **time sharding** (BASELINE config 5): every utterance is cut into N contiguous frame windows.  The
  network is a finite-receptive-field stack, so rank g
    1. runs the prenet on its window widened by ``prenet_halo`` frames (tokens are replicated, so this
       is recompute of 57 cheap 50 Hz frames, no communication) and keeps x for its own frames,
    2. exchanges ``wavegen_halo`` boundary rows of x = prenet(z_q, d) + d with its left / right
       neighbour (NCCL send/recv over NVLink: B x halo x 1024 fp32 per direction),
    3. runs the WaveGenerator on the widened window and keeps the samples of its own frames.
  Zero padding applies at true utterance edges only; rows inside the halo of a shard edge are
  discarded, so the result equals the un-sharded one.

The compute backend is any object with ``prenet(sem, glob) -> (B,T,D)``, ``wavegen(x) -> (B,1,hop*T)``,
``detokenize(sem, glob)`` and ``halo_frames()`` -- the native ``BiCodec`` on GPUs; the CPU gloo tests
plug in the oracle to exercise exactly this partition / exchange logic without a GPU.
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
def _exchange_halos(x_own: torch.Tensor, halo: int, rank: int, world: int, has_left: bool, has_right: bool,
                    group=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Send my first/last ``halo`` rows to the left/right neighbour, receive theirs.  One batched
    isend/irecv group per call (ncclGroupStart/End under NCCL)."""
    ops: List[dist.P2POp] = []
    left_in = right_in = None
    if has_left:
        left_in = torch.empty_like(x_own[:, :halo])
        send_l = x_own[:, :halo].contiguous()
        ops += [dist.P2POp(dist.isend, send_l, rank - 1, group), dist.P2POp(dist.irecv, left_in, rank - 1, group)]
    if has_right:
        right_in = torch.empty_like(x_own[:, -halo:])
        send_r = x_own[:, -halo:].contiguous()
        ops += [dist.P2POp(dist.isend, send_r, rank + 1, group), dist.P2POp(dist.irecv, right_in, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return left_in, right_in


def detokenize_time_sharded(model, semantic_tokens: torch.Tensor, global_tokens: torch.Tensor,
                            group: Optional[dist.ProcessGroup] = None, exchange: bool = True):
    """Rank g decodes frames [a, b) of EVERY utterance; returns (waveform (B,1,hop*(b-a)), (a, b)).

    ``exchange=True`` : NCCL halo exchange of the prenet output (the north-star scheme).
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
    # 1. prenet on the window widened by its own receptive field; keep rows [a, b)
    lo, hi = max(0, a - ph), min(T, b + ph)
    x = model.prenet(semantic_tokens[:, lo:hi].contiguous(), global_tokens)
    x_own = x[:, a - lo:b - lo].contiguous()
    # 2. halo exchange with the neighbours
    left_in, right_in = _exchange_halos(x_own, wh, rank, world, rank > 0, rank < world - 1, group)
    parts = [p for p in (left_in, x_own, right_in) if p is not None]
    x_wide = torch.cat(parts, dim=1) if len(parts) > 1 else x_own
    # 3. wave generator on the widened window; keep my samples
    wav = model.wavegen(x_wide)
    off = wh if left_in is not None else 0
    return wav[:, :, off * hop:(off + (b - a)) * hop].contiguous(), (a, b)
