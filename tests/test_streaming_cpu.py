"""CPU: chunk schedule / cross-fade bookkeeping of spark_tts_b200.streaming against a literal restatement of
the reference's loop (runtime/triton_trtllm/model_repo/spark_tts/1/model.py:351-385, client_grpc.py:390-415)."""
import math

import numpy as np
import pytest

from spark_tts_b200.streaming import ChunkPolicy, chunk_schedule, cross_fade, schedule_sizes


def _reference_chunks(n_tokens, frame_rate=50, dur=1.0, max_dur=30.0, scale=8.0, ov=0.1):
    arr = []
    max_chunk = math.ceil(max_dur * frame_rate)
    chunk = math.ceil(dur * frame_rate)
    overlap = math.ceil(ov * frame_rate)
    out = []
    for tok in range(n_tokens):              # tokens arrive one at a time from the LLM
        arr.append(tok)
        if len(arr) >= chunk:
            out.append(arr[:chunk])
            arr = arr[chunk - overlap:]
            chunk = min(max_chunk, int(chunk * scale))
    if len(arr) > 0:
        out.append(arr)
    return out


@pytest.mark.parametrize("n", [0, 1, 49, 50, 51, 55, 449, 450, 451, 2000, 4000])
def test_chunk_schedule_matches_reference_loop(n):
    got = chunk_schedule(n, ChunkPolicy())
    ref = _reference_chunks(n)
    assert [list(range(a, b)) for a, b in got] == ref


def test_policy_defaults_are_run_sh_values():
    p = ChunkPolicy()
    assert (p.first_chunk, p.overlap, p.max_chunk) == (50, 5, 1500)


def test_schedule_sizes_are_the_only_graphed_chunk_lengths():
    """The growth rule 50 -> 400 -> 1500 (x8, capped at 30 s): the shapes the streaming server captures as graphs."""
    assert schedule_sizes(ChunkPolicy()) == [50, 400, 1500]
    assert schedule_sizes(ChunkPolicy(scale=1.0)) == [50]


def test_cross_fade_reconstruction():
    rng = np.random.default_rng(0)
    ov = 1600
    chunks = [rng.standard_normal(16000).astype(np.float32), rng.standard_normal(8000).astype(np.float32),
              rng.standard_normal(4000).astype(np.float32)]
    out = cross_fade(chunks, ov)
    assert out.shape[0] == 16000 + (8000 - ov) + (4000 - ov)
    assert np.array_equal(out[:16000 - ov], chunks[0][:-ov])
    assert np.allclose(out[16000 - ov], chunks[0][-ov], atol=1e-6)          # fade starts on the old chunk
    assert np.array_equal(out[-ov:], chunks[2][-ov:])
    same = [np.ones(4000, np.float32)] * 3                                     # constant signal stays constant
    assert np.allclose(cross_fade(same, ov), 1.0, atol=1e-6)
    assert cross_fade([], ov).size == 0 and cross_fade(chunks[:1], ov) is chunks[0]
