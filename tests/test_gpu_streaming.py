"""GPU: streaming chunked detokenize -- every chunk equals the oracle's independent decode of that chunk
(what the reference's Triton vocoder does per chunk), eager and CUDA-graph paths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(cfg, state_dict):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spark_tts_b200 import BiCodec
    return BiCodec.from_state_dict(cfg, state_dict, device=torch.device("cuda:0"))


@pytest.mark.parametrize("use_graphs", [False, True])
def test_streams_match_oracle_per_chunk(use_graphs, model, cfg, state_dict):
    from oracle import bicodec_oracle as O
    from spark_tts_b200.streaming import StreamingDetokenizer, chunk_schedule
    from spark_tts_b200.synthetic import synthetic_tokens
    sd = StreamingDetokenizer(model, use_graphs=use_graphs, graph_min_batch=2)
    lengths = {0: 130, 1: 130, 2: 57, 3: 50}
    toks = {}
    for sid, n in lengths.items():
        sem, glob = synthetic_tokens(cfg, 1, n, 300 + sid)
        toks[sid] = (sem[0], glob[0, 0])
        sd.open(sid, glob[0, 0])
    got = {sid: [] for sid in lengths}
    pos = {sid: 0 for sid in lengths}
    while any(pos[s] < lengths[s] for s in lengths):           # tokens trickle in 10 at a time
        for sid in lengths:
            if pos[sid] < lengths[sid]:
                nxt = min(pos[sid] + 10, lengths[sid])
                sd.push(sid, toks[sid][0][pos[sid]:nxt].tolist())
                pos[sid] = nxt
                if nxt == lengths[sid]:
                    sd.close(sid)
        for sid, chunks in sd.poll().items():
            got[sid].extend(chunks)
    assert not sd.streams
    for sid, n in lengths.items():
        sched = chunk_schedule(n, sd.policy)
        assert len(got[sid]) == len(sched)
        for (a, b), wav in zip(sched, got[sid]):
            ref = O.detokenize(state_dict, cfg, toks[sid][0][a:b].unsqueeze(0), toks[sid][1].view(1, 1, -1))
            ref = ref.view(-1)
            w = torch.from_numpy(wav)
            assert w.shape == ref.shape
            assert (ref - w).abs().max().item() <= 1e-3 and O.snr_db(ref, w) >= 60.0


def test_graph_replay_is_bit_identical_to_eager(model, cfg):
    from spark_tts_b200.streaming import StreamingDetokenizer
    from spark_tts_b200.synthetic import synthetic_tokens
    sem, glob = synthetic_tokens(cfg, 16, 50, 77)
    eager = StreamingDetokenizer(model, use_graphs=False).decode_batch(sem, glob.squeeze(1)).clone()
    g = StreamingDetokenizer(model, use_graphs=True)
    a = g.decode_batch(sem, glob.squeeze(1)).clone()
    sem2, glob2 = synthetic_tokens(cfg, 16, 50, 78)
    g.decode_batch(sem2, glob2.squeeze(1))
    b = g.decode_batch(sem, glob.squeeze(1)).clone()           # replay with fresh inputs, then the old ones again
    assert torch.equal(a, eager) and torch.equal(a, b)


def test_old_graph_survives_workspace_growth(model, cfg):
    """a captured graph must not point into the model's shared workspace, which is dropped and
    re-allocated when a later call needs more room.  Replay shape A after a larger shape B grew the workspace,
    with a sentinel tensor parked in the block the old workspace freed: A still equals the eager result and the
    sentinel is untouched."""
    from spark_tts_b200.streaming import StreamingDetokenizer
    from spark_tts_b200.synthetic import synthetic_tokens
    dev = model.device
    sem_a, glob_a = synthetic_tokens(cfg, 8, 50, 501)
    eager = StreamingDetokenizer(model, use_graphs=False).decode_batch(sem_a, glob_a.squeeze(1)).clone()
    model._ws = None
    torch.cuda.empty_cache()
    model.detokenize(sem_a.to(dev), glob_a.to(dev))                            # shared workspace sized for (8, 50)
    g = StreamingDetokenizer(model, use_graphs=True, graph_min_batch=2)
    first = g.decode_batch(sem_a, glob_a.squeeze(1)).clone()                   # captures (8, 50)
    assert torch.equal(first, eager)
    old_bytes = model._ws.numel() if model._ws is not None else 0
    sem_b, glob_b = synthetic_tokens(cfg, 8, 400, 502)
    g.decode_batch(sem_b, glob_b.squeeze(1))                                   # captures (8, 400)
    model.detokenize(sem_b.to(dev), glob_b.to(dev))                            # eager call: shared workspace grows
    assert model._ws.numel() > old_bytes
    # whatever the caching allocator hands out next may be the freed block of the old workspace
    sentinels = [torch.full((max(old_bytes, 1 << 20) // 4,), 1234.5, device=dev) for _ in range(3)]
    again = g.decode_batch(sem_a, glob_a.squeeze(1)).clone()                   # replay of the OLD graph
    torch.cuda.synchronize(dev)
    assert torch.equal(again, eager)
    for s in sentinels:
        assert bool((s == 1234.5).all())


def test_graph_cache_is_bounded_and_one_off_shapes_stay_eager(model, cfg):
    """only schedule sizes (or shapes seen graph_after times) are captured; LRU eviction."""
    from spark_tts_b200.streaming import StreamingDetokenizer
    from spark_tts_b200.synthetic import synthetic_tokens
    g = StreamingDetokenizer(model, use_graphs=True, graph_min_batch=2, graph_cache_size=2, graph_after=3)
    for T in (7, 13, 21, 33):                                                  # flush chunks of arbitrary length
        sem, glob = synthetic_tokens(cfg, 4, T, 600 + T)
        g.decode_batch(sem, glob.squeeze(1))
    assert len(g._graphs) == 0
    sem, glob = synthetic_tokens(cfg, 4, 21, 700)
    ref = StreamingDetokenizer(model, use_graphs=False).decode_batch(sem, glob.squeeze(1)).clone()
    for _ in range(3):                                                         # third sighting of (4, 21): captured
        out = g.decode_batch(sem, glob.squeeze(1)).clone()
        assert torch.equal(out, ref)
    assert [k[:2] for k in g._graphs] == [(4, 21)]
    for B in (2, 3, 5):                                                        # schedule size 50: captured at once
        sem, glob = synthetic_tokens(cfg, B, 50, 710 + B)
        g.decode_batch(sem, glob.squeeze(1))
        assert len(g._graphs) <= 2
    assert [k[:2] for k in g._graphs] == [(3, 50), (5, 50)]                    # least recently used went first


def test_graphs_are_dropped_when_the_handle_is_rebuilt(cfg, state_dict):
    """graphs captured against a destroyed handle are never replayed (generation counter)."""
    from spark_tts_b200 import BiCodec
    from spark_tts_b200.streaming import StreamingDetokenizer
    from spark_tts_b200.synthetic import synthetic_tokens
    m = BiCodec.from_state_dict(cfg, state_dict, device=torch.device("cuda:0"))
    g = StreamingDetokenizer(m, use_graphs=True, graph_min_batch=2)
    sem, glob = synthetic_tokens(cfg, 4, 50, 800)
    a = g.decode_batch(sem, glob.squeeze(1)).clone()
    old = next(iter(g._graphs.values()))
    m._build(m.device)                                                         # what .to(other_device) does
    b = g.decode_batch(sem, glob.squeeze(1)).clone()
    new = next(iter(g._graphs.values()))
    assert new is not old and new.generation == m._generation and len(g._graphs) == 1
    assert torch.equal(a, b)
