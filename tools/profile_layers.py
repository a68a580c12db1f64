"""Per-launch CUDA-event profile of one detokenize pass (library profile mode): prints every launch
and an aggregate per kernel shape.  python tools/profile_layers.py [B] [T] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = torch.device("cuda:0")
cfg = BiCodecConfig()
model = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev, precision=prec)
model.validate_tokens = False
sem, glob = synthetic_tokens(cfg, B, T, 1)
sem, glob = sem.to(dev), glob.to(dev)
for _ in range(2):
    model.detokenize(sem, glob)
torch.cuda.synchronize()
passes = int(os.environ.get("PROFILE_PASSES", "5"))
rows = None
for _ in range(passes):          # min over passes per launch: the pool's GPUs / clocks are noisy
    model.profile(True)
    model.detokenize(sem, glob)
    cur = model.profile_read()
    model.profile(False)
    if rows is None:
        rows = cur
    else:
        for a, b in zip(rows, cur):
            a["ms"] = min(a["ms"], b["ms"])
total = sum(r["ms"] for r in rows)
print(f"# B={B} T={T} {prec}: {len(rows)} launches, {total:.2f} ms total (sum of per-launch events, min over {passes} passes)")
from spark_tts_b200 import _lib
work = float(_lib.load().sparkcodec_fp32_terms()) if prec == "fp32" else 1.0   # bf16-MMA equivalents per MAC
agg = {}
order = []
for r in rows:
    if r["name"] not in agg:
        order.append(r["name"])
    a = agg.setdefault(r["name"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
    a["ms"] += r["ms"]; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]; a["n"] += 1
print(f"{'kernel':86s} {'n':>3s} {'ms':>8s} {'%':>6s} {'TF/s alg':>9s} {'pipe%':>6s} {'GB/s':>8s}")
for name in order:
    a = agg[name]
    tf = a["flops"] / (a["ms"] * 1e-3) / 1e12
    gb = a["bytes"] / (a["ms"] * 1e-3) / 1e9
    print(f"{name:86s} {a['n']:3d} {a['ms']:8.3f} {100 * a['ms'] / total:6.2f} {tf:9.1f} {100 * tf * work / 1416.3:6.1f} {gb:8.0f}")
