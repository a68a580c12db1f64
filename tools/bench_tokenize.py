"""Throughput of the semantic tokenize row (sparkcodec_tokenize_semantic): features on device -> tokens on device,
CUDA events, median of 10 after 3 warm-ups, plus the per-kernel breakdown from the library's profile mode.
python tools/bench_tokenize.py [B] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_encoder_state_dict, synthetic_features, synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
dev = torch.device("cuda:0")
cfg = BiCodecConfig()
sd = {**synthetic_state_dict(cfg, 0), **synthetic_encoder_state_dict(cfg, 0)}
feat = synthetic_features(cfg, B, T, 5).to(dev)
for prec in ("fp32", "bf16"):
    m = BiCodec.from_state_dict(cfg, sd, device=dev, precision=prec)
    for _ in range(3):
        m.tokenize_semantic(feat)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.tokenize_semantic(feat)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{prec} B={B} T={T}: median {med:.3f} ms -> {B * T / 50 / (med * 1e-3):.0f} audio-s/s "
          f"({B * T / (med * 1e-3) / 1e6:.2f} M frames/s)")
    m.profile(True)
    m.tokenize_semantic(feat)
    rows = m.profile_read()
    m.profile(False)
    agg = {}
    for r in rows:
        a = agg.setdefault(r["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += r["ms"]; a[2] += r["flops"]; a[3] += r["bytes"]
    for name, (n, ms, fl, by) in agg.items():
        print(f"    {name:86s} x{n:<3d} {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TF/s  {by / ms / 1e6:7.0f} GB/s")

# ---- speaker half (sparkcodec_tokenize_speaker): reference clips on device -> global tokens on device
from spark_tts_b200.synthetic import synthetic_ref_wav, synthetic_speaker_state_dict

sd_spk = {**sd, **synthetic_speaker_state_dict(cfg, 0)}
m = BiCodec.from_state_dict(cfg, sd_spk, device=dev)
for Bs in (1, 16):
    wav = synthetic_ref_wav(cfg, Bs, 6.0, 6).to(dev)
    for _ in range(3):
        m.tokenize_speaker(wav)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.tokenize_speaker(wav)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"speaker tokenize fp32 B={Bs} x 6 s clip: median {ts[len(ts) // 2]:.3f} ms ({Bs / (ts[len(ts) // 2] * 1e-3):.0f} clips/s)")
