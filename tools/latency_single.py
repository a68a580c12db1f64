"""Single-utterance latency (the CLI case, cli/SparkTTS.py: one utterance per call): tokens on device -> waveform on
device, CUDA events, median of 50 after 10 warm-ups.  python tools/latency_single.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

dev = torch.device("cuda:0")
cfg = BiCodecConfig()
sd = synthetic_state_dict(cfg, 0)
for prec in ("fp32", "bf16"):
    m = BiCodec.from_state_dict(cfg, sd, device=dev, precision=prec)
    m.validate_tokens = False
    for B, T, graphs in ((1, 50, False), (1, 50, True), (1, 500, False), (1, 500, True), (1, 1500, False),
                         (1, 1500, True), (8, 500, False), (8, 500, True)):
        m.use_graphs = graphs
        sem, glob = synthetic_tokens(cfg, B, T, 5)
        sem, glob = sem.to(dev), glob.to(dev)
        for _ in range(10):
            m.detokenize(sem, glob)
        ts = []
        for _ in range(50):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m.detokenize(sem, glob)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"{prec} {'graph' if graphs else 'eager'} B={B} T={T} ({B * T / 50:.0f} s audio): median {ts[25]:.3f} ms  p95 {ts[47]:.3f} ms  "
              f"-> {B * T / 50 / (ts[25] * 1e-3):.0f} audio-s/s, RTF {ts[25] * 1e-3 / (B * T / 50):.2e}")
