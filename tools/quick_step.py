"""Back-to-back detokenize steps timed with one CUDA-event pair (no per-launch events): python tools/quick_step.py [B] [T] [precision] [steps]"""
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
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
cfg = BiCodecConfig()
m = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev, precision=prec)
m.validate_tokens = False
sem, glob = synthetic_tokens(cfg, B, T, 1)
sem, glob = sem.to(dev), glob.to(dev)
for _ in range(3):
    m.detokenize(sem, glob)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    wav = m.detokenize(sem, glob)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"B={B} T={T} {prec} PDL={os.environ.get('SPARKCODEC_PDL', '1')}: {ms:.3f} ms/step, {B * T / 50 / (ms * 1e-3):.0f} audio-s/s, checksum {float(wav.double().abs().sum()):.6f}")
