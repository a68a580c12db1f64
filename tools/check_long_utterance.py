import sys; sys.path.insert(0, "/root/repo")
import torch
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens
cfg = BiCodecConfig(); dev = torch.device("cuda:0")
m = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev)
sem, glob = synthetic_tokens(cfg, 2, 6001, 77)      # 120 s, odd frame count
semd, globd = sem.to(dev), glob.to(dev)
full = m.detokenize(semd, globd)
assert bool(torch.isfinite(full).all())
# interior window decoded separately must match away from the edges (receptive field 67 frames)
a, b = 3000, 3400
win = m.detokenize(semd[:, a - 80:b + 80].contiguous(), globd)
hop = cfg.hop
d = (full[:, :, a * hop:b * hop] - win[:, :, 80 * hop:(80 + b - a) * hop]).abs().max().item()
print("long-T ok", tuple(full.shape), "interior max-abs diff", d)
assert d < 1e-4
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): m.detokenize(semd, globd)
torch.cuda.synchronize(); print("ms per 2x120s:", (time.perf_counter() - t0) / 3 * 1e3)
