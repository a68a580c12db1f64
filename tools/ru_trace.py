"""Schedule trace of the fused ResidualUnit kernels (SPARKCODEC_RU_TRACE=1): per-tile event clocks of CTA 0.
python tools/ru_trace.py [fp32|bf16]"""
import os
import sys

os.environ["SPARKCODEC_RU_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
cfg = BiCodecConfig()
model = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev, precision=prec)
model.validate_tokens = False
sem, glob = synthetic_tokens(cfg, 16, 500, 1)
model.detokenize(sem.to(dev), glob.to(dev))
torch.cuda.synchronize()
