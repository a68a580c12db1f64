"""One tiny detokenize pass (B=1, T=8 by default) in both precision modes, plus the staged-wavegen and semantic-tokenize
entry points, for `compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck}` (tools/sanitize_run.sh).
Exits non-zero if the result is not finite."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_encoder_state_dict, synthetic_features, synthetic_state_dict, synthetic_tokens


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda:0")
    cfg = BiCodecConfig()
    sd = {**synthetic_state_dict(cfg, 0), **synthetic_encoder_state_dict(cfg, 0)}
    m = BiCodec.from_state_dict(cfg, sd, device=dev)
    sem, glob = synthetic_tokens(cfg, B, T, 11)
    for prec in ("fp32", "bf16"):
        wav = m.detokenize(sem.to(dev), glob.to(dev), precision=prec)
        assert bool(torch.isfinite(wav).all()), prec
    x = m.prenet(sem.to(dev), glob.to(dev))
    m.wavegen_stage(x, T, 0)
    assert bool(torch.isfinite(m.wavegen_staged(B, T)).all())
    tok = m.tokenize_semantic(synthetic_features(cfg, B, T, 12).to(dev))
    assert tok.shape == (B, T)
    torch.cuda.synchronize()
    print(f"sanitize_pass ok: B={B} T={T} launches={m.launch_count()}")


if __name__ == "__main__":
    main()
