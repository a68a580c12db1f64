"""Race hunting: many small detokenize passes (fused and one-kernel-per-conv schedules, both precisions, with taps) in
one process; a protocol bug shows up as a trapped kernel (bounded mbarrier waits).  python tools/stress_small.py [iters]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
cfg = BiCodecConfig()
m = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev)
shapes = [(2, 40), (1, 37), (3, 101), (1, 1), (5, 33), (2, 130)]
toks = [tuple(t.to(dev) for t in synthetic_tokens(cfg, B, T, 77 + B)) for B, T in shapes]
ref = {}
t0 = time.time()
for it in range(iters):
    for impl in ("tc", "tc_unfused"):
        m.set_impl(impl)
        for prec in ("fp32", "bf16"):
            for si, (sem, glob) in enumerate(toks):
                w = m.detokenize(sem, glob, precision=prec)
                if it % 4 == 0:
                    m.detokenize_tap(sem, glob, "decoder.model.3.block.4", precision=prec)
                key = (impl, prec, si)
                if key not in ref:
                    ref[key] = w.clone()
                elif not torch.equal(ref[key], w):
                    print(f"MISMATCH it={it} {key} maxdiff={(ref[key] - w).abs().max().item():.3e}", flush=True)
    torch.cuda.synchronize()
    if it % 10 == 0:
        print(f"it {it} ok ({time.time() - t0:.0f}s)", flush=True)
print("done", iters, "iterations, bit-identical throughout" , flush=True)
