"""CPU simulation of tensor-core operand splits through the oracle (test tooling; imports oracle/).

Every dense contraction of the detokenize path is replaced by a sum of products of ROUNDED operands, accumulated
in float64 (the tensor cores accumulate exactly-representable products in fp32; the split error dominates), so the
waveform SNR against the plain fp32 oracle shows what a split costs before any kernel is written.

  python tools/sim_split_precision.py [B] [T]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import oracle.bicodec_oracle as O
from spark_tts_b200.synthetic import synthetic_state_dict
from spark_tts_b200.config import BiCodecConfig

F8 = {"e5m2": torch.float8_e5m2, "e4m3": torch.float8_e4m3fn}


def rnd(x, kind):
    if kind == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if kind == "fp16":
        return x.to(torch.float16).to(torch.float32)
    if kind in F8:
        lim = 57344.0 if kind == "e5m2" else 448.0
        return x.clamp(-lim, lim).to(F8[kind]).to(torch.float32)
    raise ValueError(kind)


class Scheme:
    def __init__(self, name, hi, terms):
        self.name, self.hi, self.terms = name, hi, terms

    def products(self, a, w):
        """list of (A_part, W_part) whose products are summed"""
        a_hi, w_hi = rnd(a, self.hi), rnd(w, self.hi)
        a_lo, w_lo = a - a_hi, w - w_hi
        out = [(a_hi, w_hi)]
        for t in self.terms:
            if t == "lo*hi":
                out.append((rnd(a_lo, self.hi), w_hi))
            elif t == "hi*lo":
                out.append((a_hi, rnd(w_lo, self.hi)))
            elif t[0] == "lo8*hi8":          # (kind_a, kind_w, log2 scale on A)
                _, ka, kw, s = t
                out.append((rnd(a_lo * 2.0 ** s, ka), rnd(w_hi * 2.0 ** -s, kw)))
            elif t[0] == "hi8*lo8":
                _, ka, kw, s = t
                out.append((rnd(a_hi * 2.0 ** -s, ka), rnd(w_lo * 2.0 ** s, kw)))
        return out


SCHEMES = [
    Scheme("bf16 x1", "bf16", []),
    Scheme("bf16 x3 (current fp32 mode)", "bf16", ["lo*hi", "hi*lo"]),
    Scheme("fp16 x1", "fp16", []),
    Scheme("fp16 hi*hi + lo*hi (2 terms)", "fp16", ["lo*hi"]),
    Scheme("fp16 x3", "fp16", ["lo*hi", "hi*lo"]),
    Scheme("fp16 hi*hi + e5m2 corrections s=4/10", "fp16", [("lo8*hi8", "e5m2", "e5m2", 4), ("hi8*lo8", "e5m2", "e5m2", 10)]),
    Scheme("fp16 hi*hi + e5m2 corrections s=2/12", "fp16", [("lo8*hi8", "e5m2", "e5m2", 2), ("hi8*lo8", "e5m2", "e5m2", 12)]),
    Scheme("fp16 hi*hi + e5m2 corrections s=8/8", "fp16", [("lo8*hi8", "e5m2", "e5m2", 8), ("hi8*lo8", "e5m2", "e5m2", 8)]),
    Scheme("fp16 hi*hi + e4m3(A)/e5m2(W) s=12/6", "fp16", [("lo8*hi8", "e4m3", "e5m2", 12), ("hi8*lo8", "e4m3", "e5m2", -2)]),
]

_cur = None
_conv1d, _linear, _convT = F.conv1d, F.linear, F.conv_transpose1d


def dense(fn, a, w, b, **kw):
    if _cur is None:
        return fn(a, w, b, **kw)
    acc = None
    for ap, wp in _cur.products(a, w):
        y = fn(ap.double(), wp.double(), None, **kw)
        acc = y if acc is None else acc + y
    acc = acc.float()
    if b is not None:
        acc = acc + (b.view(1, -1, 1) if fn is not _linear else b)
    return acc


def conv1d(a, w, b=None, stride=1, padding=0, dilation=1, groups=1):
    if groups != 1 or w.shape[0] < 8 or w.shape[1] < 16:       # depthwise conv, head, out_project (8 -> 1024): FFMA kernels
        return _conv1d(a, w, b, stride, padding, dilation, groups)
    return dense(_conv1d, a, w, b, padding=padding, dilation=dilation)


def linear(a, w, b=None):
    if w.shape[1] < 16 or a.dim() < 3:                         # FSQ project_out, speaker project, AdaLN GEMVs: FFMA
        return _linear(a, w, b)
    if w.shape == (384, 1024) and a.shape[-1] == 1024:         # linear_pre is folded into the VQ table (fp32)
        return _linear(a, w, b)
    return dense(_linear, a, w, b)


def convT(a, w, b=None, stride=1, padding=0):
    return dense(_convT, a, w, b, stride=stride, padding=padding)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    global _cur
    cfg = BiCodecConfig()
    sd = synthetic_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(7)
    sem = torch.randint(0, 8192, (B, T), generator=g)
    glo = torch.randint(0, 4096, (B, 1, 32), generator=g, dtype=torch.int32)
    ref = O.detokenize(sd, cfg, sem, glo)
    F.conv1d, F.linear, F.conv_transpose1d = conv1d, linear, convT
    O.F.conv1d, O.F.linear, O.F.conv_transpose1d = conv1d, linear, convT
    print(f"B={B} T={T} ref rms {ref.pow(2).mean().sqrt():.4f} absmax {ref.abs().max():.3f}")
    for s in SCHEMES:
        _cur = s
        out = O.detokenize(sd, cfg, sem, glo)
        print(f"{s.name:55s} SNR {O.snr_db(ref, out):6.2f} dB   max-abs {float((out - ref).abs().max()):.2e}", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        main()
