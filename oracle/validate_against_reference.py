"""Pin the oracle restatement against the reference's own modules (run in THIS container only).

    python oracle/validate_against_reference.py

Builds the reference modules from /root/reference, loads the synthetic checkpoint into them and
checks stage by stage that ``oracle/bicodec_oracle.py`` reproduces their outputs.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import bicodec_oracle as O                      # noqa: E402
from oracle.reference_loader import ReferenceDetokenizer    # noqa: E402
from spark_tts_b200.config import BiCodecConfig             # noqa: E402
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens  # noqa: E402


def main():
    torch.manual_seed(0)
    cfg = BiCodecConfig()
    sd = synthetic_state_dict(cfg, seed=0)
    ref = ReferenceDetokenizer(cfg).load_checkpoint(sd)
    ok = True
    for (B, T, seed) in [(1, 50, 11), (2, 37, 12), (1, 1, 13)]:
        sem, glob = synthetic_tokens(cfg, B, T, seed)
        with torch.no_grad():
            z_ref = ref.quantizer.detokenize(sem)
            d_ref = ref.speaker_encoder.detokenize(glob, onnx_export_mode=True)
            x_ref = ref.prenet(z_ref, d_ref)
            wav_ref = ref.detokenize(sem, glob)
        taps = {}
        wav = O.detokenize(sd, cfg, sem, glob, taps)
        z = O.vq_detokenize(sd, sem)
        d = O.speaker_detokenize(sd, glob, cfg.fsq_levels)
        x = O.prenet(sd, z, d)
        rows = [
            ("z_q", z_ref, z), ("d_vector", d_ref, d), ("prenet", x_ref, x), ("wav", wav_ref, wav),
        ]
        for name, a, b in rows:
            err = (a - b).abs().max().item()
            snr = O.snr_db(a, b)
            bit = torch.equal(a, b)
            print(f"B={B} T={T} {name:9s} shape={tuple(b.shape)} max_abs={err:.3e} snr={snr:.1f} dB bit_exact={bit}")
            ok &= (snr > 100.0)
        print(f"    wav rms={wav_ref.pow(2).mean().sqrt().item():.4f} absmax={wav_ref.abs().max().item():.4f}")
    print("ORACLE PINNED" if ok else "ORACLE MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
