"""The tail of the reference CLI (cli/SparkTTS.py:213-234) on the B200 path: decoded LLM text (or a file of it) with
``<|bicodec_global_N|>`` / ``<|bicodec_semantic_N|>`` tokens in, a 16 kHz mono WAV out.

    python examples/vocode_tokens.py --model-dir pretrained_models/Spark-TTS-0.5B --text-file llm_output.txt --out out.wav
    python examples/vocode_tokens.py --synthetic --seconds 5 --out demo.wav        # no checkpoint: random-init weights

The reference writes the array with soundfile (cli/inference.py); the standard-library ``wave`` module is used here
because soundfile is not a dependency of this package (16-bit PCM, same sample rate).
"""
from __future__ import annotations

import argparse
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def write_wav(path: str, samples: np.ndarray, sample_rate: int) -> None:
    """float32 [-1, 1] mono -> 16-bit PCM WAV."""
    pcm = (np.clip(samples.astype(np.float64), -1.0, 1.0) * 32767.0).round().astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--model-dir", help="Spark-TTS-0.5B directory (holds BiCodec/config.yaml + model.safetensors)")
    ap.add_argument("--text", help="decoded LLM output containing the bicodec tokens")
    ap.add_argument("--text-file", help="file with the decoded LLM output")
    ap.add_argument("--synthetic", action="store_true", help="random-init weights and random tokens (smoke demo)")
    ap.add_argument("--seconds", type=float, default=5.0, help="--synthetic: length of the demo utterance")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--out", required=True)
    args = ap.parse_args(argv)

    import torch

    from spark_tts_b200 import BiCodec, BiCodecConfig, BiCodecTokenizer, token_feed

    dev = torch.device(args.device)
    if args.synthetic:
        from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens
        cfg = BiCodecConfig()
        tok = BiCodecTokenizer(device=dev, model=BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0),
                                                                         precision=args.precision))
        sem, glob = synthetic_tokens(cfg, 1, max(1, int(round(args.seconds * cfg.frame_rate))), 7)
        glob = glob.squeeze(1)
    else:
        if not args.model_dir or not (args.text or args.text_file):
            ap.error("--model-dir and --text/--text-file are required without --synthetic")
        text = args.text if args.text is not None else open(args.text_file).read()
        sem, glob = token_feed.codes_from_text(text)                 # cli/SparkTTS.py:213-228
        glob = glob.squeeze(0)
        tok = BiCodecTokenizer(args.model_dir, device=dev, precision=args.precision)
        cfg = tok.model.cfg
        if sem.shape[1] == 0 or glob.shape[1] != cfg.token_num:
            raise SystemExit(f"need >= 1 semantic token and exactly {cfg.token_num} global tokens, "
                             f"got {sem.shape[1]} / {glob.shape[1]}")
    wav = tok.detokenize(glob.to(dev), sem.to(dev))                  # cli/SparkTTS.py:231-234
    write_wav(args.out, np.asarray(wav).reshape(-1), cfg.sample_rate)
    print(f"{args.out}: {wav.size / cfg.sample_rate:.2f} s at {cfg.sample_rate} Hz from {sem.shape[1]} semantic tokens")
    return 0


if __name__ == "__main__":
    sys.exit(main())
