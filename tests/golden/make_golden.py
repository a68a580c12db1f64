"""Generate the golden fixtures from the REFERENCE's own modules (run in this container only):

    python tests/golden/make_golden.py

Weights are the synthetic checkpoint (seed 0, regenerated bit-identically anywhere from
``spark_tts_b200.synthetic``); inputs/outputs of the reference modules
(/root/reference sparktts/models/bicodec.py:171-189 composed as export_sparktts_onnx.py:267-312)
are stored as tests/golden/detok_*.npz.  The GPU box has no /root/reference, so GPU parity tests
compare against these files and against the oracle restatement.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bicodec_oracle as O                                          # noqa: E402
from oracle.reference_loader import (ReferenceDetokenizer, ReferenceSemanticTokenizer,   # noqa: E402
                                     ReferenceSpeakerTokenizer)
from spark_tts_b200.config import BiCodecConfig                     # noqa: E402
from spark_tts_b200.synthetic import (synthetic_encoder_state_dict, synthetic_features,   # noqa: E402
                                      synthetic_ref_wav, synthetic_speaker_state_dict, synthetic_state_dict,
                                      synthetic_tokens)

# name, batch, frames, token seed, semantic dtype, global dtype
CASES = [
    ("a_b1_t25", 1, 25, 101, torch.int64, torch.int32),
    ("b_b3_t16", 3, 16, 102, torch.int32, torch.int64),
    ("c_b1_t1", 1, 1, 103, torch.int64, torch.int32),
    ("d_b2_t130", 2, 130, 104, torch.int64, torch.int32),
]
WEIGHT_SEED = 0
# semantic tokenize (encode side): name, batch, frames, feature seed
# speaker tokenize (encode side): name, batch, clip seconds, waveform seed
SPEAKER_CASES = [("a_b2_6s", 2, 6.0, 301), ("b_b1_1s", 1, 0.96, 302), ("c_b3_2s", 3, 2.0, 303)]
TOKENIZE_CASES = [("a_b2_t60", 2, 60, 201), ("b_b1_t7", 1, 7, 202), ("c_b3_t1", 3, 1, 203), ("d_b1_t300", 1, 300, 204)]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    cfg = BiCodecConfig()
    sd = synthetic_state_dict(cfg, seed=WEIGHT_SEED)
    ref = ReferenceDetokenizer(cfg).load_checkpoint(sd)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, B, T, seed, sdt, gdt in CASES:
        sem, glob = synthetic_tokens(cfg, B, T, seed)
        sem, glob = sem.to(sdt), glob.to(gdt)
        with torch.no_grad():
            z_q = ref.quantizer.detokenize(sem)                                    # (B,1024,T)
            codebook_rows = ref.quantizer.embed_code(sem.long())                   # (B,T,8) exact gather
            fsq = ref.speaker_encoder.quantizer.get_codes_from_indices(
                glob.long().transpose(1, 2), onnx_export_mode=True).squeeze(0)    # (B,32,6) exact
            d = ref.speaker_encoder.detokenize(glob, onnx_export_mode=True)        # (B,1024)
            x = ref.prenet(z_q, d) + d.unsqueeze(-1)                               # (B,1024,T)
            wav = ref.detokenize(sem, glob)                                        # (B,1,320T)
        np.savez_compressed(
            os.path.join(out_dir, f"detok_{name}.npz"),
            weight_seed=np.int64(WEIGHT_SEED),
            semantic_tokens=sem.numpy(), global_tokens=glob.numpy(),
            codebook_rows=codebook_rows.numpy(), fsq_codes=fsq.numpy(),
            z_q_first8=z_q[:, :, :8].numpy(), d_vector=d.numpy(),
            prenet_plus_d_first8=x[:, :, :8].numpy(), output_waveform=wav.numpy(),
        )
        print(name, "wav", tuple(wav.shape), "rms", float(wav.pow(2).mean().sqrt()))
    # semantic half of BiCodec.tokenize (bicodec.py:151-169) through the reference's Encoder + quantizer.tokenize;
    # the features are regenerated from the seed (synthetic_features), only the answers are stored
    sd_all = {**sd, **synthetic_encoder_state_dict(cfg, seed=WEIGHT_SEED)}
    tok = ReferenceSemanticTokenizer(cfg).load_checkpoint(sd_all)
    for name, B, T, seed in TOKENIZE_CASES:
        feat = synthetic_features(cfg, B, T, seed)
        idx = tok.tokenize(feat)
        _, margin = O.tokenize_semantic(sd_all, cfg, feat)
        np.savez_compressed(os.path.join(out_dir, f"tokenize_{name}.npz"), weight_seed=np.int64(WEIGHT_SEED),
                            feat_seed=np.int64(seed), batch=np.int64(B), frames=np.int64(T),
                            feat_checksum=np.float64(feat.double().sum().item()),
                            semantic_tokens=idx.numpy(), margin=margin.numpy())
        print("tokenize", name, tuple(idx.shape), "min margin", float(margin.min()))
    # speaker half of BiCodec.tokenize (bicodec.py:162-167) through the reference's mel transform + SpeakerEncoder;
    # the clips are regenerated from the seed (synthetic_ref_wav), only the answers are stored
    sd_spk = {**sd, **synthetic_speaker_state_dict(cfg, seed=WEIGHT_SEED)}
    spk = ReferenceSpeakerTokenizer(cfg).load_checkpoint(sd_spk)
    for name, B, seconds, seed in SPEAKER_CASES:
        wav = synthetic_ref_wav(cfg, B, seconds, seed)
        tokens = spk.tokenize(wav)
        o_tokens, margin = O.tokenize_speaker(sd_spk, cfg, wav)
        mel = spk.mel_transformer(wav.unsqueeze(1)).squeeze(1)
        np.savez_compressed(os.path.join(out_dir, f"speaker_{name}.npz"), weight_seed=np.int64(WEIGHT_SEED),
                            wav_seed=np.int64(seed), batch=np.int64(B), seconds=np.float64(seconds),
                            wav_checksum=np.float64(wav.double().sum().item()), global_tokens=tokens.numpy(),
                            margin=margin.numpy(), mel_first4=mel[:, :, :4].numpy())
        print("speaker", name, tuple(tokens.shape), "oracle equal", bool(torch.equal(tokens, o_tokens)),
              "min margin", float(margin.min()))


if __name__ == "__main__":
    main()
