"""CPU: the oracle's restatement of the semantic half of BiCodec.tokenize (bicodec.py:151-169: encoder ->
quantizer.tokenize) reproduces the indices the reference's own Encoder + FactorizedVectorQuantize produced
(tests/golden/tokenize_*.npz, written by tests/golden/make_golden.py in the build container)."""
import numpy as np
import pytest
import torch

from conftest import tokenize_golden_cases
from oracle import bicodec_oracle as O
from spark_tts_b200.synthetic import synthetic_features


@pytest.mark.parametrize("path", tokenize_golden_cases(), ids=lambda p: p.split("tokenize_")[-1][:-4])
def test_oracle_tokenize_matches_reference_golden(path, cfg, state_dict_with_encoder):
    g = np.load(path)
    feat = synthetic_features(cfg, int(g["batch"]), int(g["frames"]), int(g["feat_seed"]))
    assert abs(feat.double().sum().item() - float(g["feat_checksum"])) < 1e-6   # same features as the generator saw
    idx, margin = O.tokenize_semantic(state_dict_with_encoder, cfg, feat)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (int(g["batch"]), int(g["frames"]))
    assert np.array_equal(idx.numpy(), g["semantic_tokens"])                    # index work: bit exact
    assert np.allclose(margin.numpy(), g["margin"], atol=1e-6)


def test_goldens_exist():
    assert len(tokenize_golden_cases()) >= 4


def test_vq_tokenize_inverts_detokenize_codes(cfg, state_dict_with_encoder):
    """Nearest-code search returns k when the latent IS code k (the quantizer's own fixed point:
    factorized_vector_quantize.py:169-187 applied to its codebook rows); checked through the search only."""
    sd = state_dict_with_encoder
    cb = sd["quantizer.codebook.weight"]
    k = torch.arange(0, cfg.codebook_size, 37)
    enc = torch.nn.functional.normalize(cb[k])
    cbn = torch.nn.functional.normalize(cb)
    dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cbn.t() + cbn.pow(2).sum(1, keepdim=True).t()
    assert torch.equal((-dist).max(1)[1], k)
