"""Every scheduling mode of the tcgen05 kernels computes the same waveform.

The modes are process-wide switches read once from the environment (SPARKCODEC_PAIR: CTA pairs in the conv
kernel never / heuristic / always; SPARKCODEC_CLUSTER: single CTAs / weight multicast / CTA pairs in the fused
ResidualUnit kernel; SPARKCODEC_HALO_STAGES: halo tiles in flight; SPARKCODEC_PDL: programmatic dependent launch of the
pass's kernels on / off; SPARKCODEC_SMALL_N: narrow N tiles for few-tile launches on / off), so each one runs in its own interpreter and
writes its waveform to a file; the test compares them with the default configuration."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from spark_tts_b200 import BiCodec, BiCodecConfig
from spark_tts_b200.synthetic import synthetic_state_dict, synthetic_tokens
cfg = BiCodecConfig()
dev = torch.device("cuda:0")
m = BiCodec.from_state_dict(cfg, synthetic_state_dict(cfg, 0), device=dev, precision={prec!r})
sem, glob = synthetic_tokens(cfg, 3, 45, 2024)          # 3 utterances: an odd number of M tiles at every stage
wav = m.detokenize(sem.to(dev), glob.to(dev)).cpu().numpy()
np.save({out!r}, wav)
"""


def _run(tmp_path, name, prec, env):
    out = str(tmp_path / f"{name}_{prec}.npy")
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT, prec=prec, out=out)], check=True, env=e, timeout=600)
    return np.load(out)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_all_kernel_modes_agree(tmp_path, prec):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle.bicodec_oracle import snr_db
    ref = _run(tmp_path, "default", prec, {})
    assert np.isfinite(ref).all()
    modes = {
        "single_ctas": {"SPARKCODEC_PAIR": "0", "SPARKCODEC_CLUSTER": "1"},
        "pairs_everywhere": {"SPARKCODEC_PAIR": "2", "SPARKCODEC_CLUSTER": "3"},
        "multicast_fused": {"SPARKCODEC_CLUSTER": "2"},
        "three_halo_tiles": {"SPARKCODEC_HALO_STAGES": "3"},
        # no programmatic dependent launch, no narrow N tiles for few-tile launches (one run for both: each is
        # bit-identical to the default on its own -- profiles/r2_pdl_ab.txt, r2_small_n_ab.txt)
        "plain_launches": {"SPARKCODEC_PDL": "0", "SPARKCODEC_SMALL_N": "0"},
    }
    # fp32 mode: every mode feeds the tensor cores the same operands in the same order -> >= 100 dB;
    # bf16 mode: last-bit differences of fp32 sums can flip a bf16 rounding (see test_gpu_parity.py) -> >= 60 dB
    floor = 100.0 if prec == "fp32" else 60.0
    for name, env in modes.items():
        got = _run(tmp_path, name, prec, env)
        snr = snr_db(torch.from_numpy(ref), torch.from_numpy(got))
        assert snr >= floor, f"{name}: {snr:.1f} dB vs default"
    if prec == "fp32":
        # the other operand split of the fp32 mode (three bf16 products instead of fp16 + two e5m2 products) is
        # different arithmetic: the two waveforms agree to the accuracy of the coarser one (~73 dB vs the oracle)
        got = _run(tmp_path, "three_term_split", prec, {"SPARKCODEC_FP32_TERMS": "3"})
        snr = snr_db(torch.from_numpy(ref), torch.from_numpy(got))
        assert snr >= 66.0, f"three_term_split: {snr:.1f} dB vs default"
