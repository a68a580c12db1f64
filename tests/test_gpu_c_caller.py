"""GPU: a plain C program (tests/c_caller/detok_main.c) drives the whole path through include/sparkcodec.h --
checkpoint tensors by key, finalize, workspace query, detokenize, token check -- with cudaMalloc'd buffers and no
Python in the process.  Its waveform must meet the same gates as the Python surface: the reference golden and the
oracle (fp32: max-abs <= 1e-3 and SNR >= 60 dB), and equal the ctypes path bit for bit."""
import subprocess

import numpy as np
import pytest
import torch

from c_caller_util import build_c_caller, write_model_bin, write_tokens_bin
from conftest import golden_cases
from oracle import bicodec_oracle as O

pytestmark = pytest.mark.gpu


def test_c_program_matches_golden_and_python_surface(tmp_path, cfg, state_dict):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    exe = build_c_caller(tmp_path)
    write_model_bin(tmp_path / "model.bin", cfg, state_dict)
    g = np.load([p for p in golden_cases() if "b3_t16" in p][0])
    sem, glob = torch.from_numpy(g["semantic_tokens"]), torch.from_numpy(g["global_tokens"])
    write_tokens_bin(tmp_path / "tokens.bin", sem, glob)
    r = subprocess.run([exe, str(tmp_path / "model.bin"), str(tmp_path / "tokens.bin"), str(tmp_path / "out.f32")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "launches=" in r.stdout and int(r.stdout.split("launches=")[1].split()[0]) > 0
    ref = torch.from_numpy(g["output_waveform"])                       # (B, 1, hop*T) from the reference modules
    wav = torch.from_numpy(np.fromfile(tmp_path / "out.f32", dtype=np.float32)).reshape(ref.shape)
    assert (wav - ref).abs().max().item() <= 1e-3
    assert O.snr_db(ref, wav) >= 60.0
    # the same library through ctypes/torch gives the same bits
    from spark_tts_b200 import BiCodec
    dev = torch.device("cuda:0")
    m = BiCodec.from_state_dict(cfg, state_dict, device=dev)
    py = m.detokenize(sem.to(dev), glob.to(dev)).cpu()
    assert torch.equal(py, wav)


def test_c_program_reports_bad_token(tmp_path, cfg, state_dict):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spark_tts_b200.synthetic import synthetic_tokens
    exe = build_c_caller(tmp_path)
    write_model_bin(tmp_path / "model.bin", cfg, state_dict)
    sem, glob = synthetic_tokens(cfg, 1, 6, 3)
    sem[0, 2] = cfg.codebook_size + 5
    write_tokens_bin(tmp_path / "tokens.bin", sem, glob)
    r = subprocess.run([exe, str(tmp_path / "model.bin"), str(tmp_path / "tokens.bin"), str(tmp_path / "out.f32")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 2 and "out of range" in r.stderr and "sparkcodec_check_tokens" in r.stderr
