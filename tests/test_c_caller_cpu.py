"""CPU: the plain-C caller of the ABI (tests/c_caller/detok_main.c: no Python, no torch) compiles as C99 against
include/sparkcodec.h and links against libsparkcodec.so; run without a GPU it fails loudly through the C error path
instead of producing a waveform.  The GPU run of the same program is tests/test_gpu_c_caller.py."""
import os
import subprocess

import pytest

from c_caller_util import build_c_caller, write_model_bin, write_tokens_bin


def test_c_caller_builds_and_fails_loudly_without_gpu(tmp_path, cfg, state_dict):
    import torch
    exe = build_c_caller(tmp_path)
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run is covered by tests/test_gpu_c_caller.py")
    small = {k: v for k, v in state_dict.items() if k.startswith("quantizer.")}
    write_model_bin(tmp_path / "model.bin", cfg, small)
    from spark_tts_b200.synthetic import synthetic_tokens
    sem, glob = synthetic_tokens(cfg, 1, 4, 3)
    write_tokens_bin(tmp_path / "tokens.bin", sem, glob)
    r = subprocess.run([exe, str(tmp_path / "model.bin"), str(tmp_path / "tokens.bin"), str(tmp_path / "out.f32")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and not os.path.exists(tmp_path / "out.f32")
    assert "sparkcodec_" in r.stderr          # the failing call and the library's message are reported
