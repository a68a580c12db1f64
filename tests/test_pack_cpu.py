"""CPU (no GPU): the C-ABI library loads, exports every declared symbol, and its host-side weight
re-layout (tap tables + K-major bf16 hi/lo planes) reproduces torch's Conv1d / ConvTranspose1d."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from spark_tts_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sparkcodec.h")).read()
    declared = set(re.findall(r"SPARKCODEC_API\s+(?:int|const char\*)\s+(sparkcodec_\w+)\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sparkcodec_abi_version() == 1


def _bf16_bits_to_f32(a: np.ndarray) -> np.ndarray:
    return (a.astype(np.uint32) << 16).view(np.float32)


def _emulate(x: torch.Tensor, pk: dict, c_out: int) -> torch.Tensor:
    """x (B, L, C_in) -> (B, L*n_phase, C_out) using only the packed weights + tap table."""
    w = torch.from_numpy(_bf16_bits_to_f32(pk["w_hi"]) + _bf16_bits_to_f32(pk["w_lo"])).double()
    B, L, c_in = x.shape
    out = torch.zeros(B, L, pk["n_total"], dtype=torch.float64)
    xd = x.double()
    for r in range(pk["n_phase"]):
        rows = slice(r * c_out, (r + 1) * c_out)
        for m in range(int(pk["ntaps"][r])):
            sh = int(pk["shifts"][r, m])
            xs = torch.zeros_like(xd)
            lo, hi = max(0, -sh), min(L, L - sh)
            if hi > lo:
                xs[:, lo:hi] = xd[:, lo + sh:hi + sh]
            out[:, :, rows] += xs @ w[rows, m * c_in:(m + 1) * c_in].T
    return out.reshape(B, L * pk["n_phase"], c_out)


@pytest.mark.parametrize("k,dil", [(7, 1), (7, 3), (7, 9), (1, 1)])
def test_pack_conv1d_matches_torch(k, dil):
    g = torch.Generator().manual_seed(k * 10 + dil)
    w = torch.randn(24, 16, k, generator=g)
    x = torch.randn(2, 40, 16, generator=g)
    pk = ops.pack_conv(w, transposed=False, param=dil)
    assert pk["n_phase"] == 1 and pk["kt"] == k and pk["n_total"] == 24
    ref = F.conv1d(x.transpose(1, 2).double(), w.double(), dilation=dil, padding=(k - 1) // 2 * dil).transpose(1, 2)
    got = _emulate(x, pk, 24)
    # hi+lo planes keep ~16 mantissa bits of every weight
    assert (got - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


@pytest.mark.parametrize("k,s", [(16, 8), (11, 5), (8, 4), (4, 2)])
def test_pack_conv_transpose1d_matches_torch(k, s):
    g = torch.Generator().manual_seed(k * 10 + s)
    w = torch.randn(16, 8, k, generator=g)           # (C_in, C_out, k)
    x = torch.randn(2, 23, 16, generator=g)
    pk = ops.pack_conv(w, transposed=True, param=s)
    assert pk["n_phase"] == s and pk["n_total"] == s * 8
    ref = F.conv_transpose1d(x.transpose(1, 2).double(), w.double(), stride=s, padding=(k - s) // 2).transpose(1, 2)
    got = _emulate(x, pk, 8)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


def test_split_planes_are_error_compensated():
    w = torch.randn(32, 32, 1, generator=torch.Generator().manual_seed(3))
    pk = ops.pack_conv(w, transposed=False, param=1)
    hi, lo = _bf16_bits_to_f32(pk["w_hi"]), _bf16_bits_to_f32(pk["w_lo"])
    ref = w[:, :, 0].numpy()
    assert np.abs(hi - ref).max() > 1e-4                      # bf16 alone is coarse ...
    assert np.abs(hi + lo - ref).max() <= np.abs(ref).max() * 2.0 ** -16   # ... hi+lo is not


def test_product_path_has_no_cpu_fallback(cfg, state_dict):
    from spark_tts_b200 import BiCodec
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = BiCodec.from_state_dict(cfg, state_dict)
    sem = torch.zeros(1, 4, dtype=torch.int64)
    glob = torch.zeros(1, 1, cfg.token_num, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        m.detokenize(sem, glob)
    with pytest.raises(RuntimeError):
        m.to("cpu")


def test_header_is_plain_c_and_error_paths_answer_without_a_gpu(tmp_path):
    """include/sparkcodec.h compiles as C99 (no C++ / CUDA / torch types in the ABI), and a C caller linked against
    libsparkcodec.so gets error CODES + messages, not crashes, for calls that cannot succeed on this machine."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "sparkcodec.h"
int main(void) {
  sparkcodec_config cfg;
  memset(&cfg, 0, sizeof cfg);
  sparkcodec_handle* h = NULL;
  if (sparkcodec_abi_version() != SPARKCODEC_ABI_VERSION) return 1;
  if (sparkcodec_create(NULL, 0, &h) != SPARKCODEC_EINVAL) return 2;          /* null config */
  if (sparkcodec_create(&cfg, 0, &h) != SPARKCODEC_EINVAL) return 3;          /* fsq_num_levels == 0 */
  if (strlen(sparkcodec_last_error()) == 0) return 4;
  if (sparkcodec_finalize(NULL) != SPARKCODEC_EINVAL) return 5;
  size_t n = 0;
  if (sparkcodec_workspace_bytes(NULL, 1, 1, &n) != SPARKCODEC_EINVAL) return 6;
  if (sparkcodec_extract_codes(NULL, 7, 1, 1, 0, 8192, 0, 4096, NULL, NULL, NULL, 32, NULL, NULL) != SPARKCODEC_EINVAL) return 7;
  if (sparkcodec_destroy(NULL) != SPARKCODEC_OK) return 8;
  printf("abi ok\n");
  return 0;
}
''')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "spark-tts_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "sparkcodec.h")], check=True)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), "-L", libdir,
                    "-lsparkcodec", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "abi ok" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_two_term_weight_planes_follow_their_definition():
    """OPFMT_F16F8 weight planes (csrc/common.cuh, csrc/pack.cpp): fp16(W), and per group of 32 K values the 64 bytes
    [e5m2(fp16(W) * 2^-4) x 32 | e5m2((W - fp16(W)) * 2^8) x 32] -- checked against torch's own fp16 / float8_e5m2
    round-to-nearest conversions, plus the error budget the two-term product relies on."""
    from spark_tts_b200 import ops
    g = torch.Generator().manual_seed(5)
    w = torch.randn(96, 64, 7, generator=g) / (64 * 7) ** 0.5
    pk = ops.pack_conv_f16f8(w, transposed=False, param=3)
    K = pk["kt"] * pk["c_in"]
    wk = w.permute(0, 2, 1).reshape(96, K).contiguous()              # (n, tap * c_in + c): the packed K order
    h16 = torch.from_numpy(pk["w_h16"].astype(np.int16)).view(torch.float16)
    assert torch.equal(h16, wk.to(torch.float16))
    p8 = torch.from_numpy(pk["w_p8"]).reshape(96, K // 32, 2, 32)
    hi8 = p8[:, :, 0, :].reshape(96, K).contiguous().view(torch.float8_e5m2)
    lo8 = p8[:, :, 1, :].reshape(96, K).contiguous().view(torch.float8_e5m2)
    hf = h16.float()
    assert torch.equal(hi8.view(torch.uint8), (hf * 2.0 ** -4).to(torch.float8_e5m2).view(torch.uint8))
    assert torch.equal(lo8.view(torch.uint8), ((wk - hf) * 2.0 ** 8).to(torch.float8_e5m2).view(torch.uint8))
    # what the tensor cores see of a weight: hi exactly, hi and lo again with 3 significant bits
    assert ((hi8.float() * 2.0 ** 4 - hf).abs() <= hf.abs() * 2.0 ** -3 + 2.0 ** -13).all()
    assert ((lo8.float() * 2.0 ** -8 - (wk - hf)).abs() <= (wk - hf).abs() * 2.0 ** -3 + 2.0 ** -25).all()
    with pytest.raises(ValueError):
        ops.pack_conv_f16f8(torch.randn(8, 24, 1), transposed=False, param=1)   # K = 24 is not a multiple of 32
