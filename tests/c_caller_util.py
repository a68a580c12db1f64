"""Helpers for the plain-C caller tests: build tests/c_caller/detok_main.c with gcc, write its input files."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_c_caller(tmp_path) -> str:
    exe = os.path.join(str(tmp_path), "detok_main")
    libdir = os.path.join(ROOT, "spark-tts_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(CUDA_HOME, "include"), os.path.join(ROOT, "tests", "c_caller", "detok_main.c"),
           "-o", exe, "-L", libdir, "-lsparkcodec", "-L", os.path.join(CUDA_HOME, "lib64"), "-lcudart",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(CUDA_HOME, 'lib64')}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_model_bin(path, cfg, state_dict) -> None:
    """sparkcodec_config struct, then {u32 key_len, key, u32 ndim, i64 dims, f32 data} per tensor."""
    import ctypes

    from spark_tts_b200 import _lib
    with open(path, "wb") as f:
        f.write(bytes(memoryview(ctypes.string_at(ctypes.byref(_lib.make_config(cfg)),
                                                  ctypes.sizeof(_lib.SparkCodecConfig)))))
        for key, t in state_dict.items():
            a = np.ascontiguousarray(t.detach().cpu().float().numpy())
            k = key.encode()
            f.write(struct.pack("<I", len(k)) + k + struct.pack("<I", a.ndim) +
                    struct.pack(f"<{a.ndim}q", *a.shape))
            f.write(a.tobytes())


def write_tokens_bin(path, semantic, global_tokens) -> None:
    sem = np.ascontiguousarray(semantic.cpu().numpy().astype(np.int64))
    glob = np.ascontiguousarray(global_tokens.reshape(sem.shape[0], -1).cpu().numpy().astype(np.int32))
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", sem.shape[0], sem.shape[1]))
        f.write(sem.tobytes())
        f.write(glob.tobytes())
