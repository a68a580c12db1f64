"""ctypes binding of libsparkcodec.so (include/sparkcodec.h).  Fails loudly when the CUDA library is
missing or the call cannot run on a GPU: there is NO CPU / PyTorch fallback on the product path."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# (SPARKCODEC_LIB: A/B timing of two builds of the same ABI on one box; the default is the in-tree library)
LIB_PATH = os.environ.get("SPARKCODEC_LIB") or os.path.join(_HERE, "libsparkcodec.so")

OK, EINVAL, EINDEX, ECUDA, ESTATE, ENOMEM, EMISSING = 0, -1, -2, -3, -4, -5, -6
I32, I64 = 0, 1
PREC_FP32, PREC_BF16, PREC_FP32X3 = 0, 1, 2
IMPL_TC, IMPL_SIMT, IMPL_TC_UNFUSED = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_SNAKE = 0, 1, 2

# "fp32x3": the fp32 mode with the three-term bf16 split whatever the process default (the encode side uses it)
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp32x3": PREC_FP32X3}


class SparkCodecConfig(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32), ("codebook_size", C.c_int32), ("codebook_dim", C.c_int32),
        ("fsq_num_levels", C.c_int32), ("fsq_levels", C.c_int32 * 8),
        ("token_num", C.c_int32), ("latent_dim", C.c_int32),
        ("vocos_dim", C.c_int32), ("vocos_intermediate_dim", C.c_int32), ("vocos_num_layers", C.c_int32),
        ("downsample_layers", C.c_int32), ("num_downsample", C.c_int32),
        ("dec_channels", C.c_int32), ("num_upsample", C.c_int32),
        ("rates", C.c_int32 * 8), ("kernel_sizes", C.c_int32 * 8),
    ]


# name -> (restype, argtypes): every symbol include/sparkcodec.h declares
_H = C.c_void_p
SIGNATURES = {
    "sparkcodec_last_error": (C.c_char_p, []),
    "sparkcodec_abi_version": (C.c_int, []),
    "sparkcodec_create": (C.c_int, [C.POINTER(SparkCodecConfig), C.c_int, C.POINTER(_H)]),
    "sparkcodec_destroy": (C.c_int, [_H]),
    "sparkcodec_set_tensor": (C.c_int, [_H, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "sparkcodec_finalize": (C.c_int, [_H]),
    "sparkcodec_workspace_bytes": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "sparkcodec_detokenize": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "sparkcodec_prenet": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "sparkcodec_wavegen": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                     C.c_void_p, C.c_void_p]),
    "sparkcodec_wavegen_stage": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    "sparkcodec_wavegen_staged": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                            C.c_void_p]),
    "sparkcodec_halo_frames": (C.c_int, [_H, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sparkcodec_check_tokens": (C.c_int, [_H, C.c_void_p]),
    "sparkcodec_extract_codes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sparkcodec_tokenize_semantic": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "sparkcodec_set_mel_params": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]),
    "sparkcodec_speaker_workspace_bytes": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "sparkcodec_tokenize_speaker": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "sparkcodec_tokenize_speaker_tap": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                                  C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.c_void_p]),
    "sparkcodec_set_impl": (C.c_int, [_H, C.c_int]),
    "sparkcodec_detokenize_tap": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_size_t, C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_int64), C.c_void_p]),
    "sparkcodec_op_conv": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_int, C.c_void_p]),
    "sparkcodec_pack_conv": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sparkcodec_pack_conv_f16f8": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_size_t]),
    "sparkcodec_launch_count": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "sparkcodec_fp32_terms": (C.c_int, []),
    "sparkcodec_tile_width": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "sparkcodec_profile": (C.c_int, [_H, C.c_int]),
    "sparkcodec_profile_read": (C.c_int, [_H, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
}

_lib: Optional[C.CDLL] = None


class SparkCodecError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen libsparkcodec.so (built by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SparkCodecError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU or PyTorch fallback for the detokenize path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = ABI mismatch, deliberately loud
        fn.restype = res
        fn.argtypes = args
    if lib.sparkcodec_abi_version() != 1:
        raise SparkCodecError("libsparkcodec ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = load().sparkcodec_last_error().decode("utf-8", "replace")
    if rc == EINDEX:
        raise IndexError(msg)
    if rc in (EINVAL, EMISSING):
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise SparkCodecError(msg)


def make_config(cfg) -> SparkCodecConfig:
    c = SparkCodecConfig()
    c.d_model, c.codebook_size, c.codebook_dim = cfg.d_model, cfg.codebook_size, cfg.codebook_dim
    c.fsq_num_levels = len(cfg.fsq_levels)
    for i, v in enumerate(cfg.fsq_levels):
        c.fsq_levels[i] = v
    c.token_num, c.latent_dim = cfg.token_num, cfg.latent_dim
    c.vocos_dim, c.vocos_intermediate_dim = cfg.vocos_dim, cfg.vocos_intermediate_dim
    c.vocos_num_layers, c.downsample_layers = cfg.vocos_num_layers, cfg.downsample_layers
    c.num_downsample = len(cfg.sample_ratios)
    c.dec_channels, c.num_upsample = cfg.dec_channels, len(cfg.rates)
    for i, (r, k) in enumerate(zip(cfg.rates, cfg.kernel_sizes)):
        c.rates[i], c.kernel_sizes[i] = r, k
    return c
