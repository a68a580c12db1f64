"""Stand-alone entry points of libsparkcodec used by the tests: the dense convolution through the same
packing + kernels the model uses, and the host-side weight re-layout (no GPU needed)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib


def conv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], *, transposed: bool, param: int,
         act: str = "none", alpha: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         precision: str = "fp32", impl: str = "tc") -> torch.Tensor:
    """x (B, L, C_in) fp32 CUDA channels-last -> (B, L_out, C_out).  Conv1d: weight (C_out, C_in, k), param =
    dilation, 'same' padding.  ConvTranspose1d: weight (C_in, C_out, k), param = stride, padding (k-stride)//2."""
    lib = _lib.load()
    if x.device.type != "cuda" or x.dtype != torch.float32:
        raise ValueError("x must be a float32 CUDA tensor")
    x = x.contiguous()
    B, L, _ = x.shape
    w = weight.detach().to("cpu", torch.float32).contiguous()
    c_out = w.shape[1] if transposed else w.shape[0]
    L_out = L * param if transposed else L
    y = torch.empty((B, L_out, c_out), dtype=torch.float32, device=x.device)
    b = bias.detach().to("cpu", torch.float32).contiguous() if bias is not None else None
    a = alpha.detach().to("cpu", torch.float32).contiguous().view(-1) if alpha is not None else None
    r = residual.contiguous() if residual is not None else None
    wshape = (C.c_int64 * 3)(*w.shape)
    actc = {"none": _lib.ACT_NONE, "gelu": _lib.ACT_GELU, "snake": _lib.ACT_SNAKE}[act]
    _lib.check(lib.sparkcodec_op_conv(
        x.device.index or 0, 1 if transposed else 0, C.c_void_p(w.data_ptr()), wshape,
        C.c_void_p(b.data_ptr()) if b is not None else None, int(param), B, L, C.c_void_p(x.data_ptr()),
        C.c_void_p(y.data_ptr()), C.c_void_p(r.data_ptr()) if r is not None else None, actc,
        C.c_void_p(a.data_ptr()) if a is not None else None, _lib.PRECISIONS[precision],
        _lib.IMPL_TC if impl == "tc" else _lib.IMPL_SIMT,
        C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return y


def pack_conv(weight: torch.Tensor, *, transposed: bool, param: int):
    """Host-side re-layout (no GPU): returns dict(w_hi, w_lo uint16 (n_total, kt*c_in), shifts (n_phase, kt),
    ntaps (n_phase,), kt, n_phase, n_total)."""
    lib = _lib.load()
    w = weight.detach().to("cpu", torch.float32).contiguous()
    wshape = (C.c_int64 * 3)(*w.shape)
    kt, n_phase, n_total = C.c_int32(), C.c_int32(), C.c_int32()
    kind = 1 if transposed else 0
    _lib.check(lib.sparkcodec_pack_conv(kind, C.c_void_p(w.data_ptr()), wshape, int(param), None, None, 0, None, None,
                                        C.byref(kt), C.byref(n_phase), C.byref(n_total)))
    c_in = w.shape[0] if transposed else w.shape[1]
    n = n_total.value * kt.value * c_in
    hi = np.empty(n, dtype=np.uint16)
    lo = np.empty(n, dtype=np.uint16)
    shifts = np.empty(n_phase.value * kt.value, dtype=np.int32)
    ntaps = np.empty(n_phase.value, dtype=np.int32)
    _lib.check(lib.sparkcodec_pack_conv(kind, C.c_void_p(w.data_ptr()), wshape, int(param),
                                        hi.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), n,
                                        shifts.ctypes.data_as(C.c_void_p), ntaps.ctypes.data_as(C.c_void_p),
                                        C.byref(kt), C.byref(n_phase), C.byref(n_total)))
    return dict(w_hi=hi.reshape(n_total.value, kt.value * c_in), w_lo=lo.reshape(n_total.value, kt.value * c_in),
                shifts=shifts.reshape(n_phase.value, kt.value), ntaps=ntaps, kt=kt.value, n_phase=n_phase.value,
                n_total=n_total.value, c_in=c_in)


def pack_conv_f16f8(weight: torch.Tensor, *, transposed: bool, param: int):
    """Host-side weight planes of the two-term fp32 mode (no GPU): dict(w_h16 uint16 (n_total, K) = fp16 bits,
    w_p8 uint8 (n_total, 2 K): per group of 32 K values 64 bytes [e5m2(fp16(W) 2^-4) x 32 | e5m2((W - fp16(W)) 2^8) x 32])."""
    lib = _lib.load()
    meta = pack_conv(weight, transposed=transposed, param=param)
    w = weight.detach().to("cpu", torch.float32).contiguous()
    wshape = (C.c_int64 * 3)(*w.shape)
    n_total, K = meta["n_total"], meta["kt"] * meta["c_in"]
    h16 = np.empty(n_total * K, dtype=np.uint16)
    p8 = np.empty(n_total * K, dtype=np.uint16)
    _lib.check(lib.sparkcodec_pack_conv_f16f8(1 if transposed else 0, C.c_void_p(w.data_ptr()), wshape, int(param),
                                              h16.ctypes.data_as(C.c_void_p), p8.ctypes.data_as(C.c_void_p), n_total * K))
    return dict(w_h16=h16.reshape(n_total, K), w_p8=p8.view(np.uint8).reshape(n_total, 2 * K), **{k: meta[k] for k in ("kt", "c_in", "n_total")})
