"""Stand-alone timing of one dense-contraction shape through sparkcodec_op_conv (for ncu captures).
python tools/bench_op.py <c_in> <c_out> <k> <param> <transposed 0/1> <B> <L> <act> <residual 0/1> <precision> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spark_tts_b200 import ops

c_in, c_out, k, param, tr, B, L = map(int, sys.argv[1:8])
act, res, prec = sys.argv[8], int(sys.argv[9]), sys.argv[10]
reps = int(sys.argv[11]) if len(sys.argv) > 11 else 3
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = (torch.randn(c_in, c_out, k, generator=g) if tr else torch.randn(c_out, c_in, k, generator=g)) / (c_in * k) ** 0.5
b = torch.randn(c_out, generator=g) * 0.1
alpha = torch.rand(c_out, generator=g) + 0.5
x = torch.randn(B, L, c_in, device=dev)
r = torch.randn(B, L * (param if tr else 1), c_out, device=dev) if res else None
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = ops.conv(x, w, b, transposed=bool(tr), param=param, act=act, alpha=alpha, residual=r, precision=prec)
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {i}: {e0.elapsed_time(e1):.3f} ms (includes split/merge helper kernels and weight upload)")
