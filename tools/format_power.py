"""fp16 vs bf16 operands at the power cap: same tcgen05 kind::f16 rate, different multiplier width.  In-process A/B of
the w1 (SwiGLU) GEMM, sustained."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import _lib

lib = _lib.lib()
M, N, K = 8 * 8192, 5632, 1056
st = torch.cuda.current_stream().cuda_stream
bufs = {}
for f16, dt in ((1, torch.float16), (0, torch.bfloat16)):
    bufs[f16] = ((torch.randn(M, K, device="cuda") * 0.5).to(dt), (torch.randn(N, K, device="cuda") * 0.05).to(dt),
                 torch.empty(M, N // 2, device="cuda", dtype=dt))


def run(f16, reps):
    A, W, o = bufs[f16]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(lib.swb200_gemm_swiglu(3, f16, A.data_ptr(), K, W.data_ptr(), o.data_ptr(), M, 1056, N // 2, st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


run(1, 100), run(0, 100)
for _ in range(3):
    print(f"w1 GEMM sustained: fp16 operands {run(1, 500):7.1f} us   bf16 operands {run(0, 500):7.1f} us")
