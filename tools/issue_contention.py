"""Does ALU work of the epilogue warps (no memory traffic) slow the tcgen05 main loop?  EPI_BUSY: accumulators discarded,
then n x 64 dependent FMAs per epilogue thread and tile.  python tools/issue_contention.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import _lib

lib = _lib.lib()
M, N, K = 24 * 8192, 5632, 1056
st = torch.cuda.current_stream().cuda_stream
A = (torch.randn(M, K, device="cuda") * 0.5).half()
W = (torch.randn(N, K, device="cuda") * 0.05).half()
out = torch.empty(M, 8, device="cuda", dtype=torch.float16)


def run(epi, reps=120):
    for _ in range(10):
        _lib.check(lib.swb200_gemm(epi, 3, 1, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), 8, M, N, K, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(lib.swb200_gemm(epi, 3, 1, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), 8, M, N, K, st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


base = run(6)
print(f"w1 shape, main loop only: {base:8.1f} us")
for n in (0, 16, 32):
    t = run(1000 + n)
    print(f"  + {n * 64:5d} dependent FMAs per epilogue thread and tile (a warp issues every 4th cycle): {t:8.1f} us ({100 * (t / base - 1):+5.1f} %)")
for n in (16, 32, 64, 96):
    t = run(2000 + n)
    print(f"  + {n * 64:5d} independent FMAs per epilogue thread and tile (a warp issues every cycle):  {t:8.1f} us ({100 * (t / base - 1):+5.1f} %)")
print(f"main loop only again: {run(6):8.1f} us")
