"""A/B of the SwiGLU GEMM between builds of the library inside ONE process (power-capped GPUs drift between runs):
    python tools/swiglu_ab.py libA.so libB.so [...]"""
import ctypes as C
import sys

import torch

libs = []
for path in sys.argv[1:]:
    l = C.CDLL(path)
    l.swb200_gemm_swiglu.restype = C.c_int
    l.swb200_gemm_swiglu.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.swb200_gemm_qkv.restype = C.c_int
    l.swb200_gemm_qkv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    libs.append((path.split("/")[-1], l))
M, N, K = 8 * 8192, 5632, 1056
st = torch.cuda.current_stream().cuda_stream
A = (torch.randn(M, K, device="cuda") * 0.5).half()
W = (torch.randn(N, K, device="cuda") * 0.05).half()
Wq = (torch.randn(3168, K, device="cuda") * 0.05).half()
qs = torch.full((12,), 10.0, device="cuda")
o2 = torch.empty(M, N // 2, device="cuda", dtype=torch.float16)
o3 = torch.empty(3 * 12 * M * 96, device="cuda", dtype=torch.float16)


def run(l, which, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if which == "swiglu":
            rc = l.swb200_gemm_swiglu(3, 1, A.data_ptr(), K, W.data_ptr(), o2.data_ptr(), M, 1056, N // 2, st)
        else:
            rc = l.swb200_gemm_qkv(3, 1, 1, A.data_ptr(), K, Wq.data_ptr(), qs.data_ptr(), o3.data_ptr(), M, 1056, 12, st)
        assert rc == 0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for which in ("swiglu", "qkv"):
    for name, l in libs:
        run(l, which, 100)
    for rnd in range(3):
        print(which, "  ".join(f"{name}: {run(l, which, 400):7.1f} us" for name, l in libs))
