"""In-process A/B of the fused w2 + LayerNorm GEMM between builds: python tools/lnfused_ab.py libA.so libB.so"""
import ctypes as C
import sys

import torch

libs = []
for path in sys.argv[1:]:
    l = C.CDLL(path)
    l.swb200_gemm_ln_residual.restype = C.c_int
    l.swb200_gemm_ln_residual.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    l.swb200_ln_workspace_bytes.restype = C.c_size_t
    l.swb200_ln_workspace_bytes.argtypes = [C.c_int, C.c_int]
    libs.append((path.split("/")[-1], l))
B, T, D, K = 8, 8192, 1056, 2816
M = B * T
st = torch.cuda.current_stream().cuda_stream
A = (torch.randn(M, K, device="cuda") * 0.5).half()
W = (torch.randn(D, K, device="cuda") * 0.05).half()
xhl = (torch.randn(M, 2 * D, device="cuda") * 0.5).half()
gain, bias = torch.randn(B, D, device="cuda"), torch.randn(B, D, device="cuda")
ws = torch.empty(libs[0][1].swb200_ln_workspace_bytes(M, D) + 256, dtype=torch.uint8, device="cuda")
wsp = (ws.data_ptr() + 255) // 256 * 256


def run(l, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        assert l.swb200_gemm_ln_residual(3, 1, A.data_ptr(), K, W.data_ptr(), K, xhl.data_ptr(), gain.data_ptr(), bias.data_ptr(),
                                         M, D, T, wsp, 0, st) == 0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, l in libs:
    run(l, 50)
for rnd in range(3):
    print("w2 + LN fused  " + "  ".join(f"{name}: {run(l, 300):7.1f} us" for name, l in libs))
