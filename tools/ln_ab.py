"""In-process A/B of the stand-alone LayerNorm-modulate-residual kernel between builds, pair and single-value residual stream:
    python tools/ln_ab.py libA.so libB.so ..."""
import ctypes as C
import sys

import torch

libs = []
for path in sys.argv[1:]:
    l = C.CDLL(path)
    l.swb200_ln_mod_residual.restype = C.c_int
    l.swb200_ln_mod_residual.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    libs.append((path.split("/")[-1], l))
B, T, D = 24, 8192, 1056
M = B * T
st = torch.cuda.current_stream().cuda_stream
branch = torch.randn(M, D, device="cuda").half()
xhl = (torch.randn(M, 2 * D, device="cuda") * 0.5).half()
gain = torch.randn(B, D, device="cuda") * 0.01
bias = torch.randn(B, D, device="cuda") * 0.01


def run(l, reps, fmt):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        assert l.swb200_ln_mod_residual(branch.data_ptr(), 1, xhl.data_ptr(), gain.data_ptr(), bias.data_ptr(), M, D, T, fmt, st) == 0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for fmt, bpe, label in ((1, 10, "[hi | lo] pair"), (3, 6, "single value")):
    for name, l in libs:
        run(l, 20, fmt)
    for rnd in range(2):
        print(f"{label:15s} " + "  ".join(f"{name}: {run(l, 200, fmt):7.1f} us ({M * D * bpe / run(l, 50, fmt) / 1e6:5.2f} TB/s)" for name, l in libs))
