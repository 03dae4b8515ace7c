"""In-process A/B of the tcgen05 window attention between builds (Swift-B shapes, chunk of 24): python tools/attn_ab.py libA.so libB.so"""
import ctypes as C
import sys

import torch

libs = []
for path in sys.argv[1:]:
    l = C.CDLL(path)
    l.swb200_window_attention.restype = C.c_int
    l.swb200_window_attention.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p, C.c_void_p]
    libs.append((path.split("/")[-1], l))
B, gh, gw, H = 24, 64, 128, 12
M = B * gh * gw
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(H, M, 88, device="cuda", generator=g), dim=-1) * 12.0
k = torch.nn.functional.normalize(torch.randn(H, M, 88, device="cuda", generator=g), dim=-1)
v = torch.randn(H, M, 88, device="cuda", generator=g)
qkv = torch.zeros(3, H, M, 96, device="cuda", dtype=torch.float16)
qkv[0, :, :, :88], qkv[1, :, :, :88], qkv[2, :, :, :88] = q.half(), k.half(), v.half()
outs = []
for name, l in libs:
    out = torch.empty(M, H * 88, device="cuda", dtype=torch.float16)
    for shift in (0, 8):
        for _ in range(3):
            assert l.swb200_window_attention(qkv.data_ptr(), out.data_ptr(), B, gh, gw, H, shift, shift, 1, 1, 2, None, st) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            l.swb200_window_attention(qkv.data_ptr(), out.data_ptr(), B, gh, gw, H, shift, shift, 1, 1, 2, None, st)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10
        print(f"{name:24s} shift {shift}: {us:7.1f} us  ({(3 * 96 + 88) * 2 * M * H / us / 1e6:5.2f} TB/s)")
    outs.append(out.float())
if len(outs) > 1:
    d = (outs[0] - outs[1]).norm() / outs[1].norm()
    print(f"relative difference between the first two builds: {d:.3e}")
