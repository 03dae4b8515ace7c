"""Where the warps of the 256x352 GEMM spend their cycles (profiling build: -DSWB_PROFILE_EPILOGUES, SWB_LIB=...):
    python tools/gemm_roles.py [chunk] [shape ...]          shapes: qkv wo w1 w2 swiglu qkvfused"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import _lib


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    names = sys.argv[2:] or ["w1", "swiglu", "qkvfused", "wo", "w2"]
    lib = _lib.lib()
    prof = lib.swb200_debug_gemm_prof
    prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    M = chunk * 8192
    st = torch.cuda.current_stream().cuda_stream
    shapes = {"qkv": (3168, 1056), "wo": (1056, 1056), "w1": (5632, 1056), "w2": (1056, 2816), "swiglu": (5632, 1056),
              "qkvfused": (3168, 1056), "lnres_wo": (1056, 1056), "lnres_w2": (1056, 2816), "lnres_wo_pair": (1056, 1056),
              "lnres_w2_pair": (1056, 2816)}
    for name in names:
        N, K = shapes[name]
        A = (torch.randn(M, K, device="cuda") * 0.5).half()
        W = (torch.randn(N, K, device="cuda") * 0.05).half()
        if name == "swiglu":
            out = torch.empty(M, N // 2, device="cuda", dtype=torch.float16)
            fn = lambda: _lib.check(lib.swb200_gemm_swiglu(3, 1, A.data_ptr(), K, W.data_ptr(), out.data_ptr(), M, 1056, N // 2, st))
        elif name == "qkvfused":
            qs = torch.full((12,), 10.0, device="cuda")
            out = torch.empty(3 * 12 * M * 96, device="cuda", dtype=torch.float16)
            fn = lambda: _lib.check(lib.swb200_gemm_qkv(3, 1, 1, A.data_ptr(), K, W.data_ptr(), qs.data_ptr(), out.data_ptr(), M, 1056, 12, st))
        elif name.startswith("lnres"):                 # lnres_wo / lnres_w2 [_pair]: fused LayerNorm + residual epilogue
            fmt = 1 if name.endswith("_pair") else 3
            T = 8192
            xhl = torch.randn(M, 2 * N, device="cuda").half()
            gain = torch.randn(M // T, N, device="cuda")
            bias = torch.randn(M // T, N, device="cuda")
            ws = torch.empty(lib.swb200_ln_workspace_bytes(M, N) + 256, dtype=torch.uint8, device="cuda")
            ws_ptr = (ws.data_ptr() + 255) // 256 * 256
            gen = [0]
            out = xhl

            def fn():
                _lib.check(lib.swb200_gemm_ln_residual(3, fmt, A.data_ptr(), K, W.data_ptr(), K, xhl.data_ptr(), gain.data_ptr(),
                                                       bias.data_ptr(), M, N, T, ws_ptr, gen[0], st))
                gen[0] += 1
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.float16)
            fn = lambda: _lib.check(lib.swb200_gemm(1, 3, 1, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, st))
        for _ in range(200):                      # reach the power-capped steady state
            fn()
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 24)()
        prof(buf, 1)
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        prof(buf, 0)
        c = [int(x) for x in buf]
        us = e0.elapsed_time(e1) * 1e3 / reps
        tiles = max(1, c[3])
        wt = max(1, c[8])
        N = min(N, 10 ** 9)
        ideal = 2.0 * 256 * 352 * K / 2 / 8192            # tensor-pipe cycles of one tile at 8192 FLOP/clk/SM
        print(f"{name:13s} {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TF/s | issuer per tile: {c[0] / tiles:7.0f} cyc (ideal {ideal:.0f}), "
              f"waiting for data {c[1] / tiles:6.0f}, for an accumulator {c[2] / tiles:6.0f} | producer: waits {100.0 * c[4] / max(1, c[10]):4.1f} % | "
              f"epilogue warp per tile: waits {c[5] / wt:6.0f}, drain {c[6] / wt:5.0f}, rest {c[7] / wt:6.0f} (of it store-buffer waits {c[9] / wt:5.0f})")
        if c[11]:
            print(f"{'':13s} fused LayerNorm epilogue per warp-tile: publish (store + fence + atomic) {c[11] / wt:6.0f}, x loads + poll "
                  f"{c[12] / wt:6.0f}, acquire fence + partials + merge {c[13] / wt:6.0f}, apply {(c[7] - c[11] - c[12] - c[13]) / wt:6.0f}")
        if c[14]:
            print(f"{'':13s} x by TMA: waiting for the 64-column tiles {c[14] / wt:6.0f}, for the tails {c[15] / wt:6.0f}; the four "
                  f"in-place updates + stores, measured directly: {c[16] / wt:6.0f}")
        del A, W, out


if __name__ == "__main__":
    main()
