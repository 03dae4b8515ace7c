"""Time the four GEMM shapes of one Swift-B layer alone (chunk x 8192 rows) and print TFLOP/s.
    python tools/gemm_bench.py [chunk] [f16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import _lib


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    f16 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cgs = [3, 2, 1] if len(sys.argv) > 3 else [3, 2]
    dt = torch.float16 if f16 else torch.bfloat16
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    M, D, Dff, H = chunk * 8192, 1056, 2816, 12
    x = (torch.randn(M, D, device="cuda") * 0.5).to(dt)
    h = (torch.randn(M, Dff, device="cuda") * 0.5).to(dt)
    wq = (torch.randn(3 * D, D, device="cuda") * 0.02).to(dt)
    wo = (torch.randn(D, D, device="cuda") * 0.02).to(dt)
    w1 = (torch.randn(2 * Dff, D, device="cuda") * 0.02).to(dt)
    w2 = (torch.randn(D, Dff, device="cuda") * 0.02).to(dt)
    qs = torch.full((H,), 10.0, device="cuda")
    qkv = torch.empty(3, H, M, 96, device="cuda", dtype=dt)
    br = torch.empty(M, D, device="cuda", dtype=dt if f16 else torch.float32)
    epi_br = 1 if f16 else 0        # the forward stores the wo / w2 branch in fp16 in fp16 mode
    hb = torch.empty(M, Dff, device="cuda", dtype=dt)
    for cg in cgs:
        cases = {
            "qkv": (lambda: lib.swb200_gemm_qkv(cg, f16, f16, x.data_ptr(), D, wq.data_ptr(), qs.data_ptr(), qkv.data_ptr(), M, D, H, st), 2.0 * M * 3 * D * D),
            "wo": (lambda: lib.swb200_gemm(epi_br, cg, f16, x.data_ptr(), D, wo.data_ptr(), D, br.data_ptr(), D, M, D, D, st), 2.0 * M * D * D),
            "w1": (lambda: lib.swb200_gemm_swiglu(cg, f16, x.data_ptr(), D, w1.data_ptr(), hb.data_ptr(), M, D, Dff, st), 2.0 * M * 2 * Dff * D),
            "w2": (lambda: lib.swb200_gemm(epi_br, cg, f16, h.data_ptr(), Dff, w2.data_ptr(), Dff, br.data_ptr(), D, M, D, Dff, st), 2.0 * M * D * Dff),
        }
        tot_ms, tot_fl = 0.0, 0.0
        for name, (fn, fl) in cases.items():
            ms = min(timeit(fn) for _ in range(3))
            tot_ms += ms
            tot_fl += fl
            print(f"tile={cg} {name:4s} M={M}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
        print(f"tile={cg} layer GEMMs: {tot_ms*1e3:8.1f} us  {tot_fl/tot_ms/1e9:7.1f} TFLOP/s")
    # torch.matmul (cuBLAS) on the same shapes as a yardstick
    for name, (a, w) in {"qkv": (x, wq), "wo": (x, wo), "w1": (x, w1), "w2": (h, w2)}.items():
        ms = timeit(lambda: torch.matmul(a, w.t()))
        print(f"cuBLAS {name:4s}: {ms*1e3:8.1f} us  {2.0*a.shape[0]*w.shape[0]*a.shape[1]/ms/1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
