"""Fused GEMM + LayerNorm-residual epilogue in isolation (wo and w2 shapes), next to the plain 16-bit store epilogue.
SWB_LN_DEBUG=1 skips the wait for the other groups' statistics, =2 the x update, =3 both (profiling only, wrong results).
    python tools/ln_phases.py [chunk] [seconds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

from gemm_phases import Smi
from swift_b200 import _lib


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lib = _lib.lib()
    B, T, D = chunk, 8192, 1056
    M = B * T
    st = torch.cuda.current_stream().cuda_stream
    xhl = (torch.randn(M, 2 * D, device="cuda") * 0.5).half()
    gain = torch.randn(B, D, device="cuda")
    bias = torch.randn(B, D, device="cuda")
    ws = torch.empty(lib.swb200_ln_workspace_bytes(M, D) + 256, dtype=torch.uint8, device="cuda")
    wsp = (ws.data_ptr() + 255) // 256 * 256
    out = torch.empty(M, D, device="cuda", dtype=torch.float16)
    print(f"M = {M}, SWB_LN_DEBUG = {os.environ.get('SWB_LN_DEBUG', '0')}")
    for name, K in (("wo", 1056), ("w2", 2816)):
        A = (torch.randn(M, K, device="cuda") * 0.5).half()
        W = (torch.randn(D, K, device="cuda") * 0.05).half()
        fl = 2.0 * M * D * K

        def plain():
            _lib.check(lib.swb200_gemm(1, 3, 1, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), D, M, D, K, st))

        def fused():
            _lib.check(lib.swb200_gemm_ln_residual(3, 1, A.data_ptr(), K, W.data_ptr(), K, xhl.data_ptr(), gain.data_ptr(),
                                                   bias.data_ptr(), M, D, T, wsp, 0, st))

        for label, fn in (("store16", plain), ("ln_fused", fused)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            reps = max(10, int(secs * 1e3 / (fl / 1.0e12)))
            with Smi() as smi:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"  {name} {label:9s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:6.0f} TF/s  {smi.mhz:5.0f} MHz {smi.watt:5.0f} W")
        del A, W


if __name__ == "__main__":
    main()
