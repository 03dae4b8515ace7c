"""Decompose the GEMM time at the four transformer shapes: main loop only (epi 6), + TMEM drain (epi 7), + stores (epi 1).
    python tools/gemm_phases.py [chunk] [seconds per variant]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import statistics
import subprocess
import threading
import time

import torch

from swift_b200 import _lib


NSM = 2 * int(os.environ["SWB_GEMM_MAX_CLUSTERS"]) if os.environ.get("SWB_GEMM_MAX_CLUSTERS") else 148   # profiling builds only


class Smi:
    """median SM clock / power while a block runs (nvidia-smi sampled every 100 ms)"""

    def __enter__(self):
        self.lines = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True).start()
        return self

    def __exit__(self, *a):
        self.p.terminate()
        v = [tuple(float(x) for x in l.split(",")) for l in self.lines[3:] if l.count(",") == 1]
        self.mhz = statistics.median(x[0] for x in v) if v else float("nan")
        self.watt = statistics.median(x[1] for x in v) if v else float("nan")


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0      # seconds of back-to-back launches per variant
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 3          # 1: 128x176, 2: 256x176 (double-buffered), 3: 256x352
    lib = _lib.lib()
    M = chunk * 8192
    st = torch.cuda.current_stream().cuda_stream
    shapes = {"qkv": (3168, 1056), "wo": (1056, 1056), "w1": (5632, 1056), "w2": (1056, 2816)}
    print(f"M = {M}; tile config {tile}; fp16 operands; each variant runs back to back for ~{secs:.0f} s with nvidia-smi sampling")
    if os.environ.get("SWB_PHASE_SHAPES"):
        shapes = {k: v for k, v in shapes.items() if k in os.environ["SWB_PHASE_SHAPES"].split(",")}
    for name, (N, K) in shapes.items():
        A = (torch.randn(M, K, device="cuda") * 0.5).half()
        W = (torch.randn(N, K, device="cuda") * 0.05).half()
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        res = {}
        variants = ((6, "mainloop"), (7, "+drain"), (8, "+smem"), (9, "direct"), (1, "store16"))
        if os.environ.get("SWB_PHASE_VARIANTS"):         # e.g. "6,1": main loop and store16 only
            keep = {int(x) for x in os.environ["SWB_PHASE_VARIANTS"].split(",")}
            variants = tuple(v for v in variants if v[0] in keep)
        if os.environ.get("SWB_LIB_ANYABI"):
            variants = ((1, "store16"),)                 # older builds do not have the profiling epilogues
        for epi, label in variants:
            for _ in range(3):
                _lib.check(lib.swb200_gemm(epi, tile, 1, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, st))
            torch.cuda.synchronize()
            reps = max(10, int(secs * 1e3 / max(1e-3, 2.0 * M * N * K / 1.3e12)))
            with Smi() as smi:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    _lib.check(lib.swb200_gemm(epi, tile, 1, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, st))
                e1.record()
                torch.cuda.synchronize()
            res[label] = (e0.elapsed_time(e1) / reps, smi.mhz, smi.watt)
        if name in ("qkv", "w1") and not os.environ.get("SWB_PHASE_VARIANTS"):   # the real fused epilogue of this shape
            if name == "qkv":
                qs = torch.full((12,), 10.0, device="cuda")
                o2 = torch.empty(3 * 12 * M * 96, device="cuda", dtype=torch.float16)
                fn = lambda: _lib.check(lib.swb200_gemm_qkv(tile, 1, 1, A.data_ptr(), K, W.data_ptr(), qs.data_ptr(), o2.data_ptr(), M, 1056, 12, st))
            else:
                o2 = torch.empty(M, N // 2, device="cuda", dtype=torch.float16)
                fn = lambda: _lib.check(lib.swb200_gemm_swiglu(tile, 1, A.data_ptr(), K, W.data_ptr(), o2.data_ptr(), M, 1056, N // 2, st))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            reps = max(10, int(secs * 1e3 / max(1e-3, 2.0 * M * N * K / 1.3e12)))
            with Smi() as smi:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
            res["fused_epi"] = (e0.elapsed_time(e1) / reps, smi.mhz, smi.watt)
            del o2
        fl = 2.0 * M * N * K
        print(f"  {name:4s} N={N:5d} K={K:5d}: " + "\n        " + "\n        ".join(
            f"{k:9s} {v[0] * 1e3:7.1f} us ({fl / v[0] / 1e9:6.0f} TF/s, {v[1]:4.0f} MHz, {v[2]:4.0f} W, "
            f"{fl / (v[0] * 1e-3) / (NSM * v[1] * 1e6):5.0f} FLOP/clk/SM)" for k, v in res.items()))
        del A, W, out


if __name__ == "__main__":
    main()
