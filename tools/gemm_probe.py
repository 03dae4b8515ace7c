"""Bring-up probe for the tcgen05 GEMM: runs tiny problems for cta_group 1 and 2, each in its own process
(a trapped kernel poisons the CUDA context), and prints where the result deviates from torch.matmul.

    python tools/gemm_probe.py            # all cases
    python tools/gemm_probe.py one CG M N K
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(cg, M, N, K):
    import torch
    from swift_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), float("nan"), device="cuda")
    rc = lib.swb200_gemm(0, cg, 0, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), ldo, M, N, K,
                         torch.cuda.current_stream().cuda_stream)
    print("rc", rc, lib.swb200_last_error())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    got = out[:, :N]
    nan = torch.isnan(got)
    print(f"cg={cg} M={M} N={N} K={K}: nan frac {nan.float().mean():.3f}")
    d = (got - ref).abs()
    d[nan] = 1e9
    ok = d < 2e-2 * ref.abs().max()
    print(f"  ok frac {ok.float().mean():.4f}  rel-L2 {((got.nan_to_num() - ref).norm() / ref.norm()):.3e}")
    if ok.float().mean() < 1.0:
        rb = ok.reshape(M // 32, 32, N).float().mean(dim=(1, 2))
        print("  ok by 32-row block:", [f"{v:.2f}" for v in rb.tolist()][:16])
        cb = ok[:, : N // 8 * 8].reshape(M, N // 8, 8).float().mean(dim=(0, 2))
        print("  ok by 8-col block:", [f"{v:.2f}" for v in cb.tolist()])
        print("  got[0,:8]", got[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        # does the result match a K-truncated product (a wrong k-advance)?
        for kk in (16, 32, 48, 64):
            if kk <= K:
                r2 = A[:, :kk].float() @ W[:, :kk].float().t()
                print(f"  rel-L2 vs first {kk} of K: {((got.nan_to_num() - r2).norm() / r2.norm()):.3e}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(*[int(v) for v in sys.argv[2:6]])
        return
    cases = [(1, 128, 176, 64), (2, 256, 176, 64), (3, 256, 352, 64), (3, 256, 352, 256), (3, 512, 352, 1056),
             (2, 512, 352, 1056), (3, 8192, 1056, 1056), (3, 300, 276, 568)]
    for c in cases:
        print("=" * 80)
        try:
            r = subprocess.run([sys.executable, __file__, "one", *map(str, c)], capture_output=True, text=True,
                               timeout=120)
            print(r.stdout[-3000:])
            if r.returncode != 0:
                print("EXIT", r.returncode, r.stderr[-2000:])
        except subprocess.TimeoutExpired:
            print("TIMEOUT", c)


if __name__ == "__main__":
    main()
