#!/bin/bash
# Round 2, second session: ncu --set full captures of the forward's kernels after the skewed sub-tiles / single-value residual
# stream / TMA LayerNorm epilogue (chunk of 8 samples, tools/profile_forward.py: fuse_ln 3).
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
prof() { # name regex skip
  timeout 300 $NCU --set full --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof_r2b_$1 python tools/profile_forward.py 3 > gpurun_out/prof_r2b_$1.log 2>&1
  echo "$1 rc=$?"
}
for k in "$@"; do
  case $k in
    wo_ln) prof wo_ln 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)12' 14 ;;
    w2_ln) prof w2_ln 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)12' 15 ;;
    w1) prof w1 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)4' 14 ;;
    qkv) prof qkv 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)3' 14 ;;
    attn) prof attn 'window_attention_tc_kernel' 14 ;;
  esac
done
ls -la gpurun_out/prof_r2b_* | head -20
