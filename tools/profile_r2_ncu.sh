#!/bin/bash
# Round 2: ncu --set full captures of the reverse-mode kernels (one training step of Swift-B, tools/profile_train.py) and of
# the headline forward's dominant kernel; summaries are read back with tools/ncu_summary.py and committed under profiles/.
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
prof() { # name regex skip script
  timeout 300 $NCU --set full --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof_r2_$1 python $4 > gpurun_out/prof_r2_$1.log 2>&1
  echo "$1 rc=$?"
}
prof attn_bwd 'attn_bwd_tc_kernel' 1 tools/profile_train.py
prof wgrad_w1 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)0, \(bool\)0' 40 tools/profile_train.py
prof dgrad_w1 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)0, \(bool\)0' 41 tools/profile_train.py
prof w1 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)4' 14 tools/profile_forward.py
ls -la gpurun_out/prof_r2_* | head
