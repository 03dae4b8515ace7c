#!/bin/bash
# Round-1 last build: ncu --set full captures of the three largest kernels (same recipe as tools/profile_r1p.sh).
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
prof() { # name regex skip
  timeout 240 $NCU --set full --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof_r1s_$1 python tools/profile_forward.py > gpurun_out/prof_r1s_$1.log 2>&1
}
prof w1 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)4' 14
prof w2ln 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)10' 14
prof attn 'window_attention_tc_kernel' 15
python -m pytest tests -q -m gpu -rs 2>&1 | grep -i "skip" | head -12
ls -la gpurun_out/prof_r1s_* | head
