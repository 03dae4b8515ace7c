#!/bin/bash
# A/B of the skewed sub-tile issue order (SWB_GEMM_SKEW) and of the pipeline depth (4 vs 5 stages) on one box.
# Needs the two profiling builds:  SWB_NVCC_DEFINES="-DSWB_GEMM_MAX_STAGES=4 -DSWB_PROFILE_EPILOGUES" SWB_LIB_OUT=.../libswb_s4.so
#                                  SWB_NVCC_DEFINES="-DSWB_PROFILE_EPILOGUES" SWB_LIB_OUT=.../libswb_s5.so  python -m swift_b200.build
mkdir -p gpurun_out
out=gpurun_out/ab_skew.txt
: > $out
echo "== correctness with skew 2 (5 stages, product lib)" >> $out
SWB_GEMM_SKEW=2 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -3 >> $out
for cfg in "s4 0" "s4 2" "s5 0" "s5 1" "s5 2" "s5 3"; do
  set -- $cfg
  echo "== lib $1 skew $2" >> $out
  SWB_LIB=$PWD/swift_b200/libswb_$1.so SWB_GEMM_SKEW=$2 timeout 300 python tools/gemm_phases.py 24 1.2 >> $out 2>&1
done
cat $out
