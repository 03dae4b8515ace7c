#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ab_roles2.txt
: > $out
echo "== A in smem, skew 0" >> $out
SWB_LIB=$PWD/swift_b200/libswb_s5.so SWB_GEMM_SKEW=0 timeout 200 python tools/gemm_roles.py 24 >> $out 2>&1
echo "== A staged in TMEM (tcgen05.cp), skew 0" >> $out
SWB_LIB=$PWD/swift_b200/libswb_at.so SWB_GEMM_SKEW=0 timeout 200 python tools/gemm_roles.py 24 >> $out 2>&1
echo "== A in smem, skew 2" >> $out
SWB_LIB=$PWD/swift_b200/libswb_s5.so SWB_GEMM_SKEW=2 timeout 200 python tools/gemm_roles.py 24 >> $out 2>&1
cat $out
