#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ab_feed5.txt
: > $out
export SWB_LIB=$PWD/swift_b200/libswb_s5.so SWB_PHASE_SHAPES=${SHAPES:-w1}
echo "== TMA feed + epilogue, no MMA, 18 clusters only" >> $out
SWB_GEMM_NOMMA=1 SWB_GEMM_MAX_CLUSTERS=18 timeout 200 python tools/gemm_phases.py 8 1.0 3 >> $out 2>&1
echo "== TMA feed + epilogue, no MMA, 37 clusters only" >> $out
SWB_GEMM_NOMMA=1 SWB_GEMM_MAX_CLUSTERS=37 timeout 200 python tools/gemm_phases.py 8 1.0 3 >> $out 2>&1
cat $out
