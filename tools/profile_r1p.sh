set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none --kernel-name-base demangled"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r1p.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_r1p.log 2>&1
prof() { # name regex skip
  $NCU --set full --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o gpurun_out/prof_r1p_$1 python tools/profile_forward.py > gpurun_out/prof_r1p_$1.log 2>&1
}
prof w1 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)4' 14
prof qkv 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)3' 14
prof w2ln 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)10' 14
prof wo 'gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)1' 14
prof attn 'window_attention_tc_kernel' 15
prof ln 'ln_mod_residual_kernel' 14
ls -la gpurun_out/prof_r1p_* | head -20
