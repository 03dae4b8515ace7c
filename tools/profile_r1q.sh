cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none --kernel-name-base demangled"
$NCU --set full --import-source on -k 'regex:gemm_tcgen05_kernel<\(int\)2, \(int\)2, \(int\)2' --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r1q_embed python tools/profile_forward.py > gpurun_out/prof_r1q_embed.log 2>&1
ls -la gpurun_out/prof_r1q_embed.ncu-rep
