#!/bin/bash
# Round-1 last session: new-feature tests, the bench line, the store demo, a fresh launch list, then the full GPU suite.
# Ordered by value: a clamped call still leaves the early artefacts in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_jvp.py -q -s -m gpu -k "rollout_and_save or reference_noise or scm_output_cotangent" > gpurun_out/r1s_newtests.log 2>&1
echo "newtests rc=$?" > gpurun_out/r1s_status.txt
python bench.py > gpurun_out/bench_r1s.json 2> gpurun_out/bench_r1s.err
echo "bench rc=$?" >> gpurun_out/r1s_status.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1s_reference.json 2>> gpurun_out/bench_r1s.err
echo "bench-ref rc=$?" >> gpurun_out/r1s_status.txt
( time python -m swift_b200.generate --output /tmp/fc.zarr --members 12 --ics 2 --steps 8 --dump zarr-step ) > gpurun_out/r1s_generate.log 2>&1
echo "generate rc=$?" >> gpurun_out/r1s_status.txt
du -sh /tmp/fc.zarr >> gpurun_out/r1s_generate.log 2>&1; ls /tmp/fc.zarr/geopotential | head -3 >> gpurun_out/r1s_generate.log
timeout 200 ncu --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01s_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r1s_ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r1s_status.txt
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r1s_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r1s_status.txt
tail -3 gpurun_out/r1s_newtests.log; cat gpurun_out/r1s_status.txt; tail -3 gpurun_out/r1s_gputests.log
