#!/bin/bash
# Round 2, second session: bench line, in-situ trace, launch list of the bench command (ncu captures: tools/profile_r2b_ncu.sh,
# at most two reports per gpurun call: 21 MB each against the 64 MB that travel back).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc=$?"
python tools/trace_step.py 24 3 > gpurun_out/r02b_trace.txt 2>&1; echo "trace rc=$?"
timeout 300 ncu --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-strong --no-extras --e2e-steps 1 > gpurun_out/r02b_ncu_bench.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2p.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
PY
