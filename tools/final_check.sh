#!/bin/bash
# What the driver runs at round end, on the final tree: smoke(), both bench arms, the GPU suite.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench.err; echo "bench-ref rc=$?"
python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err; echo "bench rc=$?"
python -m pytest tests -x -q -m gpu > gpurun_out/final_gputests.log 2>&1; echo "gputests rc=$?"
tail -1 gpurun_out/final_smoke.log; tail -1 gpurun_out/final_gputests.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "eager", d["gpu_eager_baseline"])
PY
