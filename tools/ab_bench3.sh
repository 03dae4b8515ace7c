#!/bin/bash
# same-box comparison of bench configurations: tools/ab_bench3.sh NAME "args A" "args B" "args C" ...
mkdir -p gpurun_out
name=$1; shift
for rep in 1 2; do
  i=0
  for cfg in "$@"; do
    i=$((i+1))
    python bench.py --no-cpu --no-strong --no-extras --steps 8 --warmup 3 $cfg > gpurun_out/ab_${name}_${i}_${rep}.json 2> gpurun_out/ab_${name}_${i}_${rep}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_${name}_${i}_${rep}.json"))
print("[$cfg] rep $rep", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 1), d["clocks"]["sm_mhz"])
PY
  done
done
