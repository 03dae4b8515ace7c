#!/bin/bash
# Same-box A/B of the headline bench for an env knob: tools/ab_bench.sh NAME "ENV_A" "ENV_B" [extra bench args]
mkdir -p gpurun_out
name=$1; a=$2; b=$3; shift 3
for rep in 1 2; do
  for tag in a b; do
    envs=$a; [ $tag = b ] && envs=$b
    env $envs python bench.py --no-cpu --no-strong --no-extras --steps 8 --warmup 3 "$@" > gpurun_out/ab_${name}_${tag}${rep}.json 2> gpurun_out/ab_${name}_${tag}${rep}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_${name}_${tag}${rep}.json"))
print("${tag}${rep} [$envs]", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 1), d["clocks"])
PY
  done
done
