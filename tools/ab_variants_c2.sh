#!/bin/bash
# Run on the GPU box: C2 bench line (device value, e2e pipelined / blocking) for every A/B variant library; "base" = as built.
# usage: tools/ab_variants_c2.sh <name> [<name> ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp gdpathtracing_b200/libgdpt_cuda.so /tmp/libgdpt_cuda_base.so
i=0
for v in "$@"; do
  i=$((i+1))
  if [ "$v" = base ]; then cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so; else cp gdpathtracing_b200/ab/libgdpt_cuda_$v.so gdpathtracing_b200/libgdpt_cuda.so; fi
  python bench.py --scene demo --steps 30 --warmup 5 --no-c5 --no-cpu-baseline --no-schedule3 $AB_EXTRA > gpurun_out/abc2_${i}_${v}.json 2> gpurun_out/abc2_${i}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abc2_${i}_${v}.json"))
    print("$v c2", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"], 1), "blocking", round(d["e2e"]["blocking_render"]["value"], 1))
except Exception as e:
    print("$v c2 FAILED", e)
PY
done
cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so
