#!/bin/bash
# Run on the GPU box: C4 at 1080p (instanced scene, schedule 6) for every A/B variant library; "base" = as built.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp gdpathtracing_b200/libgdpt_cuda.so /tmp/libgdpt_cuda_base.so
i=0
for v in "$@"; do
  i=$((i+1))
  if [ "$v" = base ]; then cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so; else cp gdpathtracing_b200/ab/libgdpt_cuda_$v.so gdpathtracing_b200/libgdpt_cuda.so; fi
  python bench.py --scene instanced --width 1920 --height 1080 --steps 10 --warmup 3 --no-c5 --no-cpu-baseline --no-schedule3 $AB_EXTRA > gpurun_out/abc4_${i}_${v}.json 2> gpurun_out/abc4_${i}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abc4_${i}_${v}.json"))
    print("$v c4", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("$v c4 FAILED", e)
PY
done
cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so
