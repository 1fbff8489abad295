#!/bin/bash
# Run on the GPU box: bench every A/B variant library (tools/build_variant.sh) on C2 and C4 at 1080p, short lines.
# usage: tools/ab_variants.sh <name> [<name> ...]   ("base" = the library as built)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp gdpathtracing_b200/libgdpt_cuda.so /tmp/libgdpt_cuda_base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so; else cp gdpathtracing_b200/ab/libgdpt_cuda_$v.so gdpathtracing_b200/libgdpt_cuda.so; fi
  touch gdpathtracing_b200/libgdpt_cuda.so gdpathtracing_b200/libgdpt_host.so
  for cfg in "c2:--scene demo" "c4:--scene instanced --width 1920 --height 1080 --variant 6" "c3:--scene soup --depth 2"; do
    tag=${cfg%%:*}; args=${cfg#*:}
    python bench.py $args --steps 20 --warmup 5 --no-c5 --no-cpu-baseline --no-schedule3 $AB_EXTRA > gpurun_out/ab_${v}_${tag}.json 2> gpurun_out/ab_${v}_${tag}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${v}_${tag}.json"))
    print("$v $tag", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"], 1), "blocking", round(d["e2e"]["blocking_render"]["value"], 1))
except Exception as e:
    print("$v $tag FAILED", e)
PY
  done
done
cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so
