#!/bin/bash
# Run on the GPU box: tools/k2_probe.py for every A/B variant library (tools/build_variant.sh); "base" = the library as built.
cd "$(dirname "$0")/.."
cp gdpathtracing_b200/libgdpt_cuda.so /tmp/libgdpt_cuda_base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so; else cp gdpathtracing_b200/ab/libgdpt_cuda_$v.so gdpathtracing_b200/libgdpt_cuda.so; fi
  echo "$v $(timeout 120 python tools/k2_probe.py 2>&1 | tail -1)"
done
cp /tmp/libgdpt_cuda_base.so gdpathtracing_b200/libgdpt_cuda.so
