#!/bin/bash
# Run on the GPU box: short bench lines of C2 (and optionally C1 / C4 with schedule 6) for a list of tuning settings.
# usage: tools/ab_tune.sh "<label>:<--tune A=1 --tune B=2 ...>" ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%:*}; args=${spec#*:}
  for cfg in ${AB_CONFIGS:-c2 c4v6}; do case $cfg in c2) cfg="c2:--scene demo --steps 20 --warmup 5";; c1) cfg="c1:--scene cornell32 --width 256 --height 256 --depth 4 --steps 20 --warmup 5";; c4v6) cfg="c4v6:--scene instanced --width 1920 --height 1080 --steps 8 --warmup 3 --variant 6";; c4v7) cfg="c4v7:--scene instanced --width 1920 --height 1080 --steps 8 --warmup 3 --variant 7";; c4k6) cfg="c4k6:--scene instanced --width 3840 --height 2160 --steps 6 --warmup 3 --variant 6";; c4k7) cfg="c4k7:--scene instanced --width 3840 --height 2160 --steps 6 --warmup 3 --variant 7";; c3) cfg="c3:--scene soup --depth 2 --steps 8 --warmup 3";; esac
    tag=${cfg%%:*}; cargs=${cfg#*:}
    python bench.py $cargs --no-c5 --no-cpu-baseline --no-schedule3 $args > gpurun_out/abt_${label}_${tag}.json 2> gpurun_out/abt_${label}_${tag}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abt_${label}_${tag}.json"))
    print("$label $tag", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"], 1), "blocking", round(d["e2e"]["blocking_render"]["value"], 1))
except Exception as e:
    print("$label $tag FAILED", e); print(open("gpurun_out/abt_${label}_${tag}.err").read()[-400:])
PY
  done
done
