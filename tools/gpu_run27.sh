# quick three-scene line for compile-time A/B
run() { name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline $BARGS > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -5 gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$name.json"))
    print("AB $name:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("AB $name FAILED", e)
PY
}
mkdir -p gpurun_out
TAG=${1:-x}
BARGS=""; run ${TAG}_c2
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"; run ${TAG}_c4
BARGS="--scene soup --depth 2 --steps 6"; run ${TAG}_c3
