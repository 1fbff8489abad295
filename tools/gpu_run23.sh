# After the node-step diet (one table for both levels, space restore in the crossing phase): tests + C2/C4/C3 lines.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline $BARGS > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -5 gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$name.json"))
    print("AB $name:", round(d["value"],1), "Mrays/s", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("AB $name FAILED", e)
PY
}
BARGS=""
run e6
run e5 GDPT_SCHEDULE=5
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
run c4_e6
BARGS="--scene soup --depth 2 --steps 6"
run c3_e6
