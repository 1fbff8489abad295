# Default path (schedule 6 + 8 cost classes): full GPU tests, C2/C4/C3 lines, warp profile; schedule 5 beside it.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest17.log
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
run d6
run d5 GDPT_SCHEDULE=5
run d6_w0 GDPT_POOL_WAIT=0
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
run c4_d6
run c4_d5 GDPT_SCHEDULE=5
BARGS="--scene soup --depth 2 --steps 6"
run c3_d6
run c3_d5 GDPT_SCHEDULE=5
python tools/warp_profile.py --frames 6 > gpurun_out/warp_profile_d6.json 2>&1; cat gpurun_out/warp_profile_d6.json
