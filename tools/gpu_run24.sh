# Cost hint: running mean vs last frame (C2, C4-at-1080p), 40 steps so the mean settles.
run() { name=$1; shift
  env "$@" python bench.py --steps 40 --warmup 10 --no-cpu-baseline $BARGS > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -5 gpurun_out/ab_$name.err
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
BARGS=""
run ema1 GDPT_COST_EMA=1
run ema0 GDPT_COST_EMA=0
run ema1b GDPT_COST_EMA=1
run ema0b GDPT_COST_EMA=0
BARGS="--scene instanced --width 1920 --height 1080 --steps 8 --warmup 6"
run c4_ema1 GDPT_COST_EMA=1
run c4_ema0 GDPT_COST_EMA=0
GDPT_COST_EMA=1 python tools/warp_profile.py --frames 12 2>&1 | tail -1
