# Schedule 6 (pooled paths): parity tests, then A/B against schedule 5 on C2 and C4-at-1080p, pool knobs.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest13.log; cat gpurun_out/pytest13.log
run() { # name, env..., -- bench args
  name=$1; shift
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
run s5 GDPT_SCHEDULE=5
run s6 GDPT_SCHEDULE=6
run s6_swap2 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=2
run s6_swap8 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8
run s6_low16 GDPT_SCHEDULE=6 GDPT_SHADE_AT=16
run s6_low4 GDPT_SCHEDULE=6 GDPT_SHADE_AT=4
run s6_b3 GDPT_SCHEDULE=6 GDPT_BLOCKS_PER_SM=3
run s6_b2 GDPT_SCHEDULE=6 GDPT_BLOCKS_PER_SM=2
run s6_burst4 GDPT_SCHEDULE=6 GDPT_BURST=4
run s6_burst16 GDPT_SCHEDULE=6 GDPT_BURST=16
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
run c4_s5 GDPT_SCHEDULE=5
run c4_s6 GDPT_SCHEDULE=6
