# Four-wide closest-hit tables: GPU tests, then A/B against the two-wide tables on C2 / C4-at-1080p / C3 / C1.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
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
BARGS=""
run w1 GDPT_WIDE_BVH=1
run w0 GDPT_WIDE_BVH=0
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
run c4_w1 GDPT_WIDE_BVH=1
run c4_w0 GDPT_WIDE_BVH=0
BARGS="--scene soup --depth 2 --steps 6"
run c3_w1 GDPT_WIDE_BVH=1
run c3_w0 GDPT_WIDE_BVH=0
GDPT_WIDE_BVH=1 python tools/warp_profile.py --frames 6 2>&1 | tail -1
