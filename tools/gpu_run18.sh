# Schedule 6 occupancy variants (5/6/8 blocks per SM) on C2 and C4-at-1080p.
set -x
mkdir -p gpurun_out
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
for v in 0 4 5 6; do run occ$v GDPT_POOL_VARIANT=$v; done
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
for v in 0 4 5 6; do run c4_occ$v GDPT_POOL_VARIANT=$v; done
