# Schedule 6: parked-leaf depth A/B (GDPT_POOL_PARK 1..4) on C2 and C4-at-1080p; parity tests of the pooled kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "record_the_reference" 2>&1 | tail -5
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
export GDPT_SCHEDULE=6
BARGS=""
for p in 1 2 3 4; do run park$p GDPT_POOL_PARK=$p; done
run park3_b16 GDPT_POOL_PARK=3 GDPT_BURST=16
run park2_b16 GDPT_POOL_PARK=2 GDPT_BURST=16
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
for p in 1 2 3 4; do run c4_park$p GDPT_POOL_PARK=$p; done
BARGS="--scene soup --depth 2 --steps 6"
run c3_park1 GDPT_POOL_PARK=1
run c3_park2 GDPT_POOL_PARK=2
run c3_s5 GDPT_SCHEDULE=5
