# Schedule 6: pool-slot variants and the wait-accumulator trigger on C2 / C4-at-1080p / C3.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "record_the_reference" 2>&1 | tail -3
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
for v in 0 2 3; do run var$v GDPT_POOL_VARIANT=$v; done
for w in 8 16 32 64; do run wait$w GDPT_POOL_WAIT=$w; done
run wait32_16_24 GDPT_POOL_WAIT=32 GDPT_REFILL_BELOW=16
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
for v in 0 2 3; do run c4_var$v GDPT_POOL_VARIANT=$v; done
for w in 8 16 32 64; do run c4_wait$w GDPT_POOL_WAIT=$w; done
run c4_wait16_low8 GDPT_POOL_WAIT=16 GDPT_SHADE_AT=8
BARGS="--scene soup --depth 2 --steps 6"
for w in 16 32; do run c3_wait$w GDPT_POOL_WAIT=$w; done
