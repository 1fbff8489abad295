# Four-wide tables: burst length and service knobs.
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
BARGS=""
for b in 3 4 6 8 12; do run burst$b GDPT_BURST=$b; done
run swap8 GDPT_REFILL_BELOW=8
run swap16 GDPT_REFILL_BELOW=16
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
for b in 4 6 8 12; do run c4_burst$b GDPT_BURST=$b; done
