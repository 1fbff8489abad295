# Leaf size of the closest-hit trees (GDPT_FAST_LEAF_MAX) on C2 / C4-at-1080p / C3.
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
for l in 2 3 4 6 8; do
  BARGS=""; run leaf${l}_c2 GDPT_FAST_LEAF_MAX=$l
  BARGS="--scene instanced --width 1920 --height 1080 --steps 6"; run leaf${l}_c4 GDPT_FAST_LEAF_MAX=$l
  BARGS="--scene soup --depth 2 --steps 6"; run leaf${l}_c3 GDPT_FAST_LEAF_MAX=$l
done
