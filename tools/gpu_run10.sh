set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print("$name:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1))
PY
}
run base GDPT_X=0
run shade16 GDPT_SHADE_AT=16
run shade24 GDPT_SHADE_AT=24
run shade4 GDPT_SHADE_AT=4
run refill16 GDPT_REFILL_BELOW=16
run refill28 GDPT_REFILL_BELOW=28
run refill32 GDPT_REFILL_BELOW=32
run burst1 GDPT_BURST=1
run burst2 GDPT_BURST=2
run burst8 GDPT_BURST=8
run lead64 GDPT_LEAD_MIN=64
run lead1 GDPT_LEAD_MIN=1
run minb5 GDPT_PATH_MINB=5
run bps3 GDPT_BLOCKS_PER_SM=3
run bps2 GDPT_BLOCKS_PER_SM=2
python tools/warp_profile.py --frames 4 > gpurun_out/warp_profile.json 2>&1; cat gpurun_out/warp_profile.json
