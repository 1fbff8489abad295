# schedule 5 (closest-hit + proof) bring-up: GPU tests, then A/B bench lines
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for cfg in "5 4" "5 5" "5 6" "5 8" "3 1"; do
  set -- $cfg
  GDPT_SCHEDULE=$1 GDPT_PATH_MINB=$2 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s$1_m$2.json 2> gpurun_out/bench_s$1_m$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_s$1_m$2.json"))
print("schedule $1 minb $2:", round(d["value"],1), "Mrays/s", d["ms_per_step"], "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1))
PY
done
python tools/warp_profile.py --frames 4 > gpurun_out/warp_profile.json 2>&1; cat gpurun_out/warp_profile.json
for sc in soup instanced; do python tools/profile_frame.py --scene $sc --frames 3 2>&1 | tail -2; GDPT_SCHEDULE=3 python tools/profile_frame.py --scene $sc --frames 3 2>&1 | tail -1; done
