# Schedule 6: cap on alive paths per warp (balance vs batch size) on C2; ncu of the pooled kernel on C4-at-1080p.
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
for c in 0 36 40 48 56; do run alive$c GDPT_POOL_ALIVE=$c; done
run alive40_low8 GDPT_POOL_ALIVE=40 GDPT_SHADE_AT=8
run alive48_low12 GDPT_POOL_ALIVE=48 GDPT_SHADE_AT=12
GDPT_POOL_ALIVE=40 python tools/warp_profile.py --frames 6 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 2 -c 1 -o gpurun_out/prof_k_path_pool_c4 python tools/profile_frame.py --scene instanced --frames 4 > gpurun_out/ncu_k_path_pool_c4.log 2>&1
tail -2 gpurun_out/ncu_k_path_pool_c4.log
