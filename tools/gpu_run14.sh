# Schedule 6: knob combinations, warp profile, ncu full capture of k_path_pool on the demo frame.
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
run s6_8_16 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8 GDPT_SHADE_AT=16
run s6_8_24 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8 GDPT_SHADE_AT=24
run s6_12_24 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=12 GDPT_SHADE_AT=24
run s6_8_32 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8 GDPT_SHADE_AT=32
run s6_16_32 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=16 GDPT_SHADE_AT=32
BARGS="--scene instanced --width 1920 --height 1080 --steps 6"
run c4_s6_8_16 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8 GDPT_SHADE_AT=16
run c4_s6_8_32 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8 GDPT_SHADE_AT=32
run c4_s6_2_8 GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=2 GDPT_SHADE_AT=8
export GDPT_SCHEDULE=6 GDPT_REFILL_BELOW=8 GDPT_SHADE_AT=16
python tools/warp_profile.py --frames 4 > gpurun_out/warp_profile_s6.json 2>&1; cat gpurun_out/warp_profile_s6.json
python tools/warp_profile.py --frames 3 --scene instanced > gpurun_out/warp_profile_s6_c4.json 2>&1; cat gpurun_out/warp_profile_s6_c4.json
GDPT_SCHEDULE=5 python tools/warp_profile.py --frames 3 --scene instanced > gpurun_out/warp_profile_s5_c4.json 2>&1; cat gpurun_out/warp_profile_s5_c4.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 2 -c 1 -o gpurun_out/prof_k_path_pool python tools/profile_frame.py --frames 4 > gpurun_out/ncu_k_path_pool.log 2>&1
tail -3 gpurun_out/ncu_k_path_pool.log
