# Bench lines of the other BASELINE configs (C1, C3, C4) on one B200; each is a full bench.py line.
set -x
mkdir -p gpurun_out
python bench.py --scene cornell32 --width 256 --height 256 --depth 4 --steps 30 --warmup 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --scene soup --depth 2 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --scene instanced --width 3840 --height 2160 --depth 8 --steps 16 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for c in c1 c3 c4; do tail -2 gpurun_out/bench_$c.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$c.json"))
print("$c:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"], "roof", d["roofline"]["frac"], d["config"].get("work_per_frame"))
PY
done
python tools/warp_profile.py --scene instanced --width 3840 --height 2160 --frames 3 > gpurun_out/warp_profile_c4.json 2>&1; cat gpurun_out/warp_profile_c4.json
