# Round artefacts on the final code: GPU tests, official bench line, reference arm, launch list, full ncu captures,
# warp profile, the other BASELINE configs (C1, C3, C4 at 4K).
set -x
bash tools/gpu_round1_artifacts.sh > gpurun_out/artifacts.log 2>&1
grep -v "^+" gpurun_out/artifacts.log | tail -12
python bench.py --scene cornell32 --width 256 --height 256 --depth 4 --steps 30 --warmup 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --scene soup --depth 2 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --scene instanced --width 3840 --height 2160 --depth 8 --steps 16 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for c in c1 c3 c4; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$c.json"))
    print("$c:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["value"], "roof", d["roofline"]["frac"])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/bench_$c.err").read()[-1500:])
PY
done
