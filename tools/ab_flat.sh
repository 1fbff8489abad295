cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 1 0; do
python bench.py --scene demo --steps 20 --warmup 5 --no-c5 --no-cpu-baseline --no-schedule3 --tune FLAT=$f > gpurun_out/ab_flat${f}_c2.json 2> gpurun_out/ab_flat${f}_c2.err
python bench.py --scene cornell32 --width 256 --height 256 --depth 4 --steps 20 --warmup 5 --no-c5 --no-cpu-baseline --no-schedule3 --tune FLAT=$f > gpurun_out/ab_flat${f}_c1.json 2> gpurun_out/ab_flat${f}_c1.err
python bench.py --scene instanced --width 1920 --height 1080 --steps 10 --warmup 3 --no-c5 --no-cpu-baseline --no-schedule3 --variant 6 --tune FLAT=$f > gpurun_out/ab_flat${f}_c4v6.json 2> gpurun_out/ab_flat${f}_c4v6.err
for c in c2 c1 c4v6; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_flat${f}_$c.json"))
    print("FLAT=$f $c", round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 4), "ms", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"], 1), "own", d["roofline"].get("own_work_per_launch"))
except Exception as e:
    print("FLAT=$f $c FAILED", e); print(open("gpurun_out/ab_flat${f}_$c.err").read()[-600:])
PY
done
done
