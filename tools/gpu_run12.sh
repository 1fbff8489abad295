# Multi-GPU check on 2 B200s: both partitions of bench.py under torchrun, the reference arm under torchrun, and C5-style 4K rows.
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_sample.json 2> gpurun_out/bench_n2_sample.err; echo rc=$?; tail -3 gpurun_out/bench_n2_sample.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --partition rows > gpurun_out/bench_n2_rows.json 2> gpurun_out/bench_n2_rows.err; echo rc=$?; tail -3 gpurun_out/bench_n2_rows.err
timeout 600 $TR bench.py --gpus 2 --steps 16 --warmup 3 --scene instanced --width 3840 --height 2160 --partition rows > gpurun_out/bench_n2_c5_rows.json 2> gpurun_out/bench_n2_c5_rows.err; echo rc=$?; tail -3 gpurun_out/bench_n2_c5_rows.err
timeout 600 $TR bench.py --gpus 2 --steps 16 --warmup 3 --scene instanced --width 3840 --height 2160 > gpurun_out/bench_n2_c5_sample.json 2> gpurun_out/bench_n2_c5_sample.err; echo rc=$?; tail -3 gpurun_out/bench_n2_c5_sample.err
python bench.py --steps 16 --warmup 3 --scene instanced --width 3840 --height 2160 > gpurun_out/bench_n1_c5.json 2> gpurun_out/bench_n1_c5.err; echo rc=$?
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
for f in bench_n1 bench_n2_sample bench_n2_rows bench_n1_c5 bench_n2_c5_rows bench_n2_c5_sample; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json"))
    print("$f:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"],1), d["config"]["partition"], d["config"]["stage_ms"])
except Exception as e:
    print("$f: FAILED", e)
PY
done
