# Scaling table: sample-index partition on N GPUs ($1), C2 1080p and C5 4K instanced.
set -x
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
go() { name=$1; shift
  timeout 600 $TR bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; echo rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_$name.json"))
    print("N=$N $name:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms/step e2e", round(d["e2e"]["value"],1), d["config"]["partition"][:50])
except Exception as e:
    print("N=$N $name FAILED", e); print(open("gpurun_out/bench_n${N}_$name.err").read()[-1500:])
PY
}
go c2_sample --steps 20 --warmup 3
go c5_sample --steps 16 --warmup 3 --scene instanced --width 3840 --height 2160
