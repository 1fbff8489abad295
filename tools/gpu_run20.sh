# Row bands on N GPUs: presented frame assembled by peer writes fused into K2 vs the NCCL all-gather; GPU tests first (1 GPU).
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_screens or row_sharding or progressive" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
go() { name=$1; shift
  timeout 600 $TR bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; echo rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_$name.json"))
    print("N=$N $name:", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms/step", d["config"]["stage_ms"], "e2e", round(d["e2e"]["value"],1), d["config"]["partition"][:60])
except Exception as e:
    print("N=$N $name FAILED", e); print(open("gpurun_out/bench_n${N}_$name.err").read()[-2500:])
PY
}
go c2_rows_peer --steps 20 --warmup 3 --partition rows
go c2_rows_gather --steps 20 --warmup 3 --partition rows --present gather
go c5_rows_peer --steps 8 --warmup 3 --scene instanced --width 3840 --height 2160 --partition rows
go c5_rows_gather --steps 8 --warmup 3 --scene instanced --width 3840 --height 2160 --partition rows --present gather
