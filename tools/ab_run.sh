python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --c5-spp 64 --c5-present 16 2>gpurun_out/bench_n2.err > gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config']['gather_ms'], d['config']['exchange'])
print(json.dumps(d['extra'], indent=0))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-c5 --partition rows 2>gpurun_out/bench_n2r.err > gpurun_out/bench_n2r.json; tail -3 gpurun_out/bench_n2r.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2r.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config']['gather_ms'], d['config']['partition'])
"
