run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" 2>gpurun_out/ab_$name.err | tee gpurun_out/ab_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']), d['config']['stage_ms'], round(d['e2e']['value']), round(d['e2e']['blocking_render']['value']))"; }
run c2_v7 --variant 7
run c2_v7_noproof2 --variant 7 --tune NOPROOF2=1
run c2_v7_nibble --variant 7 --tune POOL_ALIVE=1
run c2_v7_both --variant 7 --tune POOL_ALIVE=1 --tune NOPROOF2=1
run c2_v6 --variant 6
run c2_v6_noproof2 --variant 6 --tune NOPROOF2=1
