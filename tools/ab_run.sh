timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --c5-spp 32 --c5-present 16 2>gpurun_out/bench_try.err > gpurun_out/bench_try.json; tail -3 gpurun_out/bench_try.err; python -c "
import json; d=json.load(open('gpurun_out/bench_try.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d['roofline'], indent=0)[:1500])
print(json.dumps(d['roofline_accumulate']))
print(json.dumps(d['config']['reference_visiting_order']), d['config']['mrays_s_over_rays_that_enter_an_instance'])
print(json.dumps(d['extra'], indent=0))
print(json.dumps(d.get('cpu_baseline')))
"
