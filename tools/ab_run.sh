run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --steps 8 --warmup 3 "$@" 2>gpurun_out/ab_$name.err | tee gpurun_out/ab_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']), d['config']['stage_ms'], round(d['e2e']['value']), round(d['e2e']['blocking_render']['value']))"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run c3_v6 --scene soup --depth 2 --variant 6
run c3_v7 --scene soup --depth 2 --variant 7
run c2_v6 --variant 6 --steps 20
run c2_v7 --variant 7 --steps 20
run c4_v6 --scene instanced --variant 6
run c4_v7 --scene instanced --variant 7
run c1_v6 --scene cornell32 --width 256 --height 256 --depth 4 --variant 6 --steps 20
run c1_v7 --scene cornell32 --width 256 --height 256 --depth 4 --variant 7 --steps 20
python - <<'PY'
import sys
sys.path.insert(0,'.')
from gdpathtracing_b200 import PathTracingCamera, scenes
for name, mk, d in (("soup", lambda: scenes.triangle_soup(1000000), 2), ("instanced", lambda: scenes.instanced_grid(), 8), ("demo", lambda: scenes.demo_scene(), 8)):
    sc = mk(); grp = scenes.populate(sc)
    cam = PathTracingCamera(); cam.fov = sc.fov; cam.geometry_group = grp; cam.denoising_mode = 2
    cam.set_window_size(1920, 1080); cam.set_global_transform(sc.camera_transform12); cam.set_max_depth(d); cam.set_variant(7)
    cam.init()
    for _ in range(3): cam.render_device_only(); st = cam.stats()
    print(name, {k: st[k] for k in ("rays", "retraced", "k1_ms")})
    del cam
PY
