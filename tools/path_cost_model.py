#!/usr/bin/env python3
"""CPU analysis (no GPU): per-pixel step counts of the culled compact-order traversal on a scene, via the
test-only host compile of the device functions (tests/devcheck).  Used to size scheduling decisions."""
import argparse, ctypes, os, sys, json
from concurrent.futures import ThreadPoolExecutor
import numpy as np
REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, REPO)
import __graft_entry__ as entry
entry.build_checkers()
from gdpathtracing_b200 import nodes, scenes
from oracle import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="demo")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--out", default="")
ap.add_argument("--fast", type=int, default=0, help="1: closest-hit search (pt_fast.cuh) instead of the culled reference-order traversal")
a = ap.parse_args()
sc = {"demo": scenes.demo_scene, "cornell32": scenes.cornell32, "instanced": scenes.instanced_grid,
      "soup": lambda: scenes.triangle_soup(1_000_000)}[a.scene]()
grp = scenes.populate(sc)
grp.build()
osc = oracle.Scene(grp.buffers(), grp.texture_layers())
lib = ctypes.CDLL(os.path.join(REPO, "tests", "devcheck", "_build", "libgdpt_devcheck.so"))
W, H = a.width, a.height
params = np.zeros(9, np.uint32); params[4], params[5] = W, H
cam = np.frombuffer(bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, 1)), np.uint8).copy()
out = np.zeros((H, W, 6), np.uint32)
P = lambda x: x.ctypes.data_as(ctypes.c_void_p)
lib.devcheck_set_fast(a.fast)
lib.devcheck_path_costs.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
def rows(y0):
    lib.devcheck_path_costs(ctypes.byref(osc.c), P(params), P(cam), a.depth, y0, min(y0 + 8, H), P(out))
with ThreadPoolExecutor(os.cpu_count()) as ex:
    list(ex.map(rows, range(0, H, 8)))
c = out.reshape(-1, 6).astype(np.int64)
node, leaf, tri, inst, seg, big = c.T
steps4 = node + inst + leaf + (0 if a.fast else np.maximum(0, (tri - leaf + 3) // 4))   # leaf step = entry+first test, then 4 per step (approx)
live = seg > 0
print(json.dumps({
    "pixels": int(live.sum()), "rays": int(seg.sum()), "paths_with_hit": int((seg > 1).sum() + ((seg == 1) & (node + tri > 8)).sum()),
    "lane_steps": {"node": int(node.sum()), "leaf_entries": int(leaf.sum()), "tri_tests": int(tri.sum()), "instance": int(inst.sum())},
    "steps_per_path_pcts(50,90,99,99.9,100)": [float(np.percentile(steps4[seg > 1], p)) for p in (50, 90, 99, 99.9, 100)],
    "longest": {k: int(v) for k, v in zip(("node", "leaf", "tri", "inst", "seg", "bigleaf"), c[np.argmax(steps4)])},
}, indent=1))
order = np.argsort(-steps4)[:12]
for i in order:
    print("pixel", int(i % W), int(i // W), dict(zip(("node", "leaf", "tri", "inst", "seg", "bigleaf"), map(int, c[i]))), "steps~", int(steps4[i]))
if a.out:
    np.save(a.out, out)
