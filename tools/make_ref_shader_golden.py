#!/usr/bin/env python3
"""Writes tests/golden/ref_shader_hashes.json: SHA-256 digests of what the REFERENCE'S OWN SHADER TEXT
(oracle/_ref/libgdpt_refshader.so = main.glsl + brdfs.glsl compiled as C++, oracle/Makefile) produces on the
cases of tests/test_ref_shader.py -- RGBA8 frame, depth bits, radiance bits before quantisation, node-visit
order of the camera rays, per-segment hit ids / counters / visit hashes, frame totals.  Run in the authoring
container (needs /root/reference for the build); the GPU box, where the reference is absent, checks the
restatement and the CUDA kernels against these digests."""
import json
import os
import sys

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import __graft_entry__ as entry  # noqa: E402

entry.build()
import test_ref_shader as t  # noqa: E402
from oracle import oracle  # noqa: E402

assert oracle.ref_shader_available(), "build oracle/_ref/libgdpt_refshader.so first (make -C oracle)"
out = {"_source": "tools/make_ref_shader_golden.py: outputs of the reference's shader text compiled as C++ "
                  "(oracle/ref_shader_bridge.cpp); segments per pixel = %d, visits per camera ray = %d" % (t.SEGS, t.VISITS),
       "frames": {}}
for name, make, W, H, depth, frame in t.CASES:
    sc, osc = t.build(make)
    out["frames"][name] = t.frame_digest(t.render(osc, sc, W, H, depth, frame, "reference"))
    print(name, out["frames"][name]["rays"], "rays")
with open(t.GOLDEN, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
    f.write("\n")
