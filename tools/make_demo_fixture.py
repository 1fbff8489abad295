#!/usr/bin/env python3
"""Generate gdpathtracing_b200/assets/demo_scene.npz from the reference's demo ASSETS (data, not source).

Run in the authoring container only (needs /root/reference).  The GPU box has no reference
checkout, so the benchmark scene of BASELINE config C2 ("project demo scene: Gobot character +
ground, albedo textures") travels as this fixture:

  * cornell.obj and suzanne.obj            project/demo/geometry/*.obj
  * the Gobot ArrayMesh (3 surfaces)       project/demo/demo.tscn105485445.tmp:59-95
  * grass / icon albedo textures           project/demo/materials/grass_wales_diffuse.png,
                                           project/addons/jar_path_tracing/icons/icon.png
    stored box-filtered to 256x256 / 128x128 to keep the fixture small; GeometryGroup3D
    resizes every layer to texture_array_resolution anyway (geometry_group3d.cpp:295-299).

OBJ import (triangulation, winding, vertex de-duplication) is done by the Godot engine, which
is not in the reference tree, so it is not pinned: we fan-triangulate and order every triangle
so that cross(e1, e2) opposes the vertex normal, i.e. Godot's clockwise front faces, which is
what makes `front` (main.glsl:254-255) true for rays arriving against the normal.
Parity is defined from the flat triangle list onward.
"""
import base64
import os
import re
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/project"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gdpathtracing_b200", "assets", "demo_scene.npz")


def load_obj(path):
    v, vt, vn, surfaces, cur = [], [], [], [], None
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            v.append([float(x) for x in p[1:4]])
        elif p[0] == "vt":
            vt.append([float(x) for x in p[1:3]])
        elif p[0] == "vn":
            vn.append([float(x) for x in p[1:4]])
        elif p[0] == "usemtl":
            cur = {"key": {}, "pos": [], "nrm": [], "uv": [], "idx": []}
            surfaces.append(cur)
        elif p[0] == "f":
            if cur is None:
                cur = {"key": {}, "pos": [], "nrm": [], "uv": [], "idx": []}
                surfaces.append(cur)
            corners = []
            for tok in p[1:]:
                a = (tok.split("/") + ["", ""])[:3]
                key = (int(a[0]) - 1, int(a[1]) - 1 if a[1] else -1, int(a[2]) - 1 if a[2] else -1)
                if key not in cur["key"]:
                    cur["key"][key] = len(cur["pos"])
                    cur["pos"].append(v[key[0]])
                    # OBJ v-flip like Godot's importer: uv.y = 1 - vt.y
                    cur["uv"].append([vt[key[1]][0], 1.0 - vt[key[1]][1]] if key[1] >= 0 else [0.0, 0.0])
                    cur["nrm"].append(vn[key[2]] if key[2] >= 0 else [0.0, 0.0, 0.0])
                corners.append(cur["key"][key])
            for k in range(1, len(corners) - 1):
                tri = [corners[0], corners[k], corners[k + 1]]
                P = np.array([cur["pos"][i] for i in tri], np.float64)
                n = np.array(cur["nrm"][tri[0]], np.float64)
                if np.dot(np.cross(P[1] - P[0], P[2] - P[0]), n) > 0:
                    tri = [tri[0], tri[2], tri[1]]
                cur["idx"] += tri
    out = []
    for s in surfaces:
        out.append({"positions": np.array(s["pos"], np.float32), "normals": np.array(s["nrm"], np.float32),
                    "uvs": np.array(s["uv"], np.float32), "indices": np.array(s["idx"], np.int32)})
    return out


def oct_decode(e):
    e = e.astype(np.float64) / 65535.0 * 2.0 - 1.0
    v = np.stack([e[:, 0], e[:, 1], 1.0 - np.abs(e[:, 0]) - np.abs(e[:, 1])], axis=1)
    t = np.maximum(-v[:, 2], 0.0)
    v[:, 0] += np.where(v[:, 0] >= 0, -t, t)
    v[:, 1] += np.where(v[:, 1] >= 0, -t, t)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def load_gobot(path):
    text = open(path).read()
    start = text.index('[sub_resource type="ArrayMesh" id="ArrayMesh_x31ta"]')
    end = text.index("blend_shape_mode", start)
    block = text[start:end]
    surfaces = []
    for chunk in block.split('"aabb":')[1:]:
        def field(name):
            m = re.search(r'"%s": PackedByteArray\("([^"]*)"\)' % name, chunk)
            return base64.b64decode(m.group(1))
        vcount = int(re.search(r'"vertex_count": (\d+)', chunk).group(1))
        icount = int(re.search(r'"index_count": (\d+)', chunk).group(1))
        fmt = int(re.search(r'"format": (\d+)', chunk).group(1))
        assert fmt == 34359745559, fmt
        vd, ad, idd = field("vertex_data"), field("attribute_data"), field("index_data")
        pos = np.frombuffer(vd[:vcount * 12], np.float32).reshape(vcount, 3).copy()
        nt = np.frombuffer(vd[vcount * 12:vcount * 20], np.uint16).reshape(vcount, 4)
        nrm = oct_decode(nt[:, 0:2])
        stride = len(ad) // vcount
        uv = np.frombuffer(ad, np.uint8).reshape(vcount, stride)[:, :8].copy().view(np.float32).reshape(vcount, 2)
        idx = np.frombuffer(idd[:icount * 2], np.uint16).astype(np.int32)
        surfaces.append({"positions": pos, "normals": nrm, "uvs": uv.copy(), "indices": idx})
    return surfaces


def shrink(path, size):
    im = Image.open(path).convert("RGBA").resize((size, size), Image.BOX)
    return np.asarray(im, np.uint8).copy()


def main():
    cornell = load_obj(f"{REF}/demo/geometry/cornell.obj")
    suzanne = load_obj(f"{REF}/demo/geometry/suzanne.obj")
    gobot = load_gobot(f"{REF}/demo/demo.tscn105485445.tmp")
    data = {}
    for name, surfs in (("cornell", cornell), ("suzanne", suzanne), ("gobot", gobot)):
        data[f"{name}_n"] = np.array(len(surfs), np.int32)
        for i, s in enumerate(surfs):
            for k, v in s.items():
                data[f"{name}_{i}_{k}"] = v
        print(name, "surfaces", len(surfs), "triangles", [len(s["indices"]) // 3 for s in surfs])
    data["tex_grass"] = shrink(f"{REF}/demo/materials/grass_wales_diffuse.png", 256)
    data["tex_icon"] = shrink(f"{REF}/addons/jar_path_tracing/icons/icon.png", 128)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **data)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
