"""Scenes used to pin the acceleration-structure builder, and the canonical byte views compared.

Canonical views (what both the product builder and the reference are reduced to):
  nodes      48 B BVHNode records, all bytes
  instances  176 B BLASInstance records, all bytes (unused material slots are zeroed on both sides)
  tlas       32 B TLASNode records with `blas` of INTERNAL nodes zeroed (indeterminate upstream)
  tri_geom   48 B per triangle (vertices)
  tri_attr   normals (3 x 12 B... as stored), uvs and surface index per triangle, packed 76 B
"""
import numpy as np

from gdpathtracing_b200 import scenes


def _degenerate():
    """All centroids identical on two axes and a >4-triangle coincident cluster: exercises the
    degenerate-axis early-out and the nth_element median fallback (bvh.cpp:54-55,170-177)."""
    sc = scenes.SceneDesc("degenerate", camera_transform12=scenes.transform12(None, (0, 0, 6)), fov=60.0)
    sc.materials = [dict()]
    sc.default_material = 0
    tris = []
    for k in range(24):  # a stack of identical triangles (same centroid) plus a few spread along x only
        x = 0.0 if k < 12 else float(k - 12) * 0.37
        tris.append([[x - 0.5, -0.5, 0.0], [x + 0.5, -0.5, 0.0], [x, 0.5, 0.0]])
    p = np.array(tris, np.float32).reshape(-1, 3)
    n = np.tile(np.array([0, 0, -1], np.float32), (len(p), 1))
    sc.meshes = [[{"positions": p, "normals": n, "uvs": np.zeros((len(p), 2), np.float32),
                   "indices": np.arange(len(p), dtype=np.int32)}]]
    sc.instances = [dict(mesh=0), dict(mesh=0, transform12=scenes.transform12(None, (0.0, 1.5, -1.0)))]
    return sc


CASES = {
    "cornell32": scenes.cornell32,
    "demo": scenes.demo_scene,
    "soup20k": lambda: scenes.triangle_soup(20000, seed=1),
    "instanced27x800": lambda: scenes.instanced_grid(3, 800, seed=3),
    "degenerate": _degenerate,
}


def material_ids(sc, blas_bytes):
    blas = np.frombuffer(blas_bytes, np.uint8).reshape(-1, 176)
    mids = blas[:, 164:176].copy().view(np.uint32)
    return [mids[k, :min(len(sc.meshes[i["mesh"]]), 3)] for k, i in enumerate(sc.instances)]


def canonical(nodes, instances, tlas, tri_geom, normals0, normals12, uvs, surf):
    t = np.frombuffer(np.ascontiguousarray(tlas).tobytes(), np.uint8).reshape(-1, 32).copy()
    internal = t[:, 12:16].copy().view(np.uint32)[:, 0] != 0
    t[internal, 28:32] = 0
    attr = np.concatenate([normals0, normals12, uvs, surf], axis=1)
    return {"nodes": np.frombuffer(np.ascontiguousarray(nodes).tobytes(), np.uint8),
            "instances": np.frombuffer(np.ascontiguousarray(instances).tobytes(), np.uint8),
            "tlas": t.reshape(-1), "tri_geom": np.ascontiguousarray(tri_geom).reshape(-1),
            "tri_attr": np.ascontiguousarray(attr).reshape(-1)}


def product_buffers(sc):
    grp = scenes.populate(sc)
    grp.build()
    b = grp.buffers()
    td = np.frombuffer(b["triangles_data"], np.uint8).reshape(-1, 80)
    tg = np.frombuffer(b["triangles_geometry"], np.uint8).reshape(-1, 48)
    return canonical(np.frombuffer(b["bvh"], np.uint8), np.frombuffer(b["blas"], np.uint8), np.frombuffer(b["tlas"], np.uint8),
                     tg, td[:, 0:12], td[:, 16:48], td[:, 48:72], td[:, 12:16]), b


def reference_buffers(sc):
    from oracle import oracle
    _, b = product_buffers(sc)  # only to learn which material ids GeometryGroup3D assigned to each instance
    ref = oracle.reference_arrays(sc, material_ids(sc, b["blas"]))
    tri = ref["triangles"].reshape(-1, 144)
    return canonical(ref["nodes"], ref["instances"], ref["tlas"], tri[:, 0:48], tri[:, 64:76], tri[:, 80:112], tri[:, 112:136],
                     tri[:, 136:140])
