"""Deterministic scene descriptions for the BASELINE.json configurations.

A scene is plain data (numpy arrays + dicts): textures, StandardMaterial3D-like materials, meshes
(lists of surfaces), MeshInstance3D-like instances, a camera pose.  ``populate`` registers it with
the host layer's GeometryGroup3D twin, which then builds BLAS/TLAS and the GPU buffers exactly as
the reference's GeometryGroup3D::build does.  No rendering code lives here.

Random geometry uses a counter-based splitmix64 -> float32 stream (NOT a library distribution), so
every platform generates the same bytes.
"""
import os
from dataclasses import dataclass, field

import numpy as np

from .nodes import IDENTITY12, GeometryGroup3D

_REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
DEMO_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "demo_scene.npz")  # workload input (the demo scene's meshes and textures), not a golden output


@dataclass
class SceneDesc:
    name: str
    textures: list = field(default_factory=list)
    materials: list = field(default_factory=list)
    meshes: list = field(default_factory=list)
    instances: list = field(default_factory=list)
    default_material: int = -1
    texture_array_resolution: int = 512
    material_ext: bool = False  # SURVEY 8f-4 material breadth (nodes.GeometryGroup3D.material_ext)
    camera_transform12: np.ndarray = field(default_factory=lambda: IDENTITY12.copy())
    fov: float = 90.0

    def triangle_count(self):
        return sum(len(s["indices"]) // 3 for m in self.meshes for s in m)


def populate(scene):
    g = GeometryGroup3D()
    g.texture_array_resolution = scene.texture_array_resolution
    g.material_ext = scene.material_ext
    tex = [g.add_texture(t) for t in scene.textures]
    mats = []
    for m in scene.materials:
        m = dict(m)
        for key in ("albedo_texture", "roughness_texture", "metallic_texture"):
            if m.get(key, -1) >= 0:
                m[key] = tex[m[key]]
        mats.append(g.add_material(**m))
    if scene.default_material >= 0:
        g.set_default_material(mats[scene.default_material])
    meshes = [g.add_mesh(m) for m in scene.meshes]
    for inst in scene.instances:
        ov = [mats[i] if i >= 0 else -1 for i in inst.get("surface_overrides", ())]
        mo = inst.get("material_override", -1)
        g.add_mesh_instance(meshes[inst["mesh"]], inst.get("transform12", IDENTITY12),
                            mats[mo] if mo >= 0 else -1, ov)
    return g


# ------------------------------------------------------------------------------------ helpers
def splitmix64_floats(seed, n):
    """n float32 values in [0,1): the top 24 bits of successive splitmix64 outputs."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def transform12(basis_rows=None, origin=(0.0, 0.0, 0.0)):
    b = np.eye(3, dtype=np.float32) if basis_rows is None else np.asarray(basis_rows, np.float32).reshape(3, 3)
    return np.concatenate([b.reshape(9), np.asarray(origin, np.float32)]).astype(np.float32)


def godot_transform(xx, xy, xz, yx, yy, yz, zx, zy, zz, ox, oy, oz):
    """Transform3D(...) as written in a .tscn: Basis rows (row-major), then the origin.  With this
    reading the Cornell room's open +X side maps to +Z, i.e. towards the demo camera."""
    rows = np.array([[xx, xy, xz], [yx, yy, yz], [zx, zy, zz]], np.float32)
    return transform12(rows, (ox, oy, oz))


def _quad_surface(corners, normal, flip=False):
    """Two clockwise-front triangles (Godot convention: cross(e1,e2) opposes the normal)."""
    p = np.asarray(corners, np.float32)
    n = np.asarray(normal, np.float32)
    idx = [0, 1, 2, 0, 2, 3]
    if np.dot(np.cross(p[1] - p[0], p[2] - p[0]), n) > 0:
        idx = [0, 2, 1, 0, 3, 2]
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    return p, np.tile(n, (4, 1)), uv, np.array(idx, np.int32)


def _merge(quads):
    pos, nrm, uv, idx, base = [], [], [], [], 0
    for p, n, t, i in quads:
        pos.append(p); nrm.append(n); uv.append(t); idx.append(i + base)
        base += len(p)
    return {"positions": np.concatenate(pos), "normals": np.concatenate(nrm), "uvs": np.concatenate(uv),
            "indices": np.concatenate(idx).astype(np.int32)}


def _box_five_faces(half=(1.0, 1.0, 1.0)):
    """Axis-aligned box without its bottom face, outward normals: 5 quads = 10 triangles."""
    hx, hy, hz = half
    q = [
        _quad_surface([[-hx, hy, -hz], [hx, hy, -hz], [hx, hy, hz], [-hx, hy, hz]], [0, 1, 0]),
        _quad_surface([[hx, -hy, -hz], [hx, hy, -hz], [hx, hy, hz], [hx, -hy, hz]], [1, 0, 0]),
        _quad_surface([[-hx, -hy, -hz], [-hx, hy, -hz], [-hx, hy, hz], [-hx, -hy, hz]], [-1, 0, 0]),
        _quad_surface([[-hx, -hy, hz], [hx, -hy, hz], [hx, hy, hz], [-hx, hy, hz]], [0, 0, 1]),
        _quad_surface([[-hx, -hy, -hz], [hx, -hy, -hz], [hx, hy, -hz], [-hx, hy, -hz]], [0, 0, -1]),
    ]
    return [_merge(q)]


def _cornell_room():
    """The geometry of project/demo/geometry/cornell.obj: a +-5 cube open towards +X, inward normals,
    three surfaces (usemtl 1: top, x=-5 wall, floor; usemtl 2: z=-5 wall; usemtl 3: z=+5 wall)."""
    s0 = _merge([
        _quad_surface([[5, 5, -5], [5, 5, 5], [-5, 5, 5], [-5, 5, -5]], [0, -1, 0]),
        _quad_surface([[-5, -5, 5], [-5, -5, -5], [-5, 5, -5], [-5, 5, 5]], [1, 0, 0]),
        _quad_surface([[-5, -5, -5], [-5, -5, 5], [5, -5, 5], [5, -5, -5]], [0, 1, 0]),
    ])
    s1 = _merge([_quad_surface([[-5, -5, -5], [5, -5, -5], [5, 5, -5], [-5, 5, -5]], [0, 0, 1])])
    s2 = _merge([_quad_surface([[5, -5, 5], [-5, -5, 5], [-5, 5, 5], [5, 5, 5]], [0, 0, -1])])
    return [s0, s1, s2]


ROOM_TRANSFORM = godot_transform(-2.62268e-08, 0, -0.6, 0, 0.6, 0, 0.6, 0, -2.62268e-08, 0, 0, 0)  # demo.tscn:79
LIGHT_TRANSFORM = godot_transform(1, 0, 0, 0, -1, 1.50996e-07, 0, -1.50996e-07, -1, 0, 2.95581, 0)  # demo.tscn:74
CAMERA_TRANSFORM = godot_transform(1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 9.7694)                          # demo.tscn:52


# ------------------------------------------------------------------------------------ C1
def cornell32():
    """BASELINE config C1: 32 triangles, 4 BLAS / 4 instances, albedo-only Lambertian materials."""
    sc = SceneDesc("cornell32", camera_transform12=CAMERA_TRANSFORM.copy(), fov=79.5)
    sc.materials = [
        dict(),                                                       # 0 default StandardMaterial3D (demo.tscn:19)
        dict(emission=(0.832472, 0.8072, 0.719802), emission_energy_multiplier=10.0),  # 1 light (demo.tscn:23-26)
        dict(albedo=(1.0, 1.0, 1.0)),                                 # 2 white walls
        dict(albedo=(1.0, 0.16, 0.16)),                               # 3 red   (demo.tscn:31-32)
        dict(albedo=(0.42, 1.0, 0.13)),                               # 4 green (demo.tscn:34-35)
    ]
    sc.default_material = 0
    light = [_merge([_quad_surface([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1]], [0, 1, 0])])]  # PlaneMesh 2x2
    sc.meshes = [light, _cornell_room(), _box_five_faces((0.8, 1.6, 0.8)), _box_five_faces((0.8, 0.8, 0.8))]
    c, s = np.cos(0.3), np.sin(0.3)
    sc.instances = [
        dict(mesh=0, transform12=LIGHT_TRANSFORM, surface_overrides=[1]),
        dict(mesh=1, transform12=ROOM_TRANSFORM, surface_overrides=[2, 3, 4]),
        dict(mesh=2, transform12=transform12([[c, 0, s], [0, 1, 0], [-s, 0, c]], (-1.0, -1.4, -0.8))),
        dict(mesh=3, transform12=transform12([[c, 0, -s], [0, 1, 0], [s, 0, c]], (1.1, -2.2, 0.6))),
    ]
    return sc


# ------------------------------------------------------------------------------------ C2
def _checker(size=256, cells=16, seed=7):
    """Stand-in for Gobot's diffuse texture, which is not in the reference tree."""
    r = splitmix64_floats(seed, cells * cells * 3).reshape(cells, cells, 3)
    y, x = np.mgrid[0:size, 0:size]
    cy, cx = y * cells // size, x * cells // size
    base = np.where(((cx + cy) & 1)[..., None] == 0, 0.85, 0.35) * (0.6 + 0.4 * r[cy, cx])
    img = np.empty((size, size, 4), np.uint8)
    img[..., :3] = np.clip(base * 255.0 + 0.5, 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


def demo_scene(fixture=DEMO_FIXTURE):
    """BASELINE config C2: the banner scene of demo.tscn105485445.tmp:196-244 (Cornell room with grass
    albedo on surface 0, two Suzannes, Gobot), lit by the hard-coded sky; fov 90 (property default)."""
    d = np.load(fixture)

    def mesh(name):
        return [{k: d[f"{name}_{i}_{k}"] for k in ("positions", "normals", "uvs", "indices")}
                for i in range(int(d[f"{name}_n"]))]

    sc = SceneDesc("demo", camera_transform12=CAMERA_TRANSFORM.copy(), fov=90.0, texture_array_resolution=1024)
    sc.textures = [d["tex_grass"], d["tex_icon"], _checker()]
    sc.materials = [
        dict(),                                    # 0 default (StandardMaterial3D_avnmi)
        dict(albedo_texture=0),                    # 1 grass  (StandardMaterial3D_ahy1k)
        dict(albedo=(1.0, 0.16, 0.16)),            # 2 red
        dict(albedo=(0.42, 1.0, 0.13)),            # 3 green
        dict(albedo_texture=1),                    # 4 icon   (StandardMaterial3D_qeiev)
        dict(albedo_texture=2),                    # 5 checker: stands in for StandardMaterial3D_k1rqu / eye_mat
    ]
    sc.default_material = 0
    sc.meshes = [mesh("cornell"), mesh("suzanne"), mesh("gobot")]
    sc.instances = [
        dict(mesh=0, transform12=ROOM_TRANSFORM, surface_overrides=[1, 2, 3]),
        dict(mesh=1, transform12=godot_transform(0.918709, 0, 0.394936, -0.0683152, 0.984926, 0.158916, -0.388983,
                                                 -0.172978, 0.90486, -1.30439, -1.75104, -0.608295)),
        dict(mesh=1, transform12=godot_transform(1.44381, 0.158618, -0.332176, 0, 1.34457, 0.642048, 0.368104,
                                                 -0.622146, 1.30289, 0.790388, -1.15649, -1.17417),
             surface_overrides=[4]),
        # Gobot: overrides exist only for surfaces 1 and 2, surface 0 resolves to the default material
        dict(mesh=2, transform12=godot_transform(1, 0, 0, 0, 1, 0, 0, 0, 1, 1.63994, -2.53071, 1.88098),
             surface_overrides=[-1, 5, 5]),
    ]
    return sc


# ------------------------------------------------------------------------------------ C3
def _soup_surface(n_tris, seed, extent, half):
    r = splitmix64_floats(seed, n_tris * 12).reshape(n_tris, 12)
    centre = (r[:, 0:3] * np.float32(2.0) - np.float32(1.0)) * np.float32(extent)
    verts = centre[:, None, :] + ((r[:, 3:12].reshape(n_tris, 3, 3) * np.float32(2.0) - np.float32(1.0)) * np.float32(half))
    verts = verts.astype(np.float32)
    g = np.cross(verts[:, 1] - verts[:, 0], verts[:, 2] - verts[:, 0])
    ln = np.linalg.norm(g, axis=1, keepdims=True)
    n = np.where(ln > 0, -g / np.maximum(ln, 1e-30), np.array([0, 1, 0], np.float32)).astype(np.float32)
    return {"positions": verts.reshape(-1, 3), "normals": np.repeat(n, 3, axis=0), "uvs": np.zeros((n_tris * 3, 2), np.float32),
            "indices": np.arange(n_tris * 3, dtype=np.int32)}


def triangle_soup(n_tris=1_000_000, seed=1):
    """BASELINE config C3: centres uniform in [-10,10]^3, vertices = centre + uniform[-0.05,0.05]^3,
    one BLAS, identity instance, camera at (0,0,30) looking down -Z; render with MAX_DEPTH 2
    (primary + one cosine-weighted diffuse bounce)."""
    sc = SceneDesc(f"soup{n_tris}", camera_transform12=transform12(None, (0.0, 0.0, 30.0)), fov=45.0)
    sc.materials = [dict(albedo=(0.8, 0.8, 0.8), roughness=1.0, metallic=0.0)]
    sc.default_material = 0
    sc.meshes = [[_soup_surface(n_tris, seed, 10.0, 0.05)]]
    sc.instances = [dict(mesh=0, transform12=IDENTITY12.copy())]
    return sc


# ------------------------------------------------------------------------------------ C4 / C5
def instanced_grid(n_side=10, tris_per_blas=10_000, seed=3):
    """BASELINE config C4/C5: one seeded blob BLAS instanced on an n^3 grid with seeded rotations
    (1,000 instances x 10,000 triangles = 10 M instanced triangles; TLAS = 2,000 nodes < 65,535)."""
    n_inst = n_side ** 3
    sc = SceneDesc(f"instanced{n_inst}x{tris_per_blas}", fov=60.0)
    sc.materials = [dict(albedo=(0.75, 0.75, 0.75), roughness=0.8),
                    dict(albedo=(0.9, 0.4, 0.2), roughness=0.4, metallic=0.6)]
    sc.default_material = 0
    blob = _soup_surface(tris_per_blas, seed, 0.9, 0.06)
    # pull the blob into a ball so instances do not interpenetrate much
    p = blob["positions"].reshape(-1, 3, 3)
    c = p.mean(axis=1, keepdims=True)
    scale = np.minimum(1.0, 0.9 / np.maximum(np.linalg.norm(c, axis=2, keepdims=True), 1e-6)).astype(np.float32)
    blob["positions"] = (c * scale + (p - c)).reshape(-1, 3).astype(np.float32)
    sc.meshes = [[blob]]
    r = splitmix64_floats(seed + 1, n_inst * 3).reshape(n_inst, 3)
    spacing = 2.4
    for i in range(n_inst):
        ix, iy, iz = i % n_side, (i // n_side) % n_side, i // (n_side * n_side)
        ang = r[i] * np.float32(2.0 * np.pi)
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], np.float32)
        ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], np.float32)
        rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], np.float32)
        origin = ((np.array([ix, iy, iz], np.float32) - (n_side - 1) / 2.0) * spacing).astype(np.float32)
        sc.instances.append(dict(mesh=0, transform12=transform12((rz @ ry @ rx).astype(np.float32), origin),
                                 surface_overrides=[1 if (ix + iy + iz) % 3 == 0 else 0]))
    sc.camera_transform12 = transform12(None, (0.0, 0.0, n_side * spacing * 1.1))
    return sc


SCENES = {"cornell32": cornell32, "demo": demo_scene, "soup": triangle_soup, "instanced": instanced_grid}
