"""Scenes and camera blocks aimed at the places where the closest-hit search + proof (pt_fast.cuh) could part from the
reference traversal (main.glsl:259-350): equal-t ties (quirk Q8, main.glsl:247), instances that coincide, rays with zero
direction components (1/0 = inf, 0*inf = NaN in intersectAABB, main.glsl:259-268), triangles the |det| < 1e-5 test
rejects (main.glsl:232), a TLAS whose root is a leaf (main.glsl:309), a tiny instance far from the camera (margins of the
search's boxes against Moller-Trumbore rounding).  Used by the CPU tier (device functions compiled for the host, the
restatement against the reference's shader text) and by the GPU tier (every rendering schedule against the oracle)."""
import numpy as np

from gdpathtracing_b200 import _lib, nodes, scenes


def _surface(tris, normal=(0, 0, 1)):
    p = np.asarray(tris, np.float32).reshape(-1, 3)
    n = np.tile(np.asarray(normal, np.float32), (len(p), 1))
    uv = np.tile(np.array([[0, 0], [1, 0], [0, 1]], np.float32), (len(p) // 3, 1))
    return {"positions": p, "normals": n, "uvs": uv, "indices": np.arange(len(p), dtype=np.int32)}


def _doubled(surface):
    """Every triangle twice, the copy right behind the original in the triangle list."""
    p = surface["positions"][surface["indices"]].reshape(-1, 3, 3)
    n = surface["normals"][surface["indices"]].reshape(-1, 3, 3)
    t = surface["uvs"][surface["indices"]].reshape(-1, 3, 2)
    rep = lambda a: np.repeat(a, 2, axis=0).reshape(-1, a.shape[-1])
    p2 = rep(p)
    return {"positions": p2, "normals": rep(n), "uvs": rep(t), "indices": np.arange(len(p2), dtype=np.int32)}


def coplanar_duplicates():
    """The Cornell room and a box inside it with every triangle stored twice: every hit is an equal-t tie, the reference
    keeps the copy it tests last."""
    sc = scenes.SceneDesc("coplanar_duplicates", camera_transform12=scenes.transform12(None, (0.0, 0.0, 9.5)), fov=50.0)
    sc.materials = [dict(albedo=(0.8, 0.8, 0.8), roughness=0.9), dict(albedo=(0.9, 0.3, 0.2), roughness=0.3, metallic=0.7),
                    dict(albedo=(1, 1, 1), emission=(1, 1, 1), emission_energy_multiplier=6.0)]
    sc.default_material = 0
    room = [_doubled(s) for s in scenes._cornell_room()]
    box = [_doubled(s) for s in scenes._box_five_faces((0.8, 0.8, 0.8))]
    sc.meshes = [room, box]
    sc.instances = [dict(mesh=0, transform12=scenes.ROOM_TRANSFORM, surface_overrides=[0, 1, 0]),
                    dict(mesh=1, transform12=scenes.transform12(None, (0.4, -2.2, 0.3)), surface_overrides=[1]),
                    dict(mesh=1, transform12=scenes.LIGHT_TRANSFORM, surface_overrides=[2])]
    return sc


def coincident_instances():
    """Two instances of one BLAS at the same transform (and a third elsewhere): equal t from two different instances,
    the TLAS visiting order decides which instance id the reference reports."""
    sc = scenes.instanced_grid(2, 300, seed=9)
    first = sc.instances[0]
    sc.instances.insert(5, dict(mesh=0, transform12=first["transform12"].copy(), surface_overrides=[1]))
    sc.instances.append(dict(mesh=0, transform12=sc.instances[3]["transform12"].copy(), surface_overrides=[0]))
    sc.name = "coincident_instances"
    return sc


def decal_on_a_wall():
    """A quad of a second instance lying exactly in the room's back wall (a decal, an object resting on a floor): equal t
    across two instances.  The reference then reports the triangle, position and facing of the pair it tested last but the
    instance -- transform and materials -- of the pair that lowered t first (main.glsl:247-255 against :322-325)."""
    sc = scenes.SceneDesc("decal_on_a_wall", camera_transform12=scenes.transform12(None, (0.3, 0.2, 8.0)), fov=50.0)
    sc.materials = [dict(albedo=(0.8, 0.8, 0.8), roughness=0.9), dict(albedo=(0.9, 0.3, 0.2), roughness=0.3, metallic=0.7),
                    dict(albedo=(0.2, 0.4, 0.9), roughness=0.5)]
    sc.default_material = 0
    quad = [scenes._merge([scenes._quad_surface([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], [0, 0, 1])])]
    # the room's own back wall (cornell.obj usemtl 2, z = -5) as a mesh of its own, moved by 4 units inside its plane; the
    # instance's origin moves it back, so both instances test triangles of the same world plane from different spaces
    wall = [scenes._merge([scenes._quad_surface([[-5, -1, -5], [5, -1, -5], [5, 9, -5], [-5, 9, -5]], [0, 0, 1])])]
    shifted = scenes.ROOM_TRANSFORM.copy()
    shifted[10] = np.float32(-2.4)  # origin.y: 0.6 * (-4)
    sc.meshes = [scenes._cornell_room(), quad, wall]
    sc.instances = [dict(mesh=1, transform12=scenes.transform12([[1.5, 0, 0], [0, 1.2, 0], [0, 0, 1]], (0.5, 0.0, -3.0)), surface_overrides=[1]),
                    dict(mesh=0, transform12=scenes.ROOM_TRANSFORM, surface_overrides=[0, 2, 0]),
                    dict(mesh=1, transform12=scenes.transform12([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], (-3.0, -1.0, 0.0)), surface_overrides=[2]),
                    dict(mesh=2, transform12=shifted, surface_overrides=[1])]
    return sc


def degenerate_cluster():
    """Triangles the determinant test throws away next to ones it keeps: needles and specks whose |det| < 1e-5, zero-area
    triangles, and walls seen edge-on, in front of an ordinary backdrop."""
    sc = scenes.SceneDesc("degenerate_cluster", camera_transform12=scenes.transform12(None, (0.0, 0.0, 6.0)), fov=55.0)
    sc.materials = [dict(albedo=(0.7, 0.7, 0.7), roughness=0.8), dict(albedo=(0.2, 0.6, 0.9), roughness=0.2, metallic=0.9)]
    sc.default_material = 0
    r = scenes.splitmix64_floats(21, 4000 * 12).reshape(4000, 12)
    centre = (r[:, 0:3] * 2 - 1) * np.float32(2.0)
    size = np.float32(10.0) ** (-(r[:, 3:4] * 4 + 1))                      # edge length 1e-1 .. 1e-5
    specks = (centre[:, None, :] + (r[:, 3:12].reshape(-1, 3, 3) * 2 - 1) * size[:, None, :]).astype(np.float32)
    needles = specks[:1000].copy()
    needles[:, 1] = needles[:, 0] + np.float32([1.5, 0, 0])                 # long and thin
    needles[:, 2] = needles[:, 0] + np.float32([0.75, 1e-6, 0])
    zero_area = specks[1000:1200].copy()
    zero_area[:, 2] = zero_area[:, 1]
    edge_on = []                                                             # walls containing the camera's z axis
    for k in range(40):
        y = -2.0 + 0.1 * k
        edge_on += [[[0.0, y, -3.0], [0.0, y, 3.0], [0.0, y + 0.05, 3.0]], [[-3.0, 0.0, y], [3.0, 0.0, y], [3.0, 0.0, y + 0.05]]]
    backdrop = [[[-4, -4, -3], [4, -4, -3], [4, 4, -3]], [[-4, -4, -3], [4, 4, -3], [-4, 4, -3]]]
    sc.meshes = [[_surface(np.concatenate([specks, needles, zero_area, np.array(edge_on, np.float32)])), _surface(backdrop)]]
    sc.instances = [dict(mesh=0, surface_overrides=[1, 0]),
                    dict(mesh=0, transform12=scenes.transform12([[0.6, 0.0, 0.8], [0.0, 1.0, 0.0], [-0.8, 0.0, 0.6]], (0.3, 0.2, -0.5)),
                         surface_overrides=[0, 1])]
    return sc


def tlas_root_is_a_leaf():
    """One rotated, scaled instance: the TLAS root is its only node and is never box-tested (main.glsl:309-314)."""
    sc = scenes.triangle_soup(3000, seed=13)
    sc.name = "tlas_root_is_a_leaf"
    sc.meshes = [[scenes._soup_surface(3000, 13, 10.0, 0.7)]]
    c, s = np.float32(np.cos(0.7)), np.float32(np.sin(0.7))
    sc.instances = [dict(mesh=0, transform12=scenes.transform12([[0.5 * c, -0.5 * s, 0.0], [0.5 * s, 0.5 * c, 0.0], [0.0, 0.0, 0.5]], (0.5, -1.0, 2.0)))]
    return sc


def far_tiny_instance():
    """A 2 cm object seen from 4 000 of its diameters away next to an ordinary one: beyond the reach the search's culling
    margins are built for (derived_layout.h fast_reach), so those rays must take the exact traversal."""
    sc = scenes.SceneDesc("far_tiny_instance", camera_transform12=scenes.transform12(None, (0.0, 0.0, 80.0)), fov=0.05)
    sc.materials = [dict(albedo=(0.8, 0.5, 0.3), roughness=0.5, metallic=0.3)]
    sc.default_material = 0
    blob = scenes._soup_surface(600, 5, 0.008, 0.004)
    sc.meshes = [[blob], scenes._box_five_faces((0.02, 0.02, 0.02))]
    sc.instances = [dict(mesh=0), dict(mesh=1, transform12=scenes.transform12(None, (0.02, 0.0, -0.1)))]
    return sc


def small_prop_in_a_large_room():
    """A 4 cm object on the floor of the 6 m room, seen from across it: 150 of its own diameters away, far beyond what a
    margin of extent / 4096 covers by itself.  The margins of such a prop are built on a larger effective extent
    (derived_layout.h root_extent), so its rays stay with the search -- no cliff into the exact traversal."""
    sc = scenes.SceneDesc("small_prop_in_a_large_room", camera_transform12=scenes.transform12(None, (0.0, -2.0, 2.9)), fov=12.0)
    sc.materials = [dict(albedo=(0.8, 0.8, 0.8), roughness=0.9), dict(albedo=(0.9, 0.5, 0.2), roughness=0.4, metallic=0.5)]
    sc.default_material = 0
    sc.meshes = [scenes._cornell_room(), [scenes._soup_surface(400, 17, 0.015, 0.006)]]
    sc.instances = [dict(mesh=0, transform12=scenes.ROOM_TRANSFORM),
                    dict(mesh=1, transform12=scenes.transform12(None, (0.0, -2.97, -2.9)), surface_overrides=[1])]
    return sc


def axis_aligned_boxes():
    """Boxes whose faces lie in coordinate planes through the origin, seen by the degenerate cameras below."""
    sc = scenes.SceneDesc("axis_aligned_boxes", camera_transform12=scenes.transform12(None, (0.0, 0.0, 7.0)), fov=50.0)
    sc.materials = [dict(albedo=(0.8, 0.8, 0.8), roughness=0.05, metallic=1.0), dict(albedo=(0.3, 0.8, 0.4), roughness=0.7)]
    sc.default_material = 0
    sc.meshes = [scenes._box_five_faces((1.0, 1.0, 1.0)), scenes._cornell_room()]
    sc.instances = [dict(mesh=0, transform12=scenes.transform12(None, (1.0, 0.0, 0.0))),     # a face in the plane x = 0
                    dict(mesh=0, transform12=scenes.transform12(None, (-1.0, 1.0, -2.0))),   # faces in x = 0 and y = 0
                    dict(mesh=0, transform12=scenes.transform12(None, (0.0, -3.0, 1.0)), surface_overrides=[1]),
                    dict(mesh=1, transform12=scenes.ROOM_TRANSFORM, surface_overrides=[1, 0, 1])]
    return sc


def planar_camera(sc, W, H, frame_index, zero_axes):
    """The scene's camera block with the rows of the inverse view-projection that produce the world x (and y) coordinate of
    the far point zeroed and the camera moved into those planes: every camera ray has direction component(s) exactly 0
    (main.glsl:414-421 with such a matrix), so rD holds inf and the slab tests meet 0 * inf at the box faces in x = 0."""
    c = nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, frame_index)
    for a in zero_axes:
        for col in range(4):
            c.ivp[col * 4 + a] = 0.0
        c.position[a] = 0.0
    return c


# (name, scene, W, H, depth, camera: None = the scene's own, else the axes whose direction component is forced to 0)
CASES = [
    ("coplanar_duplicates", coplanar_duplicates, 96, 96, 5, None),
    ("coincident_instances", coincident_instances, 128, 72, 4, None),
    ("decal_on_a_wall", decal_on_a_wall, 128, 96, 5, None),
    ("degenerate_cluster", degenerate_cluster, 128, 96, 4, None),
    ("tlas_root_is_a_leaf", tlas_root_is_a_leaf, 96, 64, 3, None),
    ("far_tiny_instance", far_tiny_instance, 96, 64, 3, None),
    ("small_prop_in_a_large_room", small_prop_in_a_large_room, 96, 64, 4, None),
    ("rays_in_the_plane_x0", axis_aligned_boxes, 64, 96, 4, (0,)),
    ("rays_along_minus_z", axis_aligned_boxes, 48, 32, 4, (0, 1)),
]
IDS = [c[0] for c in CASES]


def camera_for(sc, W, H, frame_index, zero_axes):
    if zero_axes is None:
        return nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, frame_index)
    return planar_camera(sc, W, H, frame_index, zero_axes)


assert _lib  # the ctypes Camera structure the blocks above are instances of
