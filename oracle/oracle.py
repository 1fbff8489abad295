"""TEST INFRASTRUCTURE ONLY -- ctypes faces of the CPU checkers.

  libgdpt_oracle.so        our restatement of main.glsl / brdfs.glsl / progressive_rendering.glsl
                           (oracle/pt_oracle.cpp)
  _ref/libgdpt_refbvh.so   the REFERENCE's own src/bvh/bvh.cpp compiled in place (oracle/ref_bridge.cpp)
  _ref/libgdpt_refshader.so  the REFERENCE's own shader text compiled as C++ (oracle/ref_shader_bridge.cpp,
                           oracle/glsl_shim/): what pins the restatement to the reference

Never imported by the product package; used by tests/, smoke() and bench.py's CPU-baseline legs.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int32, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libgdpt_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libgdpt_refbvh.so")
REF_SHADER_LIB = os.path.join(_HERE, "_ref", "libgdpt_refshader.so")

TRACE_DTYPE = [("hit", "<u4"), ("triangle", "<u4"), ("blas", "<u4"), ("front", "<u4"), ("t", "<f4"), ("u", "<f4"),
               ("v", "<f4"), ("node_pops", "<u4"), ("box_tests", "<u4"), ("tri_tests", "<u4"), ("tlas_leaves", "<u4"),
               ("max_stack", "<u4"), ("visit_hash_lo", "<u4"), ("visit_hash_hi", "<u4")]


class OrcScene(Structure):
    _fields_ = [("tri_geom", c_void_p), ("n_tris", c_uint64), ("tri_data", c_void_p), ("materials", c_void_p),
                ("n_materials", c_uint64), ("bvh", c_void_p), ("n_nodes", c_uint64), ("blas", c_void_p),
                ("n_blas", c_uint64), ("tlas", c_void_p), ("n_tlas", c_uint64), ("textures", c_void_p),
                ("tex_w", c_int32), ("tex_h", c_int32), ("tex_layers", c_int32), ("material_ext", c_int32),
                ("surface_materials", c_void_p)]


class OrcStats(Structure):
    _fields_ = [("rays", c_uint64), ("primary_hits", c_uint64), ("node_pops", c_uint64), ("box_tests", c_uint64),
                ("tri_tests", c_uint64), ("tlas_leaves", c_uint64), ("max_stack", c_uint32), ("stack_overflow", c_uint32)]


_orc = None
_ref = None


def lib():
    global _orc
    if _orc is None:
        _orc = ctypes.CDLL(ORACLE_LIB)
        _orc.orc_path_trace.restype = c_int
        _orc.orc_path_trace.argtypes = [POINTER(OrcScene), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_int, c_void_p, c_uint32, POINTER(OrcStats), c_void_p]
        _orc.orc_trace_rays.restype = None
        _orc.orc_trace_rays.argtypes = [POINTER(OrcScene), c_uint64, c_void_p, c_void_p, c_void_p]
        _orc.orc_progressive.restype = None
        _orc.orc_progressive.argtypes = [c_void_p, c_void_p, c_int, c_int, c_uint32]
        _orc.orc_temporal.restype = None
        _orc.orc_temporal.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        _orc.orc_mat4_mul.argtypes = [c_void_p, c_void_p, c_void_p]
        _orc.orc_mat4_inverse.argtypes = [c_void_p, c_void_p]
        _orc.orc_prng_seed.argtypes = [c_uint32, c_uint32, c_uint32, c_void_p]
        _orc.orc_pcg2d.argtypes = [c_void_p, c_void_p]
        _orc.orc_sincosf.argtypes = [c_float, c_void_p]
        _orc.orc_hardware_threads.restype = ctypes.c_uint
    return _orc


def ref_available():
    return os.path.exists(REF_LIB)


def ref_shader_available():
    return os.path.exists(REF_SHADER_LIB)


_refsh = None


def ref_shader():
    """The reference's shader text compiled as C++ (same entry-point shapes as the restatement's)."""
    global _refsh
    if _refsh is None:
        _refsh = ctypes.CDLL(REF_SHADER_LIB)
        _refsh.refsh_path_trace.restype = c_int
        _refsh.refsh_path_trace.argtypes = [POINTER(OrcScene), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                            c_void_p, c_void_p, c_int, c_void_p, c_uint32, POINTER(OrcStats), c_void_p]
        _refsh.refsh_trace_rays.restype = None
        _refsh.refsh_trace_rays.argtypes = [POINTER(OrcScene), c_uint64, c_void_p, c_void_p, c_void_p]
        _refsh.refsh_progressive.restype = None
        _refsh.refsh_progressive.argtypes = [c_void_p, c_void_p, c_int, c_int, c_uint32]
        _refsh.refsh_temporal.restype = None
        _refsh.refsh_temporal.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        _refsh.refsh_prng_seed.argtypes = [c_uint32, c_uint32, c_uint32, c_void_p]
        _refsh.refsh_pcg2d.argtypes = [c_void_p, c_void_p]
    return _refsh


def ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_LIB)
        _ref.refbvh_new.restype = c_void_p
        _ref.refbvh_free.argtypes = [c_void_p]
        _ref.refbvh_add_mesh.restype = c_uint32
        _ref.refbvh_add_mesh.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        _ref.refbvh_add_instance.argtypes = [c_void_p, c_uint32, c_void_p, c_int, c_void_p]
        _ref.refbvh_build_tlas.argtypes = [c_void_p]
        for n in ("node", "triangle", "instance", "tlas"):
            getattr(_ref, f"refbvh_{n}_count").restype = c_uint64
            getattr(_ref, f"refbvh_{n}_count").argtypes = [c_void_p]
        for n in ("nodes", "triangles", "instances", "tlas"):
            getattr(_ref, f"refbvh_copy_{n}").argtypes = [c_void_p, c_void_p]
        _ref.refbvh_sizeof.restype = c_uint32
    return _ref


def hardware_threads():
    return int(lib().orc_hardware_threads())


def _p(a):
    return a.ctypes.data_as(c_void_p)


class Scene:
    """Flat scene buffers (bytes, exactly what crosses the boundary) held alive for the oracle."""

    def __init__(self, buffers, texture_layers):
        self.np = {k: np.frombuffer(v, np.uint8).copy() for k, v in buffers.items()}
        layers = np.stack(texture_layers).astype(np.uint8)
        self.tex = np.ascontiguousarray(layers)
        s = OrcScene()
        s.tri_geom = _p(self.np["triangles_geometry"]); s.n_tris = len(self.np["triangles_geometry"]) // 48
        s.tri_data = _p(self.np["triangles_data"])
        s.materials = _p(self.np["materials"]); s.n_materials = len(self.np["materials"]) // 64
        s.bvh = _p(self.np["bvh"]); s.n_nodes = len(self.np["bvh"]) // 48
        s.blas = _p(self.np["blas"]); s.n_blas = len(self.np["blas"]) // 176
        s.tlas = _p(self.np["tlas"]); s.n_tlas = len(self.np["tlas"]) // 32
        s.textures = _p(self.tex)
        s.tex_layers, s.tex_h, s.tex_w = self.tex.shape[0], self.tex.shape[1], self.tex.shape[2]
        # material-breadth extension (gdpt_wire.h, "#define GDPT_MATERIAL_EXT"): present iff the group emitted its table
        s.material_ext = 1 if "surface_materials" in self.np else 0
        s.surface_materials = _p(self.np["surface_materials"]) if s.material_ext else None
        self.c = s


def path_trace(scene, width, height, camera_bytes, max_depth=5, debug_steps=False, threads=None, rows=None,
               trace_segments=0, visits_per_ray=0, row_step=1, radiance=False, impl="restatement", counters=True):
    """K1 on the CPU.  Returns dict(rgba8, depth, stats, trace, visits[, radiance]).

    impl="restatement": oracle/pt_oracle.cpp.  impl="reference": the reference's own main.glsl (+ brdfs.glsl)
    compiled as C++ (oracle/_ref/libgdpt_refshader.so); its trace records leave t / u / v / front / max_stack
    zero (use trace_rays for those); counters=False (timing runs) keeps only the ray count of its stats."""
    threads = threads or hardware_threads()
    params = np.zeros(9, np.uint32)
    params[4], params[5] = width, height
    cam = np.frombuffer(bytes(camera_bytes), np.uint8).copy()
    out = np.zeros((height, width, 4), np.uint8)
    depth = np.zeros((height, width), np.float32)
    trace = np.zeros((trace_segments, height * width), dtype=TRACE_DTYPE) if trace_segments else None
    visits = np.full((height * width, visits_per_ray), 0xFFFFFFFF, np.uint32) if visits_per_ray else None
    raw = np.zeros((height, width, 4), np.float32) if radiance else None
    st = OrcStats()
    y0, y1 = rows if rows else (0, height)
    fn = lib().orc_path_trace if impl == "restatement" else ref_shader().refsh_path_trace
    flags = (1 if debug_steps else 0) | (2 if (counters and impl != "restatement") else 0)
    fn(ctypes.byref(scene.c), _p(params), _p(cam), max_depth, flags, threads, y0, y1, row_step,
       _p(out), _p(depth), _p(trace) if trace is not None else None, trace_segments,
       _p(visits) if visits is not None else None, visits_per_ray, ctypes.byref(st), _p(raw) if raw is not None else None)
    stats = {k: getattr(st, k) for k, _ in OrcStats._fields_}
    return dict(rgba8=out, depth=depth, stats=stats, trace=trace, visits=visits, radiance=raw)


def trace_rays(scene, origins, directions, impl="restatement"):
    """ray_trace_tlas (main.glsl:305-350) on the given rays; one TRACE_DTYPE record per ray."""
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
    out = np.zeros(len(o), dtype=TRACE_DTYPE)
    fn = lib().orc_trace_rays if impl == "restatement" else ref_shader().refsh_trace_rays
    fn(ctypes.byref(scene.c), len(o), _p(o), _p(d), _p(out))
    return out


def progressive(screen_rgba8, accum_rgba32f, frame_count, impl="restatement"):
    h, w = screen_rgba8.shape[:2]
    fn = lib().orc_progressive if impl == "restatement" else ref_shader().refsh_progressive
    fn(_p(screen_rgba8), _p(accum_rgba32f), w, h, frame_count)


def temporal(params_bytes, screen_rgba8, depth, fb1, fb2, impl="restatement"):
    """K3 on the CPU (temporal_reprojection.glsl:31-71).  params_bytes: the 88 B Params block; screen (H, W, 4) u8
    in/out; depth (H, W) f32; fb1 / fb2 (H, W, 4) f32 ping-pong buffers, one read and the other written."""
    par = np.frombuffer(bytes(params_bytes), np.uint8).copy()
    assert par.size == 88 and depth.dtype == np.float32 and fb1.dtype == np.float32 and fb2.dtype == np.float32
    fn = lib().orc_temporal if impl == "restatement" else ref_shader().refsh_temporal
    fn(_p(par), _p(screen_rgba8), _p(depth), _p(fb1), _p(fb2))


def mat4_mul(a, b):
    """Projection::operator* (godot-cpp projection.cpp:709-723) on column-major 16-float matrices."""
    a, b = np.ascontiguousarray(a, np.float32).reshape(16), np.ascontiguousarray(b, np.float32).reshape(16)
    out = np.zeros(16, np.float32)
    lib().orc_mat4_mul(_p(a), _p(b), _p(out))
    return out


def mat4_inverse(a):
    """Projection::inverse (godot-cpp projection.cpp:601-698)."""
    a = np.ascontiguousarray(a, np.float32).reshape(16)
    out = np.zeros(16, np.float32)
    lib().orc_mat4_inverse(_p(a), _p(out))
    return out


def temporal_delta(previous_vp, vp):
    """temporal_reprojection.cpp:58-61: `Transform3D deltaMatrix = previous_vp * vp.inverse()` keeps the affine
    part only (projection.cpp:886-907), then goes back to 16 floats with a (0, 0, 0, 1) bottom row (:916-936)."""
    d = mat4_mul(previous_vp, mat4_inverse(vp)).copy()
    d[3] = d[7] = d[11] = 0.0
    d[15] = 1.0
    return d


def prng_seed(px, py, frame):
    out = np.zeros(2, np.uint32)
    lib().orc_prng_seed(px, py, frame, _p(out))
    return out


def pcg2d(state):
    st = np.array(state, np.uint32)
    out = np.zeros(2, np.float32)
    lib().orc_pcg2d(_p(st), _p(out))
    return st, out


def sincos(x):
    out = np.zeros(2, np.float32)
    lib().orc_sincosf(float(np.float32(x)), _p(out))
    return out


# ------------------------------------------------------------------------- reference builder
def reference_build(scene_desc):
    """Run the REFERENCE's BuildBVH / BLASInstance::set_transform / TLAS::build on a SceneDesc
    (meshes de-duplicated and instances resolved the way GeometryGroup3D::build does).  Returns
    the raw arrays: nodes (48 B), triangles (144 B), instances (176 B), tlas (32 B)."""
    r = ref()
    h = r.refbvh_new()
    try:
        mesh_order, roots = [], {}
        for inst in scene_desc.instances:
            if inst["mesh"] not in roots:
                roots[inst["mesh"]] = None
                mesh_order.append(inst["mesh"])
        for m in mesh_order:
            surfaces = scene_desc.meshes[m]
            vc = np.array([len(s["positions"]) for s in surfaces], np.int32)
            ic = np.array([len(s["indices"]) for s in surfaces], np.int32)
            pos = np.ascontiguousarray(np.concatenate([np.asarray(s["positions"], np.float32).reshape(-1, 3) for s in surfaces]))
            nrm = np.ascontiguousarray(np.concatenate([np.asarray(s["normals"], np.float32).reshape(-1, 3) for s in surfaces]))
            uv = np.ascontiguousarray(np.concatenate([np.asarray(s["uvs"], np.float32).reshape(-1, 2) for s in surfaces]))
            idx = np.ascontiguousarray(np.concatenate([np.asarray(s["indices"], np.int32) for s in surfaces]))
            roots[m] = r.refbvh_add_mesh(h, len(surfaces), _p(vc), _p(ic), _p(pos), _p(nrm), _p(uv), _p(idx))
        return h, roots
    except Exception:
        r.refbvh_free(h)
        raise


def reference_arrays(scene_desc, material_ids_per_instance):
    r = ref()
    h, roots = reference_build(scene_desc)
    try:
        for inst, mids in zip(scene_desc.instances, material_ids_per_instance):
            t = np.ascontiguousarray(inst.get("transform12", [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]), np.float32)
            ids = np.ascontiguousarray(mids, np.int32)
            r.refbvh_add_instance(h, roots[inst["mesh"]], _p(ids), len(ids), _p(t))
        r.refbvh_build_tlas(h)
        out = {}
        for name, cnt, size in (("nodes", r.refbvh_node_count(h), 48), ("triangles", r.refbvh_triangle_count(h), 144),
                                ("instances", r.refbvh_instance_count(h), 176), ("tlas", r.refbvh_tlas_count(h), 32)):
            buf = np.zeros(cnt * size, np.uint8)
            getattr(r, f"refbvh_copy_{name}")(h, _p(buf))
            out[name] = buf
        return out
    finally:
        r.refbvh_free(h)
