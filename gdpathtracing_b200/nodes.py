"""Python faces of the host-layer twins of the reference's two scene nodes.

``GeometryGroup3D`` and ``PathTracingCamera`` here are thin ctypes wrappers over
libgdpt_host.so (C++), which in turn drives libgdpt_cuda.so through the C-ABI.  Property names
and the init()/render() life cycle are the reference's
(src/path_tracing/geometry_group3d.h:73-104, src/path_tracing/path_tracing_camera.h:56-111).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import cuda, host

BUFFER_NAMES = ["triangles_geometry", "triangles_data", "materials", "bvh", "blas", "tlas"]
RECORD_SIZES = [48, 80, 64, 48, 176, 32]

IDENTITY12 = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class GeometryGroup3D:
    """Twin of the reference node; resources are registered explicitly instead of being found in a scene tree."""

    def __init__(self):
        self._h = host.gdpt_group_create()
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            host.gdpt_group_destroy(self._h)
            self._h = None

    # resources ---------------------------------------------------------------------------
    def add_texture(self, rgba8):
        img = np.ascontiguousarray(rgba8, dtype=np.uint8)
        assert img.ndim == 3 and img.shape[2] == 4
        return host.gdpt_group_add_texture(self._h, _ptr(img), img.shape[1], img.shape[0])

    def add_material(self, albedo=(1.0, 1.0, 1.0), metallic=0.0, roughness=1.0, emission=(0.0, 0.0, 0.0),
                     emission_energy_multiplier=1.0, albedo_texture=-1, is_standard=True, roughness_texture=-1,
                     metallic_texture=-1, albedo_srgb=False):
        """The StandardMaterial3D fields the reference converts (geometry_group3d.cpp:271-292); roughness_texture,
        metallic_texture and albedo_srgb are the material-breadth extension, converted only while material_ext is on."""
        m = _lib.StandardMaterial()
        m.albedo[:] = albedo
        m.metallic, m.roughness = metallic, roughness
        m.emission[:] = emission
        m.emission_energy_multiplier = emission_energy_multiplier
        m.albedo_texture = albedo_texture
        m.is_standard = 1 if is_standard else 0
        if roughness_texture >= 0 or metallic_texture >= 0 or albedo_srgb:
            return host.gdpt_group_add_material_ext(self._h, ctypes.byref(m), int(roughness_texture), int(metallic_texture),
                                                    1 if albedo_srgb else 0)
        return host.gdpt_group_add_material(self._h, ctypes.byref(m))

    @property
    def material_ext(self):
        """Ours: roughness / metallic textures, sRGB albedo layers, any number of materials per instance."""
        return host.gdpt_group_get_material_ext(self._h) == 1

    @material_ext.setter
    def material_ext(self, on):
        host.gdpt_group_set_material_ext(self._h, 1 if on else 0)

    def add_mesh(self, surfaces):
        """surfaces: list of dicts with positions [n,3], normals [n,3], uvs [n,2], indices [m] (int32)."""
        vc = np.array([len(s["positions"]) for s in surfaces], dtype=np.int32)
        ic = np.array([len(s["indices"]) for s in surfaces], dtype=np.int32)
        pos = _f32(np.concatenate([np.asarray(s["positions"], np.float32).reshape(-1, 3) for s in surfaces]))
        nrm = _f32(np.concatenate([np.asarray(s["normals"], np.float32).reshape(-1, 3) for s in surfaces]))
        uv = _f32(np.concatenate([np.asarray(s["uvs"], np.float32).reshape(-1, 2) for s in surfaces]))
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s["indices"], np.int32).reshape(-1) for s in surfaces]),
                                   dtype=np.int32)
        return host.gdpt_group_add_mesh(self._h, len(surfaces), _ptr(vc), _ptr(ic), _ptr(pos), _ptr(nrm), _ptr(uv),
                                        _ptr(idx))

    def add_mesh_instance(self, mesh, transform12=IDENTITY12, material_override=-1, surface_overrides=()):
        t = _f32(transform12).reshape(12)
        ov = np.ascontiguousarray(surface_overrides, dtype=np.int32)
        host.gdpt_group_add_mesh_instance(self._h, mesh, _ptr(t), material_override, _ptr(ov) if len(ov) else None,
                                          len(ov))

    # properties --------------------------------------------------------------------------
    def set_default_material(self, handle):
        host.gdpt_group_set_default_material(self._h, handle)

    @property
    def texture_array_resolution(self):
        return host.gdpt_group_get_texture_array_resolution(self._h)

    @texture_array_resolution.setter
    def texture_array_resolution(self, v):
        host.gdpt_group_set_texture_array_resolution(self._h, int(v))

    @property
    def build_threads(self):
        """Threads of the BLAS build (ours; upstream is single-threaded): 0 = all, 1 = one.  Same bytes either way."""
        return host.gdpt_group_get_build_threads(self._h)

    @build_threads.setter
    def build_threads(self, v):
        host.gdpt_group_set_build_threads(self._h, int(v))

    # build + buffers ---------------------------------------------------------------------
    def build(self):
        host.gdpt_group_build(self._h)
        return host.gdpt_group_last_build_seconds(self._h)

    def buffer(self, which):
        """Copy of one GPU buffer as bytes (get_*_buffer of the reference)."""
        i = BUFFER_NAMES.index(which) if isinstance(which, str) else which
        size = host.gdpt_group_buffer_size(self._h, i)
        data = host.gdpt_group_buffer_data(self._h, i)
        return ctypes.string_at(data, size) if size else b""

    def buffers(self):
        out = {n: self.buffer(n) for n in BUFFER_NAMES}
        if self.material_ext:
            out["surface_materials"] = self.buffer(6)  # set 1 binding 6 of a GDPT_MATERIAL_EXT shader (gdpt_wire.h)
        return out

    def texture_layers(self):
        res = self.texture_array_resolution
        n = host.gdpt_group_texture_layer_count(self._h)
        return [np.frombuffer(ctypes.string_at(host.gdpt_group_texture_layer(self._h, l), res * res * 4), np.uint8)
                .reshape(res, res, 4).copy() for l in range(n)]

    def counts(self):
        return {n: host.gdpt_group_buffer_size(self._h, i) // RECORD_SIZES[i] for i, n in enumerate(BUFFER_NAMES)}


class PathTracingCamera:
    PROGRESSIVE_RENDERING, TEMPORAL_REPROJECTION, NONE = 0, 1, 2

    def __init__(self):
        self._h = host.gdpt_camera_create()
        self._group = None
        self._w = self._h_px = 0
        self._ready = False

    def __del__(self):
        if getattr(self, "_h", None):
            host.gdpt_camera_destroy(self._h)
            self._h = None

    # reference properties ----------------------------------------------------------------
    @property
    def fov(self):
        return host.gdpt_camera_get_fov(self._h)

    @fov.setter
    def fov(self, v):
        host.gdpt_camera_set_fov(self._h, float(v))

    @property
    def denoising_mode(self):
        return host.gdpt_camera_get_denoising_mode(self._h)

    @denoising_mode.setter
    def denoising_mode(self, m):
        host.gdpt_camera_set_denoising_mode(self._h, int(m))

    @property
    def geometry_group(self):
        return self._group

    @geometry_group.setter
    def geometry_group(self, g):
        self._group = g
        host.gdpt_camera_set_geometry_group(self._h, g._h if g is not None else None)

    # engine services / backend parameters -------------------------------------------------
    def set_window_size(self, w, h):
        self._w, self._h_px = int(w), int(h)
        host.gdpt_camera_set_window_size(self._h, int(w), int(h))

    def set_global_transform(self, transform12):
        t = _f32(transform12).reshape(12)
        host.gdpt_camera_set_global_transform(self._h, _ptr(t))

    def set_max_depth(self, d):
        host.gdpt_camera_set_max_depth(self._h, int(d))

    def set_cuda_device(self, ordinal):
        host.gdpt_camera_set_cuda_device(self._h, int(ordinal))

    def set_frame_index(self, f):
        host.gdpt_camera_set_frame_index(self._h, int(f))

    def set_shard(self, part, parts, band_rows=32):
        host.gdpt_camera_set_shard(self._h, int(part), int(parts), int(band_rows))

    def set_trace(self, segments, visits_per_ray=0):
        self._trace_segments, self._visits = int(segments), int(visits_per_ray)
        host.gdpt_camera_set_trace(self._h, int(segments), int(visits_per_ray))

    def set_debug_steps(self, on):
        host.gdpt_camera_set_debug_steps(self._h, 1 if on else 0)

    def set_variant(self, variant):
        """Kernel schedule (include/gdpt.h GDPT_VARIANT); -1 backend default.  Results are identical for every value."""
        host.gdpt_camera_set_variant(self._h, int(variant))

    def set_tuning(self, name, value):
        """Scheduling knob of the path kernels ("#define GDPT_TUNE_<NAME> n"); A/B measurements only, results identical."""
        host.gdpt_camera_set_tuning(self._h, str(name).upper().encode(), int(value))

    def set_count_work(self, on):
        """The closest-hit path kernel also counts the work it executes itself (stats own_*); measurement aid."""
        host.gdpt_camera_set_count_work(self._h, 1 if on else 0)

    def set_record_hits(self, segments):
        """Hit records of the first n segments from the rendering kernels (no work counters)."""
        host.gdpt_camera_set_record_hits(self._h, int(segments))

    def set_cull(self, mode):
        """-1 backend default, 0 reference visit order, 1 tight-box culling (bit-identical results)."""
        host.gdpt_camera_set_cull(self._h, int(mode))

    def set_fused_frame(self, on):
        host.gdpt_camera_set_fused_frame(self._h, 1 if on else 0)

    # life cycle -----------------------------------------------------------------------------
    def init(self):
        self._ready = host.gdpt_camera_init(self._h) == 1
        if not self._ready:
            raise _lib.GdptError("PathTracingCamera.init failed: "
                                 + (cuda.gdpt_last_error(host.gdpt_camera_device(self._h)) or b"").decode())
        return True

    def render(self):
        """One frame; returns the RGBA8 image (H, W, 4) the reference would push into its TextureRect."""
        host.gdpt_camera_render(self._h)
        return self.output_image()

    def render_device_only(self):
        host.gdpt_camera_render_device_only(self._h)

    def render_begin(self):
        """Pipelined render(): enqueue one frame including its read-back; at most MAX_FRAMES_IN_FLIGHT (4) may be
        in flight."""
        if host.gdpt_camera_render_begin(self._h) != 1:
            raise _lib.GdptError("render_begin failed: " + (cuda.gdpt_last_error(self.device) or b"").decode())

    def render_wait(self):
        """Block for the oldest frame in flight.  Returns (image, stats): the image is a zero-copy view of the
        page-locked buffer the frame was read back into, valid until three more frames have been begun."""
        st = _lib.FrameStats()
        p = host.gdpt_camera_render_wait(self._h, ctypes.byref(st))
        if not p:
            raise _lib.GdptError("render_wait failed: " + (cuda.gdpt_last_error(self.device) or b"").decode())
        return self._pinned_view(p), {k: getattr(st, k) for k, _ in _lib.FrameStats._fields_}

    def _pinned_view(self, address):
        """numpy view of one of this camera's page-locked frame buffers; the view keeps the camera alive."""
        n = self._w * self._h_px * 4
        buf = (ctypes.c_uint8 * n).from_address(address)
        buf._owner = self  # the array's base is `buf`: the buffers are not freed under a live view
        return np.ctypeslib.as_array(buf).reshape(self._h_px, self._w, 4)

    def synchronize(self):
        _lib.check(cuda.gdpt_device_synchronize(self.device), self.device, "synchronize")

    def output_image(self):
        """Zero-copy view of the page-locked buffer render() reads the frame back into (overwritten by the
        next render(); copy it to keep it)."""
        return self._pinned_view(host.gdpt_camera_output_image(self._h))

    # introspection for tests / bench ----------------------------------------------------------
    @property
    def device(self):
        return host.gdpt_camera_device(self._h)

    @property
    def main_shader(self):
        return host.gdpt_camera_main_shader(self._h)

    @property
    def progressive_shader(self):
        return host.gdpt_camera_progressive_shader(self._h)

    @property
    def temporal_shader(self):
        return host.gdpt_camera_temporal_shader(self._h)

    def temporal_params(self):
        """The 88 B Params block the last temporal-reprojection dispatch ran with (None before the first one)."""
        p = _lib.TemporalParams()
        return p if host.gdpt_camera_get_temporal_params(self._h, ctypes.byref(p)) == 1 else None

    def camera_block(self):
        c = _lib.Camera()
        host.gdpt_camera_get_camera_block(self._h, ctypes.byref(c))
        return c

    def last_frame_count(self):
        return host.gdpt_camera_last_frame_count(self._h)

    def stats(self):
        st = _lib.FrameStats()
        _lib.check(cuda.gdpt_shader_get_stats(self.main_shader, ctypes.byref(st)), self.device, "get_stats")
        return {k: getattr(st, k) for k, _ in _lib.FrameStats._fields_}

    def read_image(self, which="output"):
        rid = self._rid(which)
        n = self._w * self._h_px
        if which == "output":
            out = np.empty((self._h_px, self._w, 4), np.uint8)
        elif which == "depth":
            out = np.empty((self._h_px, self._w), np.float32)
        else:
            out = np.empty((self._h_px, self._w, 4), np.float32)
        shader = {"accum": self.progressive_shader, "history1": self.temporal_shader,
                  "history2": self.temporal_shader}.get(which, self.main_shader)
        _lib.check(cuda.gdpt_shader_get_image_uniform_buffer(shader, rid, 0, _ptr(out), out.nbytes), self.device,
                   "get_image_uniform_buffer")
        return out

    def read_trace(self, segment=0):
        n = self._w * self._h_px
        out = np.empty(n, dtype=_lib.TRACE_DTYPE)
        _lib.check(cuda.gdpt_shader_read_trace(self.main_shader, segment, _ptr(out), n), self.device, "read_trace")
        return out

    def read_visits(self):
        n = self._w * self._h_px
        out = np.empty((n, self._visits), dtype=np.uint32)
        _lib.check(cuda.gdpt_shader_read_visits(self.main_shader, _ptr(out), self._visits, out.size), self.device,
                   "read_visits")
        return out

    def _rid(self, which):
        """output / depth: the main shader's images; accum: progressive accumulation; history1 / history2: the
        temporal-reprojection ping-pong frame buffers."""
        if which in ("history1", "history2"):
            return host.gdpt_camera_temporal_rid(self._h, 0 if which == "history1" else 1)
        return {"output": host.gdpt_camera_output_rid, "depth": host.gdpt_camera_depth_rid,
                "accum": host.gdpt_camera_accum_rid}[which](self._h)

    def device_pointer(self, which="output"):
        rid = self._rid(which)
        p, s = ctypes.c_uint64(), ctypes.c_uint64()
        _lib.check(cuda.gdpt_rid_device_pointer(self.device, rid, ctypes.byref(p), ctypes.byref(s)), self.device,
                   "rid_device_pointer")
        return p.value, s.value


def _camera_ipc_methods():
    def export_output_handle(self):
        """CUDA IPC handle (64 bytes) of the RGBA8 output image, for the other per-GPU processes of a row-band frame."""
        buf = ctypes.create_string_buffer(64)
        _lib.check(cuda.gdpt_rid_ipc_export(self.device, self._rid("output"), buf), self.device, "rid_ipc_export")
        return buf.raw

    def open_peer_image(self, handle):
        p = ctypes.c_uint64()
        _lib.check(cuda.gdpt_device_ipc_open(self.device, ctypes.create_string_buffer(bytes(handle), 64), ctypes.byref(p)),
                   self.device, "device_ipc_open")
        return p.value

    def close_peer_image(self, ptr):
        _lib.check(cuda.gdpt_device_ipc_close(self.device, int(ptr)), self.device, "device_ipc_close")

    def set_peer_screens(self, ptrs):
        """Device addresses of the other GPUs' output images: the accumulate/tone-map kernel mirrors the rows this GPU
        owns into them (include/gdpt.h gdpt_shader_set_peer_screens).  Empty list = off."""
        host.gdpt_camera_prepare_post(self._h)  # the progressive shader is created lazily, as upstream
        if not self.progressive_shader:
            raise _lib.GdptError("set_peer_screens needs denoising_mode PROGRESSIVE_RENDERING and an initialised camera")
        arr = (ctypes.c_uint64 * max(len(ptrs), 1))(*[int(p) for p in ptrs])
        _lib.check(cuda.gdpt_shader_set_peer_screens(self.progressive_shader, arr, len(ptrs)), self.device, "set_peer_screens")

    def stream(self):
        return cuda.gdpt_device_stream(self.device)

    for f in (export_output_handle, open_peer_image, close_peer_image, set_peer_screens, stream):
        setattr(PathTracingCamera, f.__name__, f)


_camera_ipc_methods()


def make_camera_block(transform12, fov, width, height, frame_index):
    c = _lib.Camera()
    t = _f32(transform12).reshape(12)
    host.gdpt_make_camera_block(_ptr(t), float(fov), int(width), int(height), int(frame_index), ctypes.byref(c))
    return c


def make_temporal_delta(previous_vp16, transform12, fov, width, height):
    """(vp, delta) of one TemporalReprojection::render parameter update (temporal_reprojection.cpp:57-61)."""
    prev = _f32(previous_vp16).reshape(16)
    t = _f32(transform12).reshape(12)
    vp, delta = np.zeros(16, np.float32), np.zeros(16, np.float32)
    host.gdpt_make_temporal_delta(_ptr(prev), _ptr(t), float(fov), int(width), int(height), _ptr(vp), _ptr(delta))
    return vp, delta
