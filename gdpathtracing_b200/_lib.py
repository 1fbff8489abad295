"""ctypes bindings of the two shared libraries of the backend.

``libgdpt_cuda.so``  -- the C-ABI drop-in boundary (include/gdpt.h): CUDA kernels + engine.
``libgdpt_host.so``  -- the C++ host layer above it (include/gdpt_host.h): GeometryGroup3D /
                        PathTracingCamera twins and the BLAS/TLAS builder.

There is no Python or CPU implementation of the render path in this package: if the CUDA
library is missing the import fails loudly, and without a GPU ``gdpt_device_create`` fails.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_uint8, c_uint32,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(_HERE, "libgdpt_cuda.so")
HOST_LIB_PATH = os.path.join(_HERE, "libgdpt_host.so")


class BackendMissing(ImportError):
    pass


def _load(path):
    if not os.path.exists(path):
        raise BackendMissing(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no fallback implementation.")
    return ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)


# ----------------------------------------------------------------------------- wire structs
class RenderParams(Structure):
    _fields_ = [("background", c_float * 4), ("width", c_int32), ("height", c_int32), ("fov", c_float),
                ("triangle_count", c_uint32), ("blas_count", c_uint32)]


class Camera(Structure):
    _fields_ = [("vp", c_float * 16), ("ivp", c_float * 16), ("position", c_float * 4), ("frame_index", c_uint32),
                ("z_near", c_float), ("z_far", c_float), ("_tail_pad", c_uint32)]


class ProgressiveParams(Structure):
    _fields_ = [("width", c_int32), ("height", c_int32), ("frame_count", c_uint32)]


class TemporalParams(Structure):
    _fields_ = [("delta_matrix", c_float * 16), ("width", c_int32), ("height", c_int32), ("frame_count", c_uint32),
                ("blend_factor", c_float), ("near_plane", c_float), ("far_plane", c_float)]


class StandardMaterial(Structure):
    _fields_ = [("albedo", c_float * 3), ("metallic", c_float), ("roughness", c_float), ("emission", c_float * 3),
                ("emission_energy_multiplier", c_float), ("albedo_texture", c_int32), ("is_standard", c_int32)]


class FrameStats(Structure):
    _fields_ = [("rays", c_uint64), ("primary_hits", c_uint64), ("node_pops", c_uint64), ("box_tests", c_uint64),
                ("tri_tests", c_uint64), ("tlas_leaves", c_uint64), ("kernel_launches", c_uint32),
                ("max_stack", c_uint32), ("k1_ms", c_float), ("k2_ms", c_float), ("retraced", c_uint64),
                ("own_node_steps", c_uint64), ("own_box_tests", c_uint64), ("own_tri_tests", c_uint64),
                ("own_inst_entries", c_uint64), ("own_proofs", c_uint64)]


assert ctypes.sizeof(RenderParams) == 36 and ctypes.sizeof(Camera) == 160 and ctypes.sizeof(ProgressiveParams) == 12
assert ctypes.sizeof(TemporalParams) == 88

# numpy dtype of gdpt_trace_record (include/gdpt_wire.h)
TRACE_DTYPE = [("hit", "<u4"), ("triangle", "<u4"), ("blas", "<u4"), ("front", "<u4"), ("t", "<f4"), ("u", "<f4"),
               ("v", "<f4"), ("node_pops", "<u4"), ("box_tests", "<u4"), ("tri_tests", "<u4"), ("tlas_leaves", "<u4"),
               ("max_stack", "<u4"), ("visit_hash_lo", "<u4"), ("visit_hash_hi", "<u4")]

FORMAT_RGBA8, FORMAT_R32F, FORMAT_RGBA32F = 1, 2, 3
UNIFORM_IMAGE, UNIFORM_STORAGE_BUFFER = 3, 8
DENOISE_PROGRESSIVE, DENOISE_TEMPORAL, DENOISE_NONE = 0, 1, 2

# every symbol include/gdpt.h declares: (name, restype, argtypes)
CUDA_API = [
    ("gdpt_device_create", c_int, [c_int, POINTER(c_void_p)]),
    ("gdpt_device_destroy", None, [c_void_p]),
    ("gdpt_last_error", c_char_p, [c_void_p]),
    ("gdpt_abi_version", c_uint32, []),
    ("gdpt_shader_get_schedule", c_int, [c_void_p]),
    ("gdpt_progressive_accumulate", c_int, [c_void_p, c_uint64, c_uint64, c_uint64, c_int, c_int, c_uint32]),
    ("gdpt_shader_create", c_int, [c_void_p, c_char_p, POINTER(c_char_p), c_int, POINTER(c_void_p)]),
    ("gdpt_shader_destroy", None, [c_void_p]),
    ("gdpt_shader_create_storage_buffer_uniform", c_uint64, [c_void_p, c_void_p, c_uint64, c_int, c_int]),
    ("gdpt_shader_update_storage_buffer_uniform", c_int, [c_void_p, c_uint64, c_void_p, c_uint64]),
    ("gdpt_shader_get_storage_buffer_uniform", c_int, [c_void_p, c_uint64, c_void_p, c_uint64]),
    ("gdpt_shader_create_image_uniform", c_uint64, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    ("gdpt_shader_create_layered_image_uniform", c_uint64,
     [c_void_p, POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]),
    ("gdpt_shader_get_image_uniform_buffer", c_int, [c_void_p, c_uint64, c_int, c_void_p, c_uint64]),
    ("gdpt_shader_add_existing_buffer", c_int, [c_void_p, c_uint64, c_int, c_int, c_int]),
    ("gdpt_shader_finish_create_uniforms", c_int, [c_void_p]),
    ("gdpt_shader_check_ready", c_int, [c_void_p]),
    ("gdpt_shader_compute", c_int, [c_void_p, c_int, c_int, c_int]),
    ("gdpt_render_frame", c_int, [c_void_p, c_void_p, POINTER(Camera), c_int, c_uint32, c_void_p, c_void_p]),
    ("gdpt_render_frame_async", c_int, [c_void_p, c_void_p, POINTER(Camera), c_int, c_uint32]),
    ("gdpt_device_synchronize", c_int, [c_void_p]),
    ("gdpt_render_frame_begin", c_int, [c_void_p, c_void_p, POINTER(Camera), c_int, c_uint32, c_void_p, c_void_p]),
    ("gdpt_render_frame_wait", c_int, [c_void_p, POINTER(FrameStats)]),
    ("gdpt_shader_stage_params", c_int, [c_void_p, c_void_p, c_uint64]),
    ("gdpt_shader_set_shard", c_int, [c_void_p, c_int, c_int, c_int]),
    ("gdpt_rid_device_pointer", c_int, [c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)]),
    ("gdpt_host_alloc", c_void_p, [c_uint64]),
    ("gdpt_host_free", None, [c_void_p]),
    ("gdpt_device_stream", c_uint64, [c_void_p]),
    ("gdpt_device_create_buffer", c_uint64, [c_void_p, c_uint64]),
    ("gdpt_device_free_buffer", c_int, [c_void_p, c_uint64]),
    ("gdpt_rid_ipc_export", c_int, [c_void_p, c_uint64, c_void_p]),
    ("gdpt_device_ipc_open", c_int, [c_void_p, c_void_p, POINTER(c_uint64)]),
    ("gdpt_device_ipc_close", c_int, [c_void_p, c_uint64]),
    ("gdpt_shader_set_peer_screens", c_int, [c_void_p, POINTER(c_uint64), c_int]),
    ("gdpt_shader_get_stats", c_int, [c_void_p, POINTER(FrameStats)]),
    ("gdpt_shader_set_stage_timing", c_int, [c_void_p, c_int]),
    ("gdpt_shader_get_stage_times", c_int, [c_void_p, c_void_p, c_int]),
    ("gdpt_shader_set_warp_profile", c_int, [c_void_p, c_int]),
    ("gdpt_shader_read_warp_profile", ctypes.c_int64, [c_void_p, c_void_p, c_uint64]),
    ("gdpt_shader_read_trace", c_int, [c_void_p, c_int, c_void_p, c_uint64]),
    ("gdpt_shader_read_visits", c_int, [c_void_p, c_void_p, c_uint32, c_uint64]),
]

HOST_API = [
    ("gdpt_group_create", c_void_p, []),
    ("gdpt_group_destroy", None, [c_void_p]),
    ("gdpt_group_add_texture", c_int, [c_void_p, c_void_p, c_int, c_int]),
    ("gdpt_group_add_material", c_int, [c_void_p, POINTER(StandardMaterial)]),
    ("gdpt_group_add_material_ext", c_int, [c_void_p, POINTER(StandardMaterial), c_int, c_int, c_int]),
    ("gdpt_group_set_material_ext", None, [c_void_p, c_int]),
    ("gdpt_group_get_material_ext", c_int, [c_void_p]),
    ("gdpt_group_add_mesh", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("gdpt_group_add_mesh_instance", None, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int]),
    ("gdpt_group_set_default_material", None, [c_void_p, c_int]),
    ("gdpt_group_set_texture_array_resolution", None, [c_void_p, c_int]),
    ("gdpt_group_get_texture_array_resolution", c_int, [c_void_p]),
    ("gdpt_group_set_build_threads", None, [c_void_p, c_int]),
    ("gdpt_group_get_build_threads", c_int, [c_void_p]),
    ("gdpt_group_build", None, [c_void_p]),
    ("gdpt_group_last_build_seconds", c_double, [c_void_p]),
    ("gdpt_group_buffer_size", c_uint64, [c_void_p, c_int]),
    ("gdpt_group_buffer_data", c_void_p, [c_void_p, c_int]),
    ("gdpt_group_texture_layer_count", c_int, [c_void_p]),
    ("gdpt_group_texture_layer", c_void_p, [c_void_p, c_int]),
    ("gdpt_camera_create", c_void_p, []),
    ("gdpt_camera_destroy", None, [c_void_p]),
    ("gdpt_camera_set_fov", None, [c_void_p, c_float]),
    ("gdpt_camera_get_fov", c_float, [c_void_p]),
    ("gdpt_camera_set_geometry_group", None, [c_void_p, c_void_p]),
    ("gdpt_camera_set_denoising_mode", None, [c_void_p, c_int]),
    ("gdpt_camera_get_denoising_mode", c_int, [c_void_p]),
    ("gdpt_camera_set_window_size", None, [c_void_p, c_int, c_int]),
    ("gdpt_camera_set_global_transform", None, [c_void_p, c_void_p]),
    ("gdpt_camera_set_max_depth", None, [c_void_p, c_int]),
    ("gdpt_camera_set_cuda_device", None, [c_void_p, c_int]),
    ("gdpt_camera_set_frame_index", None, [c_void_p, c_uint32]),
    ("gdpt_camera_set_shard", None, [c_void_p, c_int, c_int, c_int]),
    ("gdpt_camera_set_trace", None, [c_void_p, c_int, c_uint32]),
    ("gdpt_camera_set_debug_steps", None, [c_void_p, c_int]),
    ("gdpt_camera_set_cull", None, [c_void_p, c_int]),
    ("gdpt_camera_set_record_hits", None, [c_void_p, c_int]),
    ("gdpt_camera_set_variant", None, [c_void_p, c_int]),
    ("gdpt_camera_set_tuning", None, [c_void_p, c_char_p, c_int]),
    ("gdpt_camera_set_count_work", None, [c_void_p, c_int]),
    ("gdpt_camera_set_fused_frame", None, [c_void_p, c_int]),
    ("gdpt_camera_init", c_int, [c_void_p]),
    ("gdpt_camera_render", None, [c_void_p]),
    ("gdpt_camera_render_device_only", None, [c_void_p]),
    ("gdpt_camera_render_begin", c_int, [c_void_p]),
    ("gdpt_camera_render_wait", c_void_p, [c_void_p, POINTER(FrameStats)]),
    ("gdpt_camera_output_image", c_void_p, [c_void_p]),
    ("gdpt_camera_main_shader", c_void_p, [c_void_p]),
    ("gdpt_camera_progressive_shader", c_void_p, [c_void_p]),
    ("gdpt_camera_prepare_post", None, [c_void_p]),
    ("gdpt_camera_temporal_shader", c_void_p, [c_void_p]),
    ("gdpt_camera_temporal_rid", c_uint64, [c_void_p, c_int]),
    ("gdpt_camera_get_temporal_params", c_int, [c_void_p, POINTER(TemporalParams)]),
    ("gdpt_camera_device", c_void_p, [c_void_p]),
    ("gdpt_camera_output_rid", c_uint64, [c_void_p]),
    ("gdpt_camera_depth_rid", c_uint64, [c_void_p]),
    ("gdpt_camera_accum_rid", c_uint64, [c_void_p]),
    ("gdpt_camera_get_camera_block", None, [c_void_p, POINTER(Camera)]),
    ("gdpt_camera_last_frame_count", c_uint32, [c_void_p]),
    ("gdpt_make_camera_block", None, [c_void_p, c_float, c_int, c_int, c_uint32, POINTER(Camera)]),
    ("gdpt_make_temporal_delta", None, [c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p]),
]


def _bind(lib, table):
    for name, restype, argtypes in table:
        fn = getattr(lib, name)  # AttributeError here = the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


cuda = _bind(_load(CUDA_LIB_PATH), CUDA_API)
host = _bind(_load(HOST_LIB_PATH), HOST_API)


class GdptError(RuntimeError):
    pass


def check(rc, device=None, what=""):
    if rc != 0:
        msg = cuda.gdpt_last_error(device)
        raise GdptError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
