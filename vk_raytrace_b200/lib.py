"""Loader for libvkrt_b200.so.  Fails loudly when the CUDA library has not been built:
there is no CPU or PyTorch fallback for the product path."""
import ctypes as C
import os

from . import structs as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LibraryMissing(RuntimeError):
    pass


def library_path():
    # VKRT_LIB_PATH: A/B builds during bring-up; the default is the in-tree library
    return os.environ.get("VKRT_LIB_PATH") or os.path.join(_HERE, "libvkrt_b200.so")


# every symbol include/vkrt_b200.h declares
EXPORTS = [
    "vkrt_create", "vkrt_destroy", "vkrt_name", "vkrt_last_error", "vkrt_scene_begin", "vkrt_scene_add_primmesh",
    "vkrt_scene_add_node", "vkrt_scene_set_materials", "vkrt_scene_set_lights", "vkrt_scene_add_texture", "vkrt_scene_commit",
    "vkrt_build_env_accel", "vkrt_set_environment", "vkrt_set_sunsky", "vkrt_set_camera", "vkrt_output_create",
    "vkrt_output_device_ptr", "vkrt_output_clear", "vkrt_output_read", "vkrt_tonemap", "vkrt_tonemap_float", "vkrt_render",
    "vkrt_sync", "vkrt_pack_vertices", "vkrt_make_camera", "vkrt_light_from_matrix", "vkrt_trace_rays", "vkrt_triangle_ids",
    "vkrt_get_stats", "vkrt_reset_stats", "vkrt_mark", "vkrt_elapsed_ms",
    "vkrt_tonemap_async", "vkrt_tonemap_wait", "vkrt_host_alloc", "vkrt_host_free",
    "vkrt_stream", "vkrt_output_shard_size", "vkrt_output_pack_shard", "vkrt_output_unpack_shard",
    "vkrt_output_export", "vkrt_peer_open", "vkrt_peer_close", "vkrt_output_set_mirror", "vkrt_output_fence",
]


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, u32, i32, f32p = C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(C.c_float)
    lib.vkrt_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.vkrt_destroy.argtypes = [vp]
    lib.vkrt_name.restype = C.c_char_p
    lib.vkrt_last_error.argtypes = [vp]
    lib.vkrt_last_error.restype = C.c_char_p
    lib.vkrt_scene_begin.argtypes = [vp]
    lib.vkrt_scene_add_primmesh.argtypes = [vp, vp, u32, vp, u32, i32]
    lib.vkrt_scene_add_node.argtypes = [vp, vp, u32]
    lib.vkrt_scene_set_materials.argtypes = [vp, vp, u32]
    lib.vkrt_scene_set_lights.argtypes = [vp, vp, u32]
    lib.vkrt_scene_add_texture.argtypes = [vp, vp, u32, u32, i32, i32, i32, i32]
    lib.vkrt_scene_commit.argtypes = [vp, i32]
    lib.vkrt_build_env_accel.argtypes = [vp, u32, u32, vp, f32p, f32p]
    lib.vkrt_set_environment.argtypes = [vp, vp, u32, u32, vp]
    lib.vkrt_set_sunsky.argtypes = [vp, C.POINTER(S.SunAndSky)]
    lib.vkrt_set_camera.argtypes = [vp, C.POINTER(S.SceneCamera)]
    lib.vkrt_output_create.argtypes = [vp, u32, u32, vp]
    lib.vkrt_output_device_ptr.argtypes = [vp, C.POINTER(vp)]
    lib.vkrt_output_clear.argtypes = [vp]
    lib.vkrt_output_read.argtypes = [vp, vp]
    lib.vkrt_tonemap.argtypes = [vp, C.POINTER(S.Tonemapper), vp]
    lib.vkrt_tonemap_float.argtypes = [vp, C.POINTER(S.Tonemapper), vp]
    lib.vkrt_render.argtypes = [vp, C.POINTER(S.RtxState), i32, C.POINTER(S.RenderOpts)]
    lib.vkrt_sync.argtypes = [vp]
    lib.vkrt_pack_vertices.argtypes = [vp, vp, vp, vp, vp, u32, vp]
    lib.vkrt_make_camera.argtypes = [vp, vp, vp, C.c_float, C.c_float, C.c_float, i32, C.POINTER(S.SceneCamera)]
    lib.vkrt_light_from_matrix.argtypes = [vp, vp]
    lib.vkrt_trace_rays.argtypes = [vp, vp, u32, i32, vp]
    lib.vkrt_triangle_ids.argtypes = [vp, vp, u32, vp]
    lib.vkrt_get_stats.argtypes = [vp, C.POINTER(S.Stats)]
    lib.vkrt_reset_stats.argtypes = [vp]
    lib.vkrt_mark.argtypes = [vp, i32]
    lib.vkrt_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.vkrt_stream.argtypes = [vp, C.POINTER(vp)]
    lib.vkrt_tonemap_async.argtypes = [vp, C.POINTER(S.Tonemapper), vp, i32]
    lib.vkrt_tonemap_wait.argtypes = [vp, i32]
    lib.vkrt_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.vkrt_host_free.argtypes = [vp]
    lib.vkrt_output_shard_size.argtypes = [vp, i32, i32, i32, C.POINTER(u32)]
    lib.vkrt_output_pack_shard.argtypes = [vp, i32, i32, i32, vp]
    lib.vkrt_output_unpack_shard.argtypes = [vp, i32, i32, i32, vp]
    lib.vkrt_output_export.argtypes = [vp, vp]
    lib.vkrt_peer_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.vkrt_peer_close.argtypes = [vp, vp]
    lib.vkrt_output_set_mirror.argtypes = [vp, vp]
    lib.vkrt_output_fence.argtypes = [vp]
    _LIB = lib
    return lib


# ---- libvkrt_host.so (include/vkrt_host.h): file ingest in front of the C ABI ----------------------------------------------
_HOST = None
HOST_EXPORTS = ["vkrt_host_load_gltf", "vkrt_host_load_environment", "vkrt_host_last_error"]


class HostSceneInfo(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("center", C.c_float * 3), ("up", C.c_float * 3), ("yfov_degrees", C.c_float), ("camera_from_file", C.c_int32),
                ("prim_meshes", C.c_uint32), ("nodes", C.c_uint32), ("materials", C.c_uint32), ("textures", C.c_uint32), ("lights", C.c_uint32),
                ("triangles", C.c_uint64), ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3)]


def load_host_library():
    """libvkrt_host.so next to libvkrt_b200.so (built by __graft_entry__.build()); loads the renderer library first so both share it."""
    global _HOST
    if _HOST is not None:
        return _HOST
    load_library()
    path = os.path.join(_HERE, "libvkrt_host.so")
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.vkrt_host_load_gltf.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.POINTER(HostSceneInfo)]
    lib.vkrt_host_load_environment.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.vkrt_host_last_error.restype = C.c_char_p
    _HOST = lib
    return lib
