"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so with the same call surface as
vk_raytrace_b200.Context, so that `upload_scene` can feed both from one SceneDesc.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    if os.environ.get("ORACLE_LIB_PATH"):        # tests: an instrumented build of the same sources (UBSan)
        return os.environ["ORACLE_LIB_PATH"]
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def load():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
        L = _LIB
        L.oracle_create.argtypes = [C.POINTER(vp)]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_set_threads.argtypes = [vp, C.c_int]
        L.oracle_scene_begin.argtypes = [vp]
        L.oracle_scene_add_primmesh.argtypes = [vp, vp, u32, vp, u32, i32]
        L.oracle_scene_add_node.argtypes = [vp, vp, u32]
        L.oracle_scene_set_materials.argtypes = [vp, vp, u32]
        L.oracle_scene_set_lights.argtypes = [vp, vp, u32]
        L.oracle_scene_add_texture.argtypes = [vp, vp, u32, u32, i32, i32, i32, i32]
        L.oracle_scene_commit.argtypes = [vp, i32]
        L.oracle_build_env_accel.argtypes = [vp, u32, u32, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_set_environment.argtypes = [vp, vp, u32, u32, vp]
        L.oracle_set_sunsky.argtypes = [vp, vp]
        L.oracle_set_camera.argtypes = [vp, vp]
        L.oracle_output_create.argtypes = [vp, u32, u32]
        L.oracle_output_read.argtypes = [vp, vp]
        L.oracle_output_write.argtypes = [vp, vp]
        L.oracle_render.argtypes = [vp, vp, i32, i32, i32, i32]
        L.oracle_render_tiles.argtypes = [vp, vp, i32, i32, i32, i32, i32]
        L.oracle_tonemap.argtypes = [vp, vp, vp, vp]
        L.oracle_trace_rays.argtypes = [vp, vp, u32, i32, vp]
        L.oracle_triangle_ids.argtypes = [vp, vp, u32, vp]
        L.oracle_get_counters.argtypes = [vp, vp]
        L.oracle_reset_counters.argtypes = [vp]
        L.oracle_num_triangles.argtypes = [vp]
        L.oracle_num_triangles.restype = u32
        L.oracle_tea.argtypes = [u32, u32]
        L.oracle_tea.restype = u32
        L.oracle_pcg.argtypes = [C.POINTER(u32)]
        L.oracle_pcg.restype = u32
        L.oracle_rand.argtypes = [C.POINTER(u32)]
        L.oracle_rand.restype = C.c_float
        L.oracle_pcg3d.argtypes = [vp]
        L.oracle_compress_unit_vec.argtypes = [C.c_float] * 3
        L.oracle_compress_unit_vec.restype = u32
        L.oracle_decompress_unit_vec.argtypes = [u32, vp]
        L.oracle_pack_unorm4x8.argtypes = [C.c_float] * 4
        L.oracle_pack_unorm4x8.restype = u32
        L.oracle_offset_ray.argtypes = [vp, vp, vp]
        L.oracle_alpha_draw.argtypes = [u32, u32]
        L.oracle_alpha_draw.restype = C.c_float
        L.oracle_pack_vertices.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.oracle_bsdf_eval.argtypes = [i32, vp, vp]
        L.oracle_bsdf_sample.argtypes = [i32, vp, u32, vp]
        L.oracle_sun_and_sky.argtypes = [vp, vp, vp]
        L.oracle_make_camera.argtypes = [vp, vp, vp, C.c_float, C.c_float, C.c_float, i32, vp]
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_vertices(m):
    from vk_raytrace_b200.structs import VertexAttributes
    n = len(m.positions)
    out = np.zeros(n, dtype=VertexAttributes)
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (m.positions, m.normals, m.tangents, m.texcoords, m.colors)]
    load().oracle_pack_vertices(*[_ptr(a) for a in arrs], n, _ptr(out))
    return out


def make_camera(eye, center, up, fov, aspect, aperture=0.0, nb_lights=0):
    """Scene::updateCamera (src/scene.cpp:629-640) restated on the oracle side (glm lookAt / perspectiveRH_ZO / inverse in float)."""
    from vk_raytrace_b200.structs import SceneCamera
    cam = SceneCamera()
    e, c, u = (np.ascontiguousarray(v, dtype=np.float32) for v in (eye, center, up))
    load().oracle_make_camera(_ptr(e), _ptr(c), _ptr(u), fov, aspect, aperture, nb_lights, C.byref(cam))
    return cam


def build_env_accel(env):
    from vk_raytrace_b200.structs import EnvAccel
    env = np.ascontiguousarray(env, dtype=np.float32)
    h, w = env.shape[:2]
    out = np.zeros(h * w, dtype=EnvAccel)
    integ, avg = C.c_float(), C.c_float()
    load().oracle_build_env_accel(_ptr(env), w, h, _ptr(out), C.byref(integ), C.byref(avg))
    return out, integ.value, avg.value


class Oracle:
    """CPU restatement with the Context call surface."""

    def __init__(self, threads=0, use_bvh=True):
        self.lib = load()
        self.h = C.c_void_p()
        self.lib.oracle_create(C.byref(self.h))
        self.lib.oracle_set_threads(self.h, threads)
        self.use_bvh = use_bvh
        self.width = self.height = 0

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"oracle {what}: error {rc}")

    def scene_begin(self):
        self._ck(self.lib.oracle_scene_begin(self.h), "scene_begin")

    def add_primmesh(self, vertices, indices, material):
        vertices = np.ascontiguousarray(vertices)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        self._ck(self.lib.oracle_scene_add_primmesh(self.h, _ptr(vertices), len(vertices), _ptr(indices), len(indices), material), "add_primmesh")

    def add_node(self, matrix16, mesh):
        m = np.ascontiguousarray(matrix16, dtype=np.float32)
        self._ck(self.lib.oracle_scene_add_node(self.h, _ptr(m), mesh), "add_node")

    def set_materials(self, mats):
        mats = np.ascontiguousarray(mats)
        self._ck(self.lib.oracle_scene_set_materials(self.h, _ptr(mats), len(mats)), "set_materials")

    def set_lights(self, lights):
        n = 0 if lights is None else len(lights)
        self._ck(self.lib.oracle_scene_set_lights(self.h, _ptr(np.ascontiguousarray(lights)) if n else None, n), "set_lights")

    def add_texture(self, t):
        rgba = np.ascontiguousarray(t.rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        self._ck(self.lib.oracle_scene_add_texture(self.h, _ptr(rgba), w, h, t.mag_filter, t.min_filter, t.wrap_s, t.wrap_t), "add_texture")

    def commit(self, build_mode=0):
        self._ck(self.lib.oracle_scene_commit(self.h, 1 if self.use_bvh else 0), "commit")

    def set_environment(self, env, accel=None):
        env = np.ascontiguousarray(env, dtype=np.float32)
        h, w = env.shape[:2]
        self._ck(self.lib.oracle_set_environment(self.h, _ptr(env), w, h, None if accel is None else _ptr(np.ascontiguousarray(accel))), "set_environment")

    def set_sunsky(self, ss):
        self._ck(self.lib.oracle_set_sunsky(self.h, C.byref(ss)), "set_sunsky")

    def set_camera(self, cam):
        self._ck(self.lib.oracle_set_camera(self.h, C.byref(cam)), "set_camera")

    def output_create(self, width, height, device_ptr=None):
        self.width, self.height = width, height
        self._ck(self.lib.oracle_output_create(self.h, width, height), "output_create")

    def output_read(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.oracle_output_read(self.h, _ptr(out)), "output_read")
        return out

    def output_write(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self._ck(self.lib.oracle_output_write(self.h, _ptr(img)), "output_write")

    def render(self, state, num_frames=1, seed_mode=0, rows=(0, 0)):
        self._ck(self.lib.oracle_render(self.h, C.byref(state), num_frames, seed_mode, rows[0], rows[1]), "render")

    def render_tiles(self, state, num_frames=1, seed_mode=0, tile=64, shard_rank=0, shard_count=1):
        self._ck(self.lib.oracle_render_tiles(self.h, C.byref(state), num_frames, seed_mode, tile, shard_rank, shard_count), "render_tiles")

    def tonemap(self, tm):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._ck(self.lib.oracle_tonemap(self.h, C.byref(tm), None, _ptr(out)), "tonemap")
        return out

    def tonemap_float(self, tm):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.oracle_tonemap(self.h, C.byref(tm), _ptr(out), None), "tonemap")
        return out

    def trace_rays(self, rays, closest=True):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.empty((len(rays), 4), dtype=np.float32)
        self._ck(self.lib.oracle_trace_rays(self.h, _ptr(rays), len(rays), 1 if closest else 0, _ptr(out)), "trace_rays")
        return out

    def triangle_ids(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty((len(ids), 2), dtype=np.uint32)
        self._ck(self.lib.oracle_triangle_ids(self.h, _ptr(ids), len(ids), _ptr(out)), "triangle_ids")
        return out

    def counters(self):
        out = np.zeros(4, dtype=np.uint64)
        self.lib.oracle_get_counters(self.h, _ptr(out))
        return dict(closest_rays=int(out[0]), shadow_rays=int(out[1]), nodes_visited=int(out[2]), tris_tested=int(out[3]))

    def reset_counters(self):
        self.lib.oracle_reset_counters(self.h)
