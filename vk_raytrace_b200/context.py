"""Python harness over the C ABI (include/vkrt_b200.h): `Context` is what tests and bench.py
drive; `SceneDesc` is the backend-neutral scene description that both the CUDA library
and the CPU oracle (tests only) are fed from, through the same call sequence."""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import structs as S
from .lib import load_library


@dataclass
class MeshDesc:
    positions: np.ndarray          # (n,3) f32
    normals: np.ndarray            # (n,3) f32
    tangents: np.ndarray           # (n,4) f32, w = handedness
    texcoords: np.ndarray          # (n,2) f32
    colors: np.ndarray             # (n,4) f32
    indices: np.ndarray            # (m,) u32, triangle triples
    material: int = 0


@dataclass
class TextureDesc:
    rgba: np.ndarray               # (h,w,4) u8
    mag_filter: int = 9729
    min_filter: int = 9729
    wrap_s: int = 10497
    wrap_t: int = 10497


@dataclass
class SceneDesc:
    meshes: List[MeshDesc] = field(default_factory=list)
    nodes: list = field(default_factory=list)            # [(4x4 math-convention float32 matrix, mesh index)]
    materials: Optional[np.ndarray] = None               # structs.GltfShadeMaterial
    lights: Optional[np.ndarray] = None                  # structs.Light
    textures: List[TextureDesc] = field(default_factory=list)
    env: Optional[np.ndarray] = None                     # (h,w,4) f32
    camera: dict = field(default_factory=dict)           # eye, center, up, fov (degrees), aperture
    name: str = "scene"

    def num_triangles(self):
        return sum(len(self.meshes[m].indices) // 3 for _, m in self.nodes)


def shard_pixel_list(width, height, tile, rank, count):
    """Pixels (y * width + x) of the tile shard (tile, rank, count) in the order the library stores them -- tiles dealt round-robin
    (t % count == rank), 8x4 pixel blocks inside a tile: the host-side mirror of vkrt_output_pack_shard / _unpack_shard, used by the
    CPU multi-process test and by hosts that post-process gathered shards themselves."""
    tile = tile if tile > 0 else 64
    tx, ty = (width + tile - 1) // tile, (height + tile - 1) // tile
    out = []
    for t in range(tx * ty):
        if count > 1 and t % count != rank:
            continue
        x0, y0 = (t % tx) * tile, (t // tx) * tile
        x1, y1 = min(x0 + tile, width), min(y0 + tile, height)
        for by in range(y0, y1, 4):
            for bx in range(x0, x1, 8):
                ys = np.arange(by, min(by + 4, y1))[:, None]
                xs = np.arange(bx, min(bx + 8, x1))[None, :]
                out.append((ys * width + xs).reshape(-1))
    return np.concatenate(out).astype(np.uint32) if out else np.zeros(0, np.uint32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def column_major(m):
    """4x4 math-convention matrix -> 16 floats in glm column-major order."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T).reshape(16)


def pack_vertices(m: MeshDesc):
    """vkrt_pack_vertices: Scene::createVertexBuffer packing (src/scene.cpp:224-246)."""
    n = len(m.positions)
    out = np.zeros(n, dtype=S.VertexAttributes)
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (m.positions, m.normals, m.tangents, m.texcoords, m.colors)]
    rc = load_library().vkrt_pack_vertices(*[_ptr(a) for a in arrs], n, _ptr(out))
    assert rc == 0
    return out


def make_camera(eye, center, up, fov, aspect, aperture=0.0, nb_lights=0):
    """vkrt_make_camera: Scene::updateCamera (src/scene.cpp:629-640)."""
    cam = S.SceneCamera()
    e, c, u = (np.ascontiguousarray(v, dtype=np.float32) for v in (eye, center, up))
    rc = load_library().vkrt_make_camera(_ptr(e), _ptr(c), _ptr(u), fov, aspect, aperture, nb_lights, C.byref(cam))
    assert rc == 0
    return cam


def build_env_accel(env):
    """vkrt_build_env_accel: HdrSampling::createEnvironmentAccel (src/hdr_sampling.cpp:187-248)."""
    env = np.ascontiguousarray(env, dtype=np.float32)
    h, w = env.shape[:2]
    out = np.zeros(h * w, dtype=S.EnvAccel)
    integ, avg = C.c_float(), C.c_float()
    rc = load_library().vkrt_build_env_accel(_ptr(env), w, h, _ptr(out), C.byref(integ), C.byref(avg))
    assert rc == 0
    return out, integ.value, avg.value


class Context:
    """One renderer context on one GPU (vkrt_create .. vkrt_destroy)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.vkrt_create(device, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"vkrt_create failed with code {rc} (is a CUDA GPU visible?)")
        self.width = self.height = 0

    def close(self):
        if self.h:
            self.lib.vkrt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: error {rc}: {self.lib.vkrt_last_error(self.h).decode()}")

    # ---- host helpers (pure functions of the library, no GPU needed) ----
    pack_vertices = staticmethod(lambda m: pack_vertices(m))
    make_camera = staticmethod(lambda *a, **k: make_camera(*a, **k))
    build_env_accel = staticmethod(lambda env: build_env_accel(env))

    # ---- scene ----
    def scene_begin(self):
        self._ck(self.lib.vkrt_scene_begin(self.h), "scene_begin")

    def add_primmesh(self, vertices, indices, material):
        vertices = np.ascontiguousarray(vertices)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        self._ck(self.lib.vkrt_scene_add_primmesh(self.h, _ptr(vertices), len(vertices), _ptr(indices), len(indices), material), "add_primmesh")

    def add_node(self, matrix16, mesh):
        m = np.ascontiguousarray(matrix16, dtype=np.float32)
        self._ck(self.lib.vkrt_scene_add_node(self.h, _ptr(m), mesh), "add_node")

    def set_materials(self, mats):
        mats = np.ascontiguousarray(mats)
        self._ck(self.lib.vkrt_scene_set_materials(self.h, _ptr(mats), len(mats)), "set_materials")

    def set_lights(self, lights):
        n = 0 if lights is None else len(lights)
        if n:
            lights = np.ascontiguousarray(lights)
            self._ck(self.lib.vkrt_scene_set_lights(self.h, _ptr(lights), n), "set_lights")
        else:
            self._ck(self.lib.vkrt_scene_set_lights(self.h, None, 0), "set_lights")

    def add_texture(self, t: TextureDesc):
        rgba = np.ascontiguousarray(t.rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        self._ck(self.lib.vkrt_scene_add_texture(self.h, _ptr(rgba), w, h, t.mag_filter, t.min_filter, t.wrap_s, t.wrap_t), "add_texture")

    def commit(self, build_mode=0):
        self._ck(self.lib.vkrt_scene_commit(self.h, build_mode), "scene_commit")

    def set_environment(self, env, accel=None):
        env = np.ascontiguousarray(env, dtype=np.float32)
        h, w = env.shape[:2]
        self._ck(self.lib.vkrt_set_environment(self.h, _ptr(env), w, h, None if accel is None else _ptr(np.ascontiguousarray(accel))), "set_environment")

    def set_sunsky(self, ss):
        self._ck(self.lib.vkrt_set_sunsky(self.h, C.byref(ss)), "set_sunsky")

    def set_camera(self, cam):
        self._ck(self.lib.vkrt_set_camera(self.h, C.byref(cam)), "set_camera")

    # ---- files (libvkrt_host.so, include/vkrt_host.h) ----
    def load_gltf(self, path, build_mode=0):
        """Scene::load from a .gltf / .glb file through the C++ ingest; returns lib.HostSceneInfo (camera, counts, bounds)."""
        from .lib import HostSceneInfo, load_host_library
        host = load_host_library()
        info = HostSceneInfo()
        rc = host.vkrt_host_load_gltf(self.h, os.fsencode(path), build_mode, C.byref(info))
        if rc != 0:
            raise RuntimeError(f"load_gltf: error {rc}: {host.vkrt_host_last_error().decode()}")
        return info

    def load_environment(self, path):
        """HdrSampling::loadEnvironment from a Radiance .hdr / .pfm file; returns (integral, width, height)."""
        from .lib import load_host_library
        host = load_host_library()
        integ, w, h = C.c_float(), C.c_uint32(), C.c_uint32()
        rc = host.vkrt_host_load_environment(self.h, os.fsencode(path), C.byref(integ), C.byref(w), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"load_environment: error {rc}: {host.vkrt_host_last_error().decode()}")
        return integ.value, w.value, h.value

    # ---- output ----
    def output_create(self, width, height, device_ptr=None):
        self.width, self.height = width, height
        self._ck(self.lib.vkrt_output_create(self.h, width, height, device_ptr), "output_create")

    def output_clear(self):
        self._ck(self.lib.vkrt_output_clear(self.h), "output_clear")

    def output_device_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.vkrt_output_device_ptr(self.h, C.byref(p)), "output_device_ptr")
        return p.value

    def output_read(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.vkrt_output_read(self.h, _ptr(out)), "output_read")
        return out

    def tonemap(self, tm, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._ck(self.lib.vkrt_tonemap(self.h, C.byref(tm), _ptr(out)), "tonemap")
        return out

    def tonemap_async(self, tm, out, slot):
        """vkrt_tonemap_async: tonemap + read-back into `out` (page-locked RGBA8 array) without waiting; `tonemap_wait(slot)` completes it."""
        self._ck(self.lib.vkrt_tonemap_async(self.h, C.byref(tm), _ptr(out), slot), "tonemap_async")

    def tonemap_wait(self, slot):
        self._ck(self.lib.vkrt_tonemap_wait(self.h, slot), "tonemap_wait")

    def host_alloc(self, shape, dtype=np.uint8):
        """Page-locked host array (vkrt_host_alloc); freed with host_free(array)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        if self.lib.vkrt_host_alloc(n, C.byref(p)) != 0:
            raise MemoryError("vkrt_host_alloc")
        buf = (C.c_uint8 * n).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[a.ctypes.data] = p.value
        return a

    def host_free(self, a):
        p = getattr(self, "_pinned", {}).pop(a.ctypes.data, None)
        if p is not None:
            self.lib.vkrt_host_free(C.c_void_p(p))

    # ---- tile shards ----
    def shard_size(self, tile, rank, count):
        n = C.c_uint32()
        self._ck(self.lib.vkrt_output_shard_size(self.h, tile, rank, count, C.byref(n)), "output_shard_size")
        return n.value

    def pack_shard(self, tile, rank, count, device_ptr):
        self._ck(self.lib.vkrt_output_pack_shard(self.h, tile, rank, count, device_ptr), "output_pack_shard")

    def unpack_shard(self, tile, rank, count, device_ptr):
        self._ck(self.lib.vkrt_output_unpack_shard(self.h, tile, rank, count, device_ptr), "output_unpack_shard")

    # ---- peer image (vkrt_output_export / vkrt_peer_open / vkrt_output_set_mirror / vkrt_output_fence) ----
    def output_export(self):
        """64-byte CUDA IPC handle of the library-owned accumulation image (what the root of a sharded render publishes)."""
        h = (C.c_uint8 * 64)()
        self._ck(self.lib.vkrt_output_export(self.h, h), "output_export")
        return bytes(h)

    def peer_open(self, handle):
        """Maps an image another process exported; returns the device pointer (int)."""
        h = (C.c_uint8 * 64).from_buffer_copy(bytes(handle))
        p = C.c_void_p()
        self._ck(self.lib.vkrt_peer_open(self.h, h, C.byref(p)), "peer_open")
        return p.value

    def peer_close(self, device_ptr):
        self._ck(self.lib.vkrt_peer_close(self.h, device_ptr), "peer_close")

    def set_mirror(self, device_ptr):
        """Every pixel accumulated from now on is also stored to the same pixel of the image at `device_ptr` (None: off)."""
        self._ck(self.lib.vkrt_output_set_mirror(self.h, device_ptr), "output_set_mirror")

    def output_fence(self):
        self._ck(self.lib.vkrt_output_fence(self.h), "output_fence")

    def tonemap_float(self, tm):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.vkrt_tonemap_float(self.h, C.byref(tm), _ptr(out)), "tonemap_float")
        return out

    # ---- render ----
    def render(self, state, num_frames=1, opts=None, sync=True, precise=None):
        """vkrt_render.  `precise` (or the context attribute `precise_default`) selects IEEE-exact shading arithmetic
        (vkrt_render_opts.precise_shading = 1); the library default is the fast shade kernel."""
        if precise is None:
            precise = getattr(self, "precise_default", False)
        if precise and (opts is None or not opts.precise_shading):
            o = S.RenderOpts()
            if opts is not None:
                C.memmove(C.byref(o), C.byref(opts), C.sizeof(o))
            o.precise_shading = 1
            opts = o
        self._ck(self.lib.vkrt_render(self.h, C.byref(state), num_frames, None if opts is None else C.byref(opts)), "render")
        if sync:
            self.sync()

    def stream(self):
        """cudaStream_t (integer handle) of the context's image-operation stream (vkrt_stream)."""
        p = C.c_void_p()
        self._ck(self.lib.vkrt_stream(self.h, C.byref(p)), "stream")
        return p.value or 0

    def sync(self):
        self._ck(self.lib.vkrt_sync(self.h), "sync")

    def trace_rays(self, rays, closest=True):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.empty((len(rays), 4), dtype=np.float32)
        self._ck(self.lib.vkrt_trace_rays(self.h, _ptr(rays), len(rays), 1 if closest else 0, _ptr(out)), "trace_rays")
        return out

    def triangle_ids(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty((len(ids), 2), dtype=np.uint32)
        self._ck(self.lib.vkrt_triangle_ids(self.h, _ptr(ids), len(ids), _ptr(out)), "triangle_ids")
        return out

    def mark(self, which):
        self._ck(self.lib.vkrt_mark(self.h, which), "mark")

    def elapsed_ms(self):
        ms = C.c_float()
        self._ck(self.lib.vkrt_elapsed_ms(self.h, C.byref(ms)), "elapsed_ms")
        return ms.value

    def stats(self):
        s = S.Stats()
        self._ck(self.lib.vkrt_get_stats(self.h, C.byref(s)), "get_stats")
        return s

    def reset_stats(self):
        self._ck(self.lib.vkrt_reset_stats(self.h), "reset_stats")


def upload_scene(target, scene: SceneDesc, width, height, packer=None, build_mode=0, env_accel=None):
    """Feeds `scene` to a backend exposing the Context call surface (the CUDA library, or the oracle in tests),
    mirroring Scene::load + AccelStructure::create + HdrSampling::loadEnvironment + Scene::updateCamera."""
    import sys
    packer = packer or sys.modules[__name__]
    target.scene_begin()
    for m in scene.meshes:
        target.add_primmesh(packer.pack_vertices(m), m.indices, m.material)
    for mat, mesh in scene.nodes:
        target.add_node(column_major(mat), mesh)
    target.set_materials(scene.materials)
    target.set_lights(scene.lights)
    for t in scene.textures:
        target.add_texture(t)
    target.commit(build_mode)
    target.set_environment(scene.env, env_accel)
    cam = scene.camera
    nb = 0 if scene.lights is None else len(scene.lights)
    target.set_camera(packer.make_camera(cam["eye"], cam["center"], cam["up"], cam["fov"], width / height, cam.get("aperture", 0.0), nb))
    target.output_create(width, height)
