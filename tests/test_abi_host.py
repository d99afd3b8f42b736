"""CPU tests of the drop-in boundary: the C ABI loads and exports every declared symbol, struct layouts match
the reference, and the host-side halves (vertex packing, alias table, camera) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import small_sponza

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(vk):
    hdr = open(os.path.join(ROOT, "include", "vkrt_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(vkrt_[a-z0-9_]+)\s*\(", hdr)))
    from vk_raytrace_b200.lib import EXPORTS
    assert declared and set(declared) == set(EXPORTS), (set(declared) ^ set(EXPORTS))
    lib = vk.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_reference_layout(vk):
    from vk_raytrace_b200 import structs as S
    assert C.sizeof(S.SceneCamera) == 140
    assert C.sizeof(S.RtxState) == 48
    assert C.sizeof(S.Tonemapper) == 48
    assert C.sizeof(S.SunAndSky) == 96
    assert S.VertexAttributes.itemsize == 32
    assert S.GltfShadeMaterial.itemsize == 216
    assert S.Light.itemsize == 64
    assert S.EnvAccel.itemsize == 16
    # C side: the header static_asserts the same numbers; compile it as plain C
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include <stddef.h>\n#include <stdio.h>\n#include "vkrt_b200.h"\nint main(void){'
                             'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vkrt_stats), sizeof(vkrt_render_opts), offsetof(vkrt_render_opts, precise_shading),'
                             ' offsetof(vkrt_render_opts, profile), offsetof(vkrt_stats, node_bytes), offsetof(vkrt_stats, bvh_width), offsetof(vkrt_stats, ms_build));'
                             'return 0;}\n')
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "t")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    want = [C.sizeof(S.Stats), C.sizeof(S.RenderOpts), S.RenderOpts.precise_shading.offset, S.RenderOpts.profile.offset,
            S.Stats.node_bytes.offset, S.Stats.bvh_width.offset, S.Stats.ms_build.offset]
    assert got == want and got[0] == 192, (got, want)


def test_create_fails_loudly_without_gpu(vk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = vk.load_library()
    h = C.c_void_p()
    rc = lib.vkrt_create(0, C.byref(h))
    assert rc != 0 and not h.value   # no CPU fallback
    with pytest.raises(RuntimeError):
        vk.Context(0)


def test_vertex_packing_matches_oracle_bit_for_bit(vk, ob):
    from vk_raytrace_b200.context import pack_vertices
    sc = small_sponza()
    for m in sc.meshes[:12]:
        a, b = pack_vertices(m), ob.pack_vertices(m)
        assert a.tobytes() == b.tobytes()
    # handedness bit and special values
    from vk_raytrace_b200.context import MeshDesc
    n = 6
    m = MeshDesc(np.zeros((n, 3), np.float32), np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -1, 0], [.5, .5, -.7071], [np.inf, 0, 0]], np.float32),
                 np.array([[1, 0, 0, 1], [1, 0, 0, -1], [0, 1, 0, 1], [0, 0, 1, -1], [0, 1, 0, 1], [0, 1, 0, -1]], np.float32),
                 np.array([[0.25, 0.75]] * n, np.float32), np.array([[0, 0.5, 1, 1.5]] * n, np.float32), np.arange(6, dtype=np.uint32), 0)
    a, b = pack_vertices(m), ob.pack_vertices(m)
    assert a.tobytes() == b.tobytes()
    assert (a["texcoord"][:, 1].view(np.uint32) & 1).tolist() == [1, 0, 1, 0, 1, 0]
    assert a["normal"][5] == 0xFFFFFFFF and a["color"][0] == 0xFFFF8000


def test_alias_table_matches_oracle_bit_for_bit(vk, ob):
    from vk_raytrace_b200 import scenes
    from vk_raytrace_b200.context import build_env_accel
    for env in (scenes.sky_environment(64, 32, 3), scenes.constant_environment(16, 8, 1.0), scenes.sky_environment(128, 64, 11, sun_power=5000.0)):
        a, ia, aa = build_env_accel(env)
        b, ib, ab = ob.build_env_accel(env)
        assert a.tobytes() == b.tobytes() and ia == ib and aa == ab


def test_alias_table_invariants(vk):
    """Walker alias table: every texel's total selection probability equals importance / sum (src/hdr_sampling.cpp:107-176)."""
    from vk_raytrace_b200 import scenes
    from vk_raytrace_b200.context import build_env_accel
    env = scenes.sky_environment(64, 32, 5)
    acc, integral, avg = build_env_accel(env)
    h, w = env.shape[:2]
    n = h * w
    assert acc["alias"].max() < n
    q = np.clip(acc["q"].astype(np.float64), 0, 1)
    prob = q / n
    np.add.at(prob, acc["alias"], (1 - q) / n)
    theta = (np.arange(h + 1)) * np.pi / h
    area = (np.cos(theta[:-1]) - np.cos(theta[1:])) * (2 * np.pi / w)
    imp = (env[..., :3].max(axis=2) * area[:, None]).reshape(-1).astype(np.float64)
    want = imp / imp.sum()
    assert abs(imp.sum() - integral) / integral < 1e-4
    # (the construction stops with a few above-average texels left at q > 1, which the sampler treats as q = 1:
    #  the reference's table is exact up to that residue)
    assert np.all(np.abs(prob - want) <= 2e-2 * want + 1e-9) and np.median(np.abs(prob - want) / want) < 1e-4 and abs(prob.sum() - 1) < 1e-3
    # pdf = max(rgb) / integral, aliasPdf = pdf[alias]
    assert np.allclose(acc["pdf"], env[..., :3].max(axis=2).reshape(-1) / integral, rtol=1e-6)
    assert np.array_equal(acc["aliasPdf"], acc["pdf"][acc["alias"]])
    assert abs(avg - (env[..., 0] * 0.2126 + env[..., 1] * 0.7152 + env[..., 2] * 0.0722).mean()) < 1e-4 * avg


def test_make_camera(vk):
    from vk_raytrace_b200.context import make_camera
    cam = make_camera((1, 2, 5), (0, 0.5, 0), (0, 1, 0), 45.0, 16 / 9, 0.1, 3)
    vi = np.array(cam.viewInverse, dtype=np.float64).reshape(4, 4).T   # column-major -> math
    pi = np.array(cam.projInverse, dtype=np.float64).reshape(4, 4).T
    assert np.allclose(vi[:3, 3], (1, 2, 5), atol=1e-5)                 # viewInverse * (0,0,0,1) = eye
    fwd = -vi[:3, 2]
    want = np.array([0, 0.5, 0]) - np.array([1, 2, 5])
    assert np.allclose(fwd, want / np.linalg.norm(want), atol=1e-5)
    assert abs(cam.focalDist - np.linalg.norm(want)) < 1e-5 and cam.nbLights == 3 and abs(cam.aperture - 0.1) < 1e-7
    # projInverse maps NDC (0,0,1) onto the -Z axis and y is flipped (proj[1][1] *= -1)
    t = pi @ np.array([0, 0, 1, 1.0])
    assert abs(t[0]) < 1e-6 and abs(t[1]) < 1e-6 and t[2] < 0
    up = pi @ np.array([0, -1, 1, 1.0])
    assert up[1] / abs(up[2]) > 0 and abs(up[1] / -up[2] - np.tan(np.radians(45.0) / 2)) < 1e-5


def test_cpp_renderer_adaptor_compiles_and_fails_loudly_without_gpu(vk):
    """The C++ `CudaRenderer : Renderer` adaptor builds against the C ABI; without a GPU it reports the error."""
    import subprocess, torch
    import __graft_entry__ as g
    host = os.path.join(ROOT, "vk_raytrace_b200", "host")
    exe = os.path.join(host, "vkrt_headless")
    if not os.path.exists(exe):
        g.build()
    assert os.path.exists(exe)
    if not torch.cuda.is_available():
        out = subprocess.run([exe, "1"], capture_output=True, text=True, cwd=host)
        assert out.returncode == 2 and "vkrt_create" in out.stderr


def test_adaptor_compiles_against_reference_header(tmp_path):
    """INTEGRATION.md's drop-in recipe, compiled for real: `CudaRenderer : Renderer` against the reference's own
    src/renderer.h + shaders/host_device.h (nvvk / Vulkan / glm reduced to the 40-line stubs under tests/stubs).  With the
    reference's structs in scope include/vkrt_host_device.h must only assert their sizes, and the adaptor must not add
    anything the real base class lacks (it has no virtual destructor).  Also links the object against the library."""
    import subprocess
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "renderer.h")):
        pytest.skip("reference checkout not present on this box")
    host = os.path.join(ROOT, "vk_raytrace_b200", "host")
    obj = str(tmp_path / "cuda_renderer_ref.o")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-DVKRT_HAVE_NVPRO", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ref, "src"),
           "-I", ref, "-c", os.path.join(host, "cuda_renderer.cpp"), "-o", obj]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    # a translation unit of the application's shape: holds the back-end through the abstract interface, like
    # std::array<Renderer*, eNone> m_pRender (src/sample_example.hpp:136)
    app = tmp_path / "app.cpp"
    app.write_text('#include "cuda_renderer.hpp"\n'
                   'static_assert(sizeof(RtxState) == 48 && sizeof(GltfShadeMaterial) == 216, "reference structs in scope");\n'
                   'int main() { CudaRenderer b(0); Renderer* r = &b; RtxState s{}; r->setPushContants(s); return r->name().empty() ? 1 : 0; }\n')
    exe = str(tmp_path / "app")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-DVKRT_HAVE_NVPRO", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ref, "src"), "-I", ref,
           "-I", host, str(app), obj, "-L", os.path.join(ROOT, "vk_raytrace_b200"), "-lvkrt_b200", "-Wl,-rpath," + os.path.join(ROOT, "vk_raytrace_b200"), "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    assert subprocess.run([exe]).returncode == 0


def test_host_library_loads_and_exports(vk):
    """libvkrt_host.so (include/vkrt_host.h): every declared symbol is exported; a bad path is an error with a message, no GPU needed."""
    from vk_raytrace_b200 import lib as L
    hdr = open(os.path.join(ROOT, "include", "vkrt_host.h")).read()
    declared = sorted(set(re.findall(r"\b(vkrt_host_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(L.HOST_EXPORTS)
    host = L.load_host_library()
    for name in declared:
        assert hasattr(host, name)
    info = L.HostSceneInfo()
    assert C.sizeof(info) == 96
    rc = host.vkrt_host_load_gltf(None, b"/nonexistent.gltf", 0, C.byref(info))
    assert rc != 0 and b"null" in host.vkrt_host_last_error()
