"""GPU parity tests (run with -m gpu on a B200): the CUDA wavefront path, called through the C ABI, against the CPU
oracle on the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full size -- through
size-independent properties (determinism, batching invariance, shard-union == whole).

Tolerances: traversal results (t, u, v, triangle id) and integer pieces are bit-exact.  Images are floating point:
BASELINE.json's north_star allows 1e-3 relative per-channel L2; CUDA and glibc transcendentals differ by ulps, so
measured values sit around 1e-7..1e-5 and the tests assert 1e-4 (images) with a per-pixel outlier budget."""
import os

import numpy as np
import pytest

from helpers import random_rays, rel_l2, small_sponza, triangle_soup

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IMG_TOL = 1e-4      # relative L2 per channel (north_star allows 1e-3)


def per_channel_rel_l2(a, b):
    return max(rel_l2(a[..., c], b[..., c]) for c in range(3))


def both(vk, ob, gpu, scene, W, H, build_mode=0):
    o = ob.Oracle()
    vk.upload_scene(gpu, scene, W, H, build_mode=build_mode)
    vk.upload_scene(o, scene, W, H)
    return o


# ---------------------------------------------------------------------------------------------- traversal
@pytest.mark.parametrize("which,build_mode", [("cornell", 0), ("sponza", 0), ("sponza", 1), ("soup", 0), ("soup_mirrored_double", 1),
                                              ("cornell", 3), ("sponza", 3), ("soup", 3), ("soup_mirrored_double", 2),
                                              ("cornell", 4), ("sponza", 4), ("soup", 4), ("soup_mirrored_double", 4)])
def test_trace_rays_bit_exact(vk, ob, gpu, which, build_mode):
    from vk_raytrace_b200 import scenes
    sc = {"cornell": lambda: scenes.cornell_box(), "sponza": small_sponza, "soup": lambda: triangle_soup(500, 1),
          "soup_mirrored_double": lambda: triangle_soup(500, 2, double_sided=True, mirrored=True)}[which]()
    o = both(vk, ob, gpu, sc, 16, 16, build_mode)
    n = 20000
    rays = random_rays(n, 11, 14.0 if which == "sponza" else 3.0, centre=(0, 3, 0) if which == "sponza" else (0, 0, 0))
    rays[: n // 4, 3] = 2.5
    for closest in (True, False):
        a, b = gpu.trace_rays(rays, closest), o.trace_rays(rays, closest)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    hit = b[:, 0] < 1e31
    assert hit.sum() > n // 20
    # flattened triangle id -> (node, primitive) mapping agrees too
    ids = gpu.trace_rays(rays, True)[:, 3].view(np.uint32)[gpu.trace_rays(rays, True)[:, 0] < 1e31][:512]
    assert np.array_equal(gpu.triangle_ids(ids), o.triangle_ids(ids))


def test_trace_matches_golden_probe(vk, gpu):
    from vk_raytrace_b200 import scenes
    g = np.load(os.path.join(G, "cornell64.npz"))
    vk.upload_scene(gpu, scenes.cornell_box(with_light=True), 64, 64)
    assert np.array_equal(gpu.trace_rays(g["probe_rays"], True).view(np.uint32), g["probe_hits_closest"].view(np.uint32))
    assert np.array_equal(gpu.trace_rays(g["probe_rays"], False).view(np.uint32), g["probe_hits_any"].view(np.uint32))


def test_degenerate_scenes(vk, ob, gpu):
    """Edge cases: empty scene (no nodes), a single triangle, degenerate (zero-area) triangles, rays with zero direction components."""
    from vk_raytrace_b200 import scenes, structs as S
    from vk_raytrace_b200.context import SceneDesc
    empty = SceneDesc()
    empty.materials = S.default_material(1)
    empty.env = scenes.constant_environment(8, 4, 0.5)
    empty.camera = dict(eye=(0, 0, 3), center=(0, 0, 0), up=(0, 1, 0), fov=45.0)
    o = both(vk, ob, gpu, empty, 16, 8)
    rays = random_rays(256, 1)
    assert np.array_equal(gpu.trace_rays(rays).view(np.uint32), o.trace_rays(rays).view(np.uint32))
    st = S.default_rtx_state(16, 8, maxDepth=3)
    gpu.render(st, 1)
    o.render(st, 1)
    assert np.allclose(gpu.output_read(), o.output_read(), rtol=1e-6) and np.allclose(gpu.output_read()[..., :3], 0.5)
    one = SceneDesc()
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [2, 2, 2], [2, 2, 2], [2, 2, 2]], np.float32)   # second triangle is degenerate
    one.meshes = [scenes.finish_mesh(pos, [0, 1, 2, 3, 4, 5], None, 0, smooth=False)]
    one.nodes = [(np.eye(4, dtype=np.float32), 0)]
    one.materials, one.env, one.camera = empty.materials, empty.env, empty.camera
    o = both(vk, ob, gpu, one, 16, 8)
    rays = random_rays(4096, 2, 1.5)
    rays[:64, 4:7] = (0, 0, -1)        # axis aligned: zero components
    rays[:64, 2] = 2.0
    assert np.array_equal(gpu.trace_rays(rays).view(np.uint32), o.trace_rays(rays).view(np.uint32))
    assert (gpu.trace_rays(rays)[:, 0] < 1e31).sum() > 10


# ---------------------------------------------------------------------------------------------- images
@pytest.mark.parametrize("pbr", [0, 1])
def test_cornell_matches_oracle_and_golden(vk, ob, gpu, pbr):
    from vk_raytrace_b200 import scenes, structs as S
    g = np.load(os.path.join(G, "cornell64.npz"))
    W = H = 64
    o = both(vk, ob, gpu, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=4, pbrMode=pbr)
    gpu.reset_stats()
    gpu.render(st, 4)
    o.render(st, 4)
    a, b = gpu.output_read(), o.output_read()
    assert per_channel_rel_l2(a[..., :3], b[..., :3]) < IMG_TOL
    assert per_channel_rel_l2(a[..., :3], g[f"radiance_pbr{pbr}"]) < IMG_TOL
    assert np.all(a[..., 3] == 1.0)
    # opaque scene: the two implementations trace exactly the same rays
    s, c = gpu.stats(), o.counters()
    assert (s.closest_rays, s.shadow_rays) == (c["closest_rays"], c["shadow_rays"])


@pytest.mark.parametrize("dbg", list(range(1, 12)))
def test_debug_views(vk, ob, gpu, dbg):
    """C1 of BASELINE.json: every DebugMode (shaders/host_device.h:88-102) at depth 1 (9-11: depth 3)."""
    from vk_raytrace_b200 import scenes, structs as S
    g = np.load(os.path.join(G, "cornell64.npz"))
    W = H = 64
    vk.upload_scene(gpu, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=1 if dbg < 9 else 3, debugging_mode=dbg)
    gpu.render(st, 1)
    a = gpu.output_read()[..., :3]
    want = g[f"debug{dbg}"]
    assert np.abs(a - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("pbr,seed_mode", [(0, 0), (1, 0), (0, 1)])
def test_sponza_with_alpha_textures_lights(vk, ob, gpu, pbr, seed_mode):
    """Textured scene with MASK-alpha drapes, punctual + spot lights, normal maps, transmission, clearcoat, anisotropy."""
    from vk_raytrace_b200 import structs as S
    W, H = 96, 64
    o = both(vk, ob, gpu, small_sponza(), W, H)
    assert gpu.stats().has_non_opaque == 1
    st = S.default_rtx_state(W, H, maxDepth=6, pbrMode=pbr)
    opts = S.RenderOpts()
    opts.seed_mode = seed_mode
    gpu.render(st, 4, opts)
    o.render(st, 4, seed_mode=seed_mode)
    a, b = gpu.output_read()[..., :3], o.output_read()[..., :3]
    assert np.isfinite(a).all()
    # a few pixels may take a different branch because of ulp differences in sin/cos/pow: bounded count, bounded L2
    # (at this tiny size one divergent path is already 1e-2 of the image norm, so the L2 bound is taken over the rest)
    off = (np.abs(a - b) > 1e-3 * (1 + np.abs(b))).any(axis=2)
    assert off.sum() <= 0.005 * W * H, off.sum()
    assert per_channel_rel_l2(a[~off], b[~off]) < IMG_TOL


def test_multi_sample_frames_and_clamp(vk, ob, gpu):
    """maxSamples > 1 (the RNG stream continues across samples, pathtrace.comp:101-106) and the firefly clamp."""
    from vk_raytrace_b200 import scenes, structs as S
    W = H = 48
    o = both(vk, ob, gpu, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=4, maxSamples=3, fireflyClampThreshold=0.4, frame=2)
    gpu.output_clear()
    gpu.render(st, 2)
    o.render(st, 2)
    assert per_channel_rel_l2(gpu.output_read()[..., :3], o.output_read()[..., :3]) < IMG_TOL


def test_sun_and_sky_environment(vk, ob, gpu):
    from vk_raytrace_b200 import scenes, structs as S
    W, H = 64, 40
    sc = scenes.cornell_box()
    sc.nodes = sc.nodes[:1] + sc.nodes[6:]     # floor + the two boxes under an open sky
    o = both(vk, ob, gpu, sc, W, H)
    ss = S.default_sunsky(in_use=1)
    gpu.set_sunsky(ss)
    o.set_sunsky(ss)
    st = S.default_rtx_state(W, H, maxDepth=4)
    gpu.render(st, 3)
    o.render(st, 3)
    a, b = gpu.output_read()[..., :3], o.output_read()[..., :3]
    assert a.max() > 0 and per_channel_rel_l2(a, b) < 1e-3
    gpu.set_sunsky(S.default_sunsky())


def test_tonemap(vk, ob, gpu):
    from vk_raytrace_b200 import scenes, structs as S
    g = np.load(os.path.join(G, "cornell64.npz"))
    W = H = 64
    o = both(vk, ob, gpu, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=4)
    gpu.render(st, 4)
    o.render(st, 4)
    for tm in (S.default_tonemapper(), S.default_tonemapper(dither=0, contrast=1.2, brightness=0.8, saturation=0.5, vignette=0.3, avgLum=1.7)):
        a8, b8 = gpu.tonemap(tm).astype(int), o.tonemap(tm).astype(int)
        assert np.abs(a8 - b8).max() <= 1 and (a8 != b8).mean() < 0.01
        af, bf = gpu.tonemap_float(tm), o.tonemap_float(tm)
        assert np.abs(af - bf).max() < 5e-3 and rel_l2(af, bf) < 1e-3      # dithering quantises: a flipped step is 1/255
    a8 = gpu.tonemap(S.default_tonemapper()).astype(int)
    assert np.abs(a8 - g["tonemapped_rgba8"].astype(int)).max() <= 1


def test_tonemap_auto_exposure(vk, ob, gpu):
    """post.frag:101-111 with autoExposure bit0 (global, toneExposure) and bit0|bit1 (toneLocalExposure) on a non-power-of-two
    image: the mip pyramid (RenderOutput::genMipmap blit chain) and both operators against the oracle."""
    from vk_raytrace_b200 import scenes, structs as S
    W, H = 100, 72
    o = both(vk, ob, gpu, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=4)
    gpu.render(st, 4)
    o.output_write(gpu.output_read())          # identical HDR input: this test is about the post chain only
    for ae, zoom in ((1, 1.0), (3, 1.0), (3, 0.75)):
        tm = S.default_tonemapper(dither=0, autoExposure=ae, zoom=zoom)
        af, bf = gpu.tonemap_float(tm), o.tonemap_float(tm)
        assert np.isfinite(af[..., :3]).all() == np.isfinite(bf[..., :3]).all()
        m = np.isfinite(bf).all(-1) & np.isfinite(af).all(-1)
        assert m.mean() > 0.95
        assert np.abs(af[m] - bf[m]).max() < 2e-3 and rel_l2(af[m], bf[m]) < 1e-4, (ae, zoom)
        a8, b8 = gpu.tonemap(tm).astype(int), o.tonemap(tm).astype(int)
        assert np.abs(a8 - b8)[m].max() <= 1
    # the operator does something: differs from the plain path
    assert rel_l2(gpu.tonemap_float(S.default_tonemapper(dither=0, autoExposure=1)), gpu.tonemap_float(S.default_tonemapper(dither=0))) > 1e-2


# ---------------------------------------------------------------------------------------------- invariances
def test_batching_and_sharding_do_not_change_the_image(vk, gpu):
    """frames_in_flight only changes scheduling; tile shards of N contexts sum to the whole image (bit-exact)."""
    from vk_raytrace_b200 import structs as S
    W, H = 200, 120
    vk.upload_scene(gpu, small_sponza(), W, H)
    st = S.default_rtx_state(W, H, maxDepth=5)
    imgs = []
    for fif in (1, 3, 8):
        opts = S.RenderOpts()
        opts.frames_in_flight = fif
        gpu.output_clear()
        gpu.render(st, 6, opts)
        imgs.append(gpu.output_read())
    assert np.array_equal(imgs[0], imgs[1]) and np.array_equal(imgs[0], imgs[2])
    # the optional ray sort (octant + origin Morton key) only reorders the queue: same image, bit for bit
    opts = S.RenderOpts()
    opts.sort_rays, opts.frames_in_flight = 1, 3
    gpu.output_clear()
    gpu.render(st, 6, opts)
    assert np.array_equal(imgs[0], gpu.output_read()) and gpu.stats().ms_sort >= 0.0
    total = np.zeros_like(imgs[0])
    for rank in range(3):
        opts = S.RenderOpts()
        opts.shard_rank, opts.shard_count, opts.tile_size = rank, 3, 32
        gpu.output_clear()
        gpu.render(st, 6, opts)
        part = gpu.output_read()
        assert (part[..., 3] == 1).sum() > 0.2 * W * H
        total += part
    assert np.array_equal(total, imgs[0])
    # the fused tail kernel (late bounces of small wavefronts in one launch) against one launch per stage: never, from
    # bounce 1 / 2 / 3, and the megakernel extreme (from bounce 0) -- bit-identical, both shade builds, same ray counts
    for precise in (True, False):
        ref_img, ref_rays = None, None
        for tail in (-1, 3, 1, 2, 0, 1 << 20):
            opts = S.RenderOpts()
            opts.tail_from, opts.frames_in_flight = tail, 2
            gpu.output_clear()
            gpu.reset_stats()
            gpu.render(st, 6, opts, precise=precise)
            img, stt = gpu.output_read(), gpu.stats()
            rays = (int(stt.closest_rays), int(stt.shadow_rays))
            if ref_img is None:
                ref_img, ref_rays = img, rays
                if precise:
                    assert np.array_equal(img, imgs[0])
            assert np.array_equal(img, ref_img), (precise, tail)
            assert rays == ref_rays, (precise, tail, rays, ref_rays)
    # frame-by-frame calls == one multi-frame call
    gpu.output_clear()
    for f in range(6):
        st1 = S.default_rtx_state(W, H, maxDepth=5, frame=f)
        gpu.render(st1, 1)
    assert np.array_equal(gpu.output_read(), imgs[0])


def test_external_output_buffer_and_errors(vk, gpu):
    import torch
    from vk_raytrace_b200 import scenes, structs as S
    W = H = 32
    vk.upload_scene(gpu, scenes.cornell_box(), W, H)
    st = S.default_rtx_state(W, H, maxDepth=2)
    gpu.render(st, 1)
    want = gpu.output_read()
    buf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
    gpu.output_create(W, H, buf.data_ptr())
    assert gpu.output_device_ptr() == buf.data_ptr()
    gpu.render(st, 1)
    assert np.array_equal(buf.cpu().numpy(), want)
    gpu.output_create(W, H)
    bad = S.default_rtx_state(W + 1, H)
    with pytest.raises(RuntimeError):
        gpu.render(bad, 1)


def _temperature(v):
    """shaders/common.glsl:38-62 in numpy."""
    def smooth(x):
        x = np.clip(x, 0.0, 1.0)
        return x * x * (3.0 - 2.0 * x)

    def fade(lo, hi, x):
        mid, rng = (lo + hi) * 0.5, (hi - lo) * 0.5
        return smooth(1.0 - np.clip(np.abs(mid - x) / rng, 0.0, 1.0))
    v = np.asarray(v, np.float64)[..., None]
    cols = np.array([[0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 1, 0], [1, 0, 0]], np.float64)
    return (fade(-0.25, 0.25, v) * cols[0] + fade(0.0, 0.5, v) * cols[1] + fade(0.25, 0.75, v) * cols[2] + fade(0.5, 1.0, v) * cols[3]
            + smooth((v - 0.75) / 0.25) * cols[4])


def test_heatmap_debug_mode(vk, gpu):
    """Debug mode 12 (pathtrace.comp:108-117): the reference maps the nanoseconds its invocation took through temperature(); the
    wavefront kernels add the nanoseconds (%globaltimer) a lane held each path -- traversal of its rays, shading of its bounces.
    Like any clock it is not reproducible; checked: the colour map, the range handling, that the measure is a time (sky pixels, whose
    paths end at the first miss, are cheaper than geometry; two runs agree on the picture) and that it leaves the renderer intact."""
    from vk_raytrace_b200 import structs as S
    W, H = 160, 96
    vk.upload_scene(gpu, small_sponza(), W, H)
    hi = 4_000_000          # ns: a latency-bound wavefront keeps a path far longer than an RT-core invocation (reference default 65 us)
    st = S.default_rtx_state(W, H, maxDepth=5, debugging_mode=12, minHeatmap=0, maxHeatmap=hi)
    gpu.render(st, 1)
    a = gpu.output_read()[..., :3]
    gpu.output_clear()
    gpu.render(st, 1)
    b = gpu.output_read()[..., :3]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    # every pixel is a point of the temperature curve: recover the intensity from the (piecewise unique) colour and re-evaluate
    grid = np.linspace(0.0, 1.0, 20001)
    lut = _temperature(grid)

    def intensity(img):
        flat = img.reshape(-1, 3).astype(np.float64)[::7]
        idx = np.array([np.argmin(((lut - c) ** 2).sum(1)) for c in flat])
        assert np.abs(lut[idx] - flat).max() < 2e-3
        return grid[idx]
    va, vb = intensity(a), intensity(b)
    assert va.max() > va.min() and len(np.unique(np.round(va, 3))) > 20          # a distribution, not a constant
    assert np.corrcoef(va, vb)[0, 1] > 0.5                                        # two runs: the same picture, not the same numbers
    # a time: paths that leave through the open roof at the first bounce cost less than paths that bounce around inside.
    # Reference picture: radiance debug view at depth 1 is black exactly where the primary ray missed everything.
    gpu.output_clear()
    gpu.render(S.default_rtx_state(W, H, maxDepth=1, debugging_mode=5), 1)        # alpha view: 0 on a miss, 1 on geometry
    hit = (gpu.output_read()[..., 0].reshape(-1)[::7] > 0.5)
    if hit.any() and (~hit).any():
        assert va[~hit].mean() < va[hit].mean()
    # range handling: everything below min -> temperature(0) = blue, everything above max -> temperature(1) = red
    for lo, hi2, want in ((1_000_000_000, 2_000_000_000, (0.0, 0.0, 1.0)), (0, 1, (1.0, 0.0, 0.0))):
        gpu.output_clear()
        gpu.render(S.default_rtx_state(W, H, maxDepth=5, debugging_mode=12, minHeatmap=lo, maxHeatmap=hi2), 1)
        img = gpu.output_read()[..., :3]
        assert np.abs(img - np.array(want, np.float32)).max() < 1e-6
    # the renderer is untouched by the instrumentation: a normal render afterwards equals one before
    st0 = S.default_rtx_state(W, H, maxDepth=5)
    gpu.output_clear()
    gpu.render(st0, 2)
    ref = gpu.output_read()
    gpu.output_clear()
    gpu.render(st, 1)
    gpu.output_clear()
    gpu.render(st0, 2)
    assert np.array_equal(gpu.output_read(), ref)


@pytest.mark.parametrize("build_mode", [0, 1, 3, 4])
def test_full_size_properties_sponza_1080p(vk, gpu, build_mode):
    """BASELINE.json configs[2] at full size (262k triangles, 1920x1080, depth 8): determinism, batching invariance,
    ray accounting, and agreement of sampled pixels with the oracle (the oracle only renders a band of rows)."""
    from oracle import binding as ob
    from vk_raytrace_b200 import scenes, structs as S
    W, H = 1920, 1080
    sc = scenes.sponza_standin(detail=1.0, tex_size=256, env_size=(512, 256))
    vk.upload_scene(gpu, sc, W, H, build_mode=build_mode)
    s = gpu.stats()
    assert 240_000 < s.num_triangles < 290_000 and s.bvh_max_depth < 128
    st = S.default_rtx_state(W, H, maxDepth=8)
    gpu.reset_stats()
    gpu.render(st, 3)
    a = gpu.output_read()
    s1 = gpu.stats()
    opts = S.RenderOpts()
    opts.frames_in_flight = 1
    gpu.output_clear()
    gpu.render(st, 3, opts)
    b = gpu.output_read()
    assert np.array_equal(a, b) and np.isfinite(a).all()
    assert 3 * W * H <= s1.closest_rays <= 3 * W * H * 8 and s1.shadow_rays <= s1.closest_rays
    o = ob.Oracle()
    vk.upload_scene(o, sc, W, H)
    rows = (500, 508)
    o.render(st, 3, rows=rows)
    ref = o.output_read()[rows[0]:rows[1], :, :3]
    got = a[rows[0]:rows[1], :, :3]
    off = (np.abs(got - ref) > 1e-3 * (1 + np.abs(ref))).any(axis=2)
    assert off.mean() < 0.002 and per_channel_rel_l2(got[~off], ref[~off]) < IMG_TOL


def test_full_size_properties_10m_triangles_and_helmet(vk, gpu):
    """BASELINE.json configs[4] (10.4 M triangles, BVH in HBM, 4-wide tree) and configs[1] (single PBR mesh) at full size:
    the three GPU builders give bit-identical hits for a million random rays (the result does not depend on the tree), and
    a render is deterministic, finite and accounts for its rays."""
    from vk_raytrace_b200 import scenes, structs as S
    W, H = 1920, 1080
    sc = scenes.instanced_stress(instances_per_axis=10)
    rays = random_rays(1_000_000, 7, extent=6.5, centre=(0.0, 5.5, 0.0))
    hits = []
    for mode in (0, 1, 3, 4):
        vk.upload_scene(gpu, sc, W, H, build_mode=mode)
        s = gpu.stats()
        assert s.num_triangles > 10_000_000 and s.bvh_width == 4 and s.bvh_max_depth <= 120
        hits.append((gpu.trace_rays(rays, True).view(np.uint32), gpu.trace_rays(rays, False).view(np.uint32)))
    assert (hits[0][0][:, 3] != 0xffffffff).mean() > 0.01            # the probe does hit geometry
    for h in hits[1:]:
        assert np.array_equal(hits[0][0], h[0]) and np.array_equal(hits[0][1], h[1])
    st = S.default_rtx_state(W, H, maxDepth=8)
    gpu.reset_stats()
    gpu.render(st, 2, precise=False)
    a = gpu.output_read()
    s1 = gpu.stats()
    gpu.output_clear()
    gpu.render(st, 2, precise=False)
    assert np.array_equal(a, gpu.output_read()) and np.isfinite(a).all()
    assert 2 * W * H <= s1.closest_rays <= 2 * W * H * 8 and s1.shadow_rays <= s1.closest_rays
    # configs[1] stand-in
    vk.upload_scene(gpu, scenes.helmet_standin(level=6, tex_size=512, env_size=(512, 256)), W, H, build_mode=2)
    st = S.default_rtx_state(W, H, maxDepth=6)
    imgs = []
    for fif in (0, 1):
        o = S.RenderOpts()
        o.frames_in_flight = fif
        gpu.output_clear()
        gpu.render(st, 2, o, precise=False)
        imgs.append(gpu.output_read())
    assert np.array_equal(imgs[0], imgs[1]) and np.isfinite(imgs[0]).all() and imgs[0][..., :3].mean() > 0.01


@pytest.mark.parametrize("pbr", [0, 1])
def test_fast_shading_against_precise_and_oracle(vk, ob, gpu, pbr):
    """The default shade kernel (approximate division / sqrt, SFU sin/cos/exp/log/pow -- what a GLSL compiler emits too) against
    the IEEE-exact one and the oracle.  Primary rays and traversal are exact in both, so frame-0 geometry is identical; radiance
    differs by rounding, and a small fraction of multi-bounce paths takes another branch (chaotic in glossy scenes)."""
    from vk_raytrace_b200 import scenes, structs as S
    for sc, W, H, depth, max_off in ((scenes.cornell_box(with_light=True), 64, 64, 4, 0.002), (small_sponza(), 96, 64, 6, 0.02)):
        o = both(vk, ob, gpu, sc, W, H)
        st = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=pbr)
        gpu.render(st, 4, precise=True)
        exact = gpu.output_read()[..., :3]
        gpu.output_clear()
        gpu.render(st, 4, precise=False)
        fast = gpu.output_read()[..., :3]
        gpu.output_clear()
        gpu.render(st, 4, precise=False)
        assert np.array_equal(fast, gpu.output_read()[..., :3]) and np.isfinite(fast).all()      # deterministic
        o.render(st, 4)
        ref = o.output_read()[..., :3]
        for other in (exact, ref):
            off = (np.abs(fast - other) > 1e-3 * (1 + np.abs(other))).any(axis=2)
            assert off.mean() <= max_off, off.mean()
            assert per_channel_rel_l2(fast[~off], other[~off]) < 1e-3        # north_star tolerance
        # debug views that do not depend on the path history are identical or within rounding
        for dbg in (1, 2, 5, 7):   # base colour, normal, alpha, texcoord (shaders/host_device.h:88-102)
            sd = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=pbr, debugging_mode=dbg)
            imgs = []
            for precise in (True, False):
                gpu.output_clear()
                gpu.render(sd, 1, precise=precise)
                imgs.append(gpu.output_read()[..., :3])
            bad = (np.abs(imgs[0] - imgs[1]) > 1e-4).any(axis=2)
            assert bad.mean() <= 0.002, (dbg, bad.mean())


@pytest.mark.parametrize("env", [{"VKRT_BVH_WIDTH": "4"}, {"VKRT_NODE_FORMAT": "q32"}, {"VKRT_NODE_FORMAT": "lohi"}, {"VKRT_TRACE_VARIANT": "200"},
                                 {"VKRT_OVERLAP": "0"}, {"VKRT_DUAL_SLOT": "0"}])
def test_non_default_traversal_kernels_in_subprocess(env):
    """The traversal kernel is chosen once per process (environment): the 4-wide tree, the 16-bit grid nodes, the min/max fp32
    nodes, the if-if loop and the single-stream schedule must give the same bit-exact hits and the same images as the default."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                          "-k", "trace_rays_bit_exact or degenerate or cornell_matches or sponza_with_alpha"],
                         capture_output=True, text=True, cwd=root, timeout=600, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_cpp_renderer_adaptor_headless(vk):
    """`CudaRenderer : Renderer` (vk_raytrace_b200/host/) driven like src/main.cpp drives a back-end."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "vk_raytrace_b200", "host", "vkrt_headless")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([exe, "8"], capture_output=True, text=True, cwd=os.path.dirname(exe), timeout=120)
    assert out.returncode == 0, out.stderr
    assert "8 frames" in out.stdout and float(out.stdout.split("mean green")[1].split()[0]) > 20.0


@pytest.mark.parametrize("pbr", [0, 1])
def test_gltf_file_to_image_end_to_end(vk, ob, tmp_path, pbr):
    """The north-star path: a glTF 2.0 file on disk -> C++ ingest -> `CudaRenderer : Renderer` -> image, against the oracle fed
    from the in-memory scene the file was exported from (textures via PNG, environment via PFM)."""
    import subprocess
    from vk_raytrace_b200 import structs as S
    from vk_raytrace_b200.gltf_export import export_gltf
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "vk_raytrace_b200", "host", "vkrt_gltf_render")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    sc = small_sponza()
    gltf, env = export_gltf(sc, str(tmp_path), "s")
    W, H = 96, 64
    _, integral, _ = vk.context.build_env_accel(sc.env)
    o = ob.Oracle()
    vk.upload_scene(o, sc, W, H)
    # The file stores the camera as a node matrix, so the importer's eye/centre/up differ from the generator's by float
    # rounding (~1e-7).  With one bounce that is invisible; with five, the near-specular materials of this scene (roughness
    # 0.02-0.05: D ~ 1/alpha^4) amplify it to percent-level differences in a percent of the pixels -- the same spread two
    # runs of the reference with a nudged camera would show -- so the deep comparison is statistical.
    for frames, depth, max_off, tol in ((4, 1, 0.0005, 1e-4), (4, 5, 0.03, 2e-3)):
        out = subprocess.run([exe, gltf, env, str(W), str(H), str(frames), str(depth), str(tmp_path / "o"), str(pbr)], capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, VKRT_SHADING="precise"))   # the oracle comparison below is for the IEEE-exact shade kernel
        assert out.returncode == 0, out.stderr + out.stdout
        got = np.fromfile(str(tmp_path / "o.f32"), dtype=np.float32).reshape(H, W, 4)[..., :3]
        st = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=pbr, fireflyClampThreshold=integral * 4.0)
        o.output_create(W, H)
        o.render(st, frames)
        ref = o.output_read()[..., :3]
        off = (np.abs(got - ref) > 2e-3 * (1 + np.abs(ref))).any(axis=2)
        assert off.mean() <= max_off, (depth, off.mean())
        assert per_channel_rel_l2(got[~off], ref[~off]) < tol
        assert abs(got.mean() - ref.mean()) < 0.02 * ref.mean()
    assert os.path.getsize(str(tmp_path / "o.ppm")) > W * H * 3


# ---------------------------------------------------------------------------------------------- round 2: reference branches, the timed path
@pytest.mark.parametrize("kind", ["blend", "dof", "directional", "unlit", "hdr0"])
@pytest.mark.parametrize("pbr", [0, 1])
def test_reference_branches(vk, ob, gpu, kind, pbr):
    """Branches of the reference no other scene reaches (tests/helpers.py::branch_scene): BLEND alpha (traceray_rq.glsl:87-99),
    thin lens with aperture 0.1 (pathtrace.glsl:366-374), a directional light (:129-146), an unlit material (:259-262),
    hdrMultiplier = 0 (:113-119).  Both sides are fed by the oracle's own packer and camera (packer=ob)."""
    from helpers import branch_scene
    from vk_raytrace_b200 import structs as S
    W = H = 64
    sc, kw = branch_scene(kind)
    o = ob.Oracle()
    vk.upload_scene(gpu, sc, W, H, packer=ob)
    vk.upload_scene(o, sc, W, H, packer=ob)
    assert gpu.stats().has_non_opaque == (1 if kind == "blend" else 0)
    st = S.default_rtx_state(W, H, maxDepth=4, pbrMode=pbr, **kw)
    gpu.reset_stats()
    gpu.render(st, 4)
    o.render(st, 4)
    a, b = gpu.output_read()[..., :3], o.output_read()[..., :3]
    assert np.isfinite(a).all() and b.mean() > 0.01
    s, c = gpu.stats(), o.counters()
    off = (np.abs(a - b) > 1e-3 * (1 + np.abs(b))).any(axis=2)
    whole = per_channel_rel_l2(a, b)
    info = (kind, pbr, float(whole), int(off.sum()), (int(s.closest_rays), int(s.shadow_rays)), (c["closest_rays"], c["shadow_rays"]))
    if kind == "blend":
        # stochastic opacity multiplies the branches a last-bit difference can flip (every pass-through is a decision): a few
        # pixels may diverge at 4 frames; everything else is exact and the divergent ones are bounded
        assert off.sum() <= 0.005 * W * H, info
        assert per_channel_rel_l2(a[~off], b[~off]) < IMG_TOL, info
        assert abs(int(s.closest_rays) - c["closest_rays"]) <= 0.002 * c["closest_rays"], info
    else:
        assert whole < IMG_TOL, info          # whole image, no mask
        assert (s.closest_rays, s.shadow_rays) == (c["closest_rays"], c["shadow_rays"]), info


def test_blend_probe_bit_exact_and_statistical(vk, ob, gpu):
    from helpers import blend_probe_scene
    alpha = 0.3
    o = ob.Oracle()
    sc = blend_probe_scene(alpha)
    vk.upload_scene(gpu, sc, 8, 8, packer=ob)
    vk.upload_scene(o, sc, 8, 8, packer=ob)
    rng = np.random.default_rng(5)
    n = 40000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:2] = rng.uniform(-0.9, 0.9, (n, 2))
    rays[:, 2], rays[:, 3], rays[:, 6] = 1.0, 1e32, -1.0
    rays[:, 7] = rng.integers(0, 2 ** 32, n, dtype=np.uint32).view(np.float32)
    for closest in (True, False):
        a, b = gpu.trace_rays(rays, closest), o.trace_rays(rays, closest)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert abs((a[:, 0] < 1e31).mean() - alpha) < 4 * np.sqrt(alpha * (1 - alpha) / n) + 1e-3


@pytest.mark.parametrize("pbr", [0])
def test_timed_configuration_matches_oracle(vk, gpu, pbr):
    """Exactly what bench.py times -- the sponza_1080p workload (same generator arguments), build_mode 2 (host SAH tree, the bench
    default below 2 M triangles) or the GPU builder the bench falls back to, the default FAST shade kernel, multi-frame batches --
    against the oracle on a 64-row band.  The fast kernel differs from IEEE arithmetic by rounding, so a bounded fraction of
    multi-bounce paths diverges; everything else must be within the north-star 1e-3, and the band means must agree."""
    from oracle import binding as ob
    from vk_raytrace_b200 import scenes, structs as S
    import bench
    cfg = bench.WORKLOADS["sponza_1080p"]
    W, H = cfg["w"], cfg["h"]
    sc = bench.make_scene(cfg)
    mode = bench.default_build_mode(sc.num_triangles())
    vk.upload_scene(gpu, sc, W, H, build_mode=mode, packer=ob)
    _, integral, _ = ob.build_env_accel(sc.env)
    st = S.default_rtx_state(W, H, maxDepth=cfg["depth"], pbrMode=pbr, fireflyClampThreshold=integral * 4.0)
    gpu.reset_stats()
    gpu.render(st, 4, precise=False)
    a = gpu.output_read()
    o = ob.Oracle()
    vk.upload_scene(o, sc, W, H, packer=ob)
    rows = (508, 572)
    o.render(st, 4, rows=rows)
    ref = o.output_read()[rows[0]:rows[1], :, :3]
    got = a[rows[0]:rows[1], :, :3]
    off = (np.abs(got - ref) > 1e-3 * (1 + np.abs(ref))).any(axis=2)
    assert off.mean() < 0.02, off.mean()
    assert per_channel_rel_l2(got[~off], ref[~off]) < 1e-3
    for c in range(3):
        assert abs(got[..., c].mean() - ref[..., c].mean()) < 0.01 * ref[..., c].mean()


# frames at which the whole-image error is below 1e-3, from the measured curve (scripts/gpu_converge.py, profiles/converge_r02.json).
# Cornell (diffuse + one glossy box) sits at 1e-7 for any count.  small_sponza (64x40, near-specular materials, alpha, two lights) is
# chaotic: a path whose branch flips on a last-bit difference (glibc vs CUDA libm for the IEEE kernel, SFU arithmetic for the fast
# one) becomes an independent sample, and the error falls as 1/sqrt(frames): IEEE 1.3e-3 / 7.6e-4 / 3.5e-4 and fast 3.7e-3 / 2.1e-3 /
# 9.9e-4 at 256 / 1024 / 4096 frames.  Two Vulkan drivers rendering the reference's own GLSL would differ the same way.
CONVERGED = {("cornell", True): 256, ("cornell", False): 256, ("small_sponza", True): 1024, ("small_sponza", False): 8192}


@pytest.mark.parametrize("which", ["cornell", "small_sponza"])
@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("pbr", [0, 1])
def test_converged_whole_image(vk, ob, gpu, which, precise, pbr):
    """North-star tolerance as written: whole image, no outlier mask, per-channel relative L2 < 1e-3 -- at a converged sample
    count, where the few paths that diverge by rounding average out."""
    from vk_raytrace_b200 import scenes, structs as S
    sc, W, H, depth = (scenes.cornell_box(with_light=True), 48, 48, 4) if which == "cornell" else (small_sponza(), 64, 40, 6)
    n = CONVERGED[(which, precise)]
    o = ob.Oracle()
    vk.upload_scene(gpu, sc, W, H, packer=ob)
    vk.upload_scene(o, sc, W, H, packer=ob)
    st = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=pbr)
    gpu.render(st, n, precise=precise)
    o.render(st, n)
    a, b = gpu.output_read()[..., :3], o.output_read()[..., :3]
    assert per_channel_rel_l2(a, b) < 1e-3


def test_async_tonemap_dual_slots_and_shard_exchange(vk, gpu):
    """Round-2 entry points: (1) vkrt_tonemap_async / _wait into page-locked memory gives the bytes of vkrt_tonemap, also while the
    next frames are already rendering on other wavefront slots; (2) one-frame calls rotating through the slots give the
    image of one batched call, bit for bit; (3) pack_shard on N 'ranks' + unpack_shard on the root rebuilds the unsharded image."""
    import torch
    from vk_raytrace_b200 import structs as S
    W, H = 200, 120
    vk.upload_scene(gpu, small_sponza(), W, H)
    tm = S.default_tonemapper()
    st = S.default_rtx_state(W, H, maxDepth=5)
    gpu.output_clear()
    gpu.render(st, 6)
    want_img = gpu.output_read()
    # (1)+(2): interactive loop, frame f's read-back overlaps frame f+1
    bufs = [gpu.host_alloc((H, W, 4)), gpu.host_alloc((H, W, 4))]
    gpu.output_clear()
    frames8 = []
    for f in range(6):
        gpu.render(S.default_rtx_state(W, H, maxDepth=5, frame=f), 1, sync=False)
        gpu.tonemap_wait(f & 1)
        if f >= 2:
            frames8.append(bufs[f & 1].copy())       # frame f-2, complete after the wait
        gpu.tonemap_async(tm, bufs[f & 1], f & 1)
    gpu.tonemap_wait(0)
    gpu.tonemap_wait(1)
    assert np.array_equal(gpu.output_read(), want_img)
    assert np.array_equal(bufs[1], gpu.tonemap(tm))          # frame 5 = the final image
    # the frames read back in flight are the running means after 1, 2, 3, 4 frames
    gpu.output_clear()
    for f in range(4):
        gpu.render(S.default_rtx_state(W, H, maxDepth=5, frame=f), 1)
        assert np.array_equal(frames8[f], gpu.tonemap(tm)), f
    for b in bufs:
        gpu.host_free(b)
    # the same loop four frames deep: a ring of four staging slots, one-frame calls rotating through four wavefront slots
    ring = [gpu.host_alloc((H, W, 4)) for _ in range(4)]
    gpu.output_clear()
    got = {}
    for f in range(10):
        gpu.render(S.default_rtx_state(W, H, maxDepth=5, frame=f), 1, sync=False)
        gpu.tonemap_wait(f % 4)
        if f >= 4:
            got[f - 4] = ring[f % 4].copy()
        gpu.tonemap_async(tm, ring[f % 4], f % 4)
    for k in range(4):
        gpu.tonemap_wait(k)
    for f in range(6, 10):
        got[f] = ring[f % 4].copy()
    gpu.output_clear()
    for f in range(10):
        gpu.render(S.default_rtx_state(W, H, maxDepth=5, frame=f), 1)
        assert np.array_equal(got[f], gpu.tonemap(tm)), f
    with pytest.raises(Exception):
        gpu.tonemap_async(tm, ring[0], 4)                    # only slots 0..3 exist
    for b in ring:
        gpu.host_free(b)
    # (3) shards
    total = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
    parts, n = [], 3
    for rank in range(n):
        o = S.RenderOpts()
        o.shard_rank, o.shard_count, o.tile_size = rank, n, 32
        gpu.output_clear()
        gpu.render(st, 6, o)
        k = gpu.shard_size(32, rank, n)
        buf = torch.zeros((k, 4), dtype=torch.float32, device="cuda:0")
        gpu.pack_shard(32, rank, n, buf.data_ptr())
        gpu.sync()
        parts.append(buf)
    assert sum(len(p) for p in parts) == W * H
    gpu.output_create(W, H, total.data_ptr())
    for rank in range(n):
        gpu.unpack_shard(32, rank, n, parts[rank].data_ptr())
    gpu.sync()
    assert np.array_equal(total.cpu().numpy(), want_img)
    # shard order == the host-side mirror (vk_raytrace_b200.context.shard_pixel_list) that the gloo test exchanges with
    from vk_raytrace_b200.context import shard_pixel_list
    ramp = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
    ramp[..., 0] = torch.arange(H * W, device="cuda:0", dtype=torch.float32).reshape(H, W)
    gpu.output_create(W, H, ramp.data_ptr())
    for rank in range(n):
        gpu.pack_shard(32, rank, n, parts[rank].data_ptr())
        gpu.sync()
        assert np.array_equal(parts[rank][:, 0].cpu().numpy().astype(np.uint32), shard_pixel_list(W, H, 32, rank, n))
    gpu.output_create(W, H)


def test_mirror_stores_shards_into_a_second_image(vk, gpu):
    """vkrt_output_set_mirror (the exchange fused into the accumulate kernel; across processes the target is a vkrt_peer_open
    mapping of rank 0's image, bench.py --gpus N): two contexts render the two tile shards, the second mirrors into the first
    one's image, and that image is, bit for bit, the unsharded render.  Also vkrt_output_fence / _export argument handling."""
    import torch
    from vk_raytrace_b200 import Context, structs as S
    W, H = 200, 120
    sc = small_sponza()
    vk.upload_scene(gpu, sc, W, H)
    st = S.default_rtx_state(W, H, maxDepth=5)
    gpu.output_clear()
    gpu.render(st, 6)
    want = gpu.output_read()
    other = Context(0)
    other.precise_default = True                             # like the `gpu` fixture: the IEEE-exact shade kernel
    try:
        vk.upload_scene(other, sc, W, H)
        gpu.output_clear()
        gpu.sync()                                           # the other context's stores must not overtake the clear
        other.set_mirror(gpu.output_device_ptr())
        for f0, n in ((0, 4), (4, 2)):                       # two calls: the running mean continues through the mirror too
            for rank, c in ((0, gpu), (1, other)):
                o = S.RenderOpts()
                o.shard_rank, o.shard_count, o.tile_size, o.frames_in_flight = rank, 2, 32, 2
                c.render(S.default_rtx_state(W, H, maxDepth=5, frame=f0), n, o, sync=False)
                c.output_fence()
        other.sync()
        gpu.sync()
        assert np.array_equal(gpu.output_read(), want)
        mine = other.output_read()                           # the mirroring context keeps its own shard locally
        own = mine[..., 3] == 1
        assert 0.3 * W * H < own.sum() < 0.7 * W * H and np.array_equal(mine[own], want[own])
        other.set_mirror(None)
        other.render(S.default_rtx_state(W, H, maxDepth=5, frame=6), 1)
        assert np.array_equal(gpu.output_read(), want)       # switched off: the first image is no longer touched
        assert len(gpu.output_export()) == 64                # library-owned image: exportable (opening it needs another process)
        ext = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
        gpu.output_create(W, H, ext.data_ptr())
        with pytest.raises(Exception):
            gpu.output_export()                              # caller-owned buffers are the caller's to export
        gpu.output_create(W, H)
    finally:
        other.close()


def test_forced_tree_width_gives_identical_hits(vk, ob, gpu):
    """VKRT_BUILD_FORCE_BVH2 / _BVH4 (what bench.py's HBM roofline record uses) select the tree per scene; results do not change."""
    sc = small_sponza()
    rays = random_rays(20000, 3, 14.0, centre=(0, 3, 0))
    hits = []
    for flag, width in ((0x100, 2), (0x200, 4)):
        vk.upload_scene(gpu, sc, 16, 16, build_mode=1 | flag)
        assert gpu.stats().bvh_width == width
        hits.append(gpu.trace_rays(rays, True).view(np.uint32))
    assert np.array_equal(hits[0], hits[1])
    o = ob.Oracle()
    vk.upload_scene(o, sc, 16, 16)
    assert np.array_equal(hits[0], o.trace_rays(rays, True).view(np.uint32))


def test_scene_file_through_host_library_and_bench(vk, gpu, tmp_path):
    """`bench.py --scene-file` / vkrt_host_load_gltf: a glTF file fed through libvkrt_host.so gives the geometry of the in-memory
    upload (same hit distances and barycentrics for 20 k rays; triangle numbering follows the importer's mesh order), and the bench
    prints a valid line for it."""
    import json
    import subprocess
    import sys
    from vk_raytrace_b200.gltf_export import export_gltf
    sc = small_sponza()
    gltf, env = export_gltf(sc, str(tmp_path), "s")
    rays = random_rays(20000, 5, 14.0, centre=(0, 3, 0))
    vk.upload_scene(gpu, sc, 32, 32, build_mode=4)
    want = gpu.trace_rays(rays, True)
    n_tri = gpu.stats().num_triangles
    info = gpu.load_gltf(gltf, 4)
    integral, ew, eh = gpu.load_environment(env)
    assert info.triangles == n_tri == gpu.stats().num_triangles and info.lights == len(sc.lights) and info.camera_from_file == 1
    assert (ew, eh) == (sc.env.shape[1], sc.env.shape[0]) and integral > 0
    got = gpu.trace_rays(rays, True)
    assert np.array_equal(got[:, :3].view(np.uint32), want[:, :3].view(np.uint32))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--scene-file", gltf, "--env-file", env, "--width", "320", "--height", "200", "--steps", "4",
                          "--warmup", "3", "--no-hbm-record"], capture_output=True, text=True, cwd=root, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"]["workload"].startswith("file:") and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["config"]["triangles"] == n_tri and line["rays"] >= 4 * 320 * 200
