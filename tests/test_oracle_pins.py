"""Pins for the oracle that do not originate in this repository's own renderer (CPU, no GPU): the reference has no golden
vectors and cannot run here, so -- besides the RNG / packing known answers of test_oracle.py -- the restatement is checked
against facts that follow from the reference's *documented algorithm* by independent mathematics:

  * camera            glm::lookAt / perspectiveRH_ZO / inverse (src/scene.cpp:629-640) in float64 numpy
  * integral of pdf   closed form for GGX half-vector sampling at normal incidence (pbr_gltf.glsl:365-434, pbr_disney.glsl:524-599)
  * Sample vs Eval    E[f cos / pdf] over PbrSample == quadrature of PbrEval (where the reference itself is self-consistent)
  * furnace           a convex surface under a constant environment: rendered radiance == the value the integrator's structure
                      (NEE with power heuristic + un-weighted environment on a miss, pathtrace.glsl:97-188,204-231) predicts
                      from the BSDF by quadrature
  * BLEND alpha       stochastic opacity passes a ray with probability 1 - alpha (traceray_rq.glsl:87-99)
"""
import ctypes as C

import numpy as np
import pytest

from helpers import blend_probe_scene, branch_scene, random_rays


def _state(V, albedo=(1, 1, 1), metallic=0.0, roughness=0.5, transmission=0.0, ior=1.5, eta=1 / 1.5, clearcoat=0.0, ccr=0.1):
    a = np.zeros(32, np.float32)
    a[0:3] = V
    a[3:6] = (0, 0, 1)
    a[9:12] = (1, 0, 0)
    a[12:15] = (0, 1, 0)
    a[15:18] = albedo
    a[18], a[19], a[20], a[21], a[22], a[23], a[24] = metallic, roughness, transmission, ior, eta, clearcoat, ccr
    a[27:30] = (1, 1, 1)
    return a


def _eval(lib, mode, st, dirs):
    out = np.zeros((len(dirs), 4), np.float32)
    o = np.zeros(4, np.float32)
    st = st.copy()
    for i, l in enumerate(dirs):
        st[6:9] = l
        lib.oracle_bsdf_eval(mode, st.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        out[i] = o
    return out


def _hemisphere(nt, nphi):
    z = (np.arange(nt) + 0.5) / nt
    ph = (np.arange(nphi) + 0.5) / nphi * 2 * np.pi
    Z, P = np.meshgrid(z, ph, indexing="ij")
    r = np.sqrt(1 - Z * Z)
    d = np.stack([r * np.cos(P), r * np.sin(P), Z], -1).reshape(-1, 3).astype(np.float32)
    return d, 2 * np.pi / (nt * nphi)


def _sample_albedo(lib, mode, st, n, seed0):
    o = np.zeros(8, np.float32)
    acc, sq = 0.0, 0.0
    for i in range(n):
        lib.oracle_bsdf_sample(mode, st.ctypes.data_as(C.c_void_p), (seed0 * 7919 + i * 2654435761) & 0xffffffff, o.ctypes.data_as(C.c_void_p))
        if o[6] > 0:
            v = float(o[3]) * abs(float(o[2])) / float(o[6])
            acc += v
            sq += v * v
    mean = acc / n
    return mean, np.sqrt(max(sq / n - mean * mean, 0.0) / n)


def test_camera_against_glm_semantics(ob, vk):
    """oracle make_camera (float, glm's algorithms) against float64 numpy; the product's vkrt_make_camera against both."""
    rng = np.random.default_rng(3)
    for _ in range(8):
        eye, center = rng.uniform(-5, 5, 3), rng.uniform(-1, 1, 3)
        up = np.array([0.1 * rng.normal(), 1.0, 0.1 * rng.normal()])
        fov, aspect = float(rng.uniform(20, 90)), float(rng.uniform(0.5, 2.5))
        f = (center - eye) / np.linalg.norm(center - eye)
        s = np.cross(f, up)
        s /= np.linalg.norm(s)
        u = np.cross(s, f)
        view = np.eye(4)
        view[0, :3], view[1, :3], view[2, :3] = s, u, -f
        view[:3, 3] = [-s @ eye, -u @ eye, f @ eye]
        zn, zf, th = 0.001, 100000.0, np.tan(np.radians(fov) / 2)
        proj = np.zeros((4, 4))
        proj[0, 0], proj[1, 1], proj[2, 2], proj[3, 2], proj[2, 3] = 1 / (aspect * th), -1 / th, zf / (zn - zf), -1.0, -(zf * zn) / (zf - zn)
        want_v, want_p = np.linalg.inv(view).T.reshape(16), np.linalg.inv(proj).T.reshape(16)   # column-major
        for cam in (ob.make_camera(eye, center, up, fov, aspect, 0.05, 2), vk.context.make_camera(eye, center, up, fov, aspect, 0.05, 2)):
            assert np.allclose(np.array(cam.viewInverse), want_v, rtol=2e-6, atol=2e-6)
            assert np.allclose(np.array(cam.projInverse), want_p, rtol=2e-6, atol=1e-7 * np.abs(want_p).max())
            assert abs(cam.focalDist - np.linalg.norm(center - eye)) < 1e-5 and cam.nbLights == 2 and abs(cam.aperture - 0.05) < 1e-8


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("metallic,roughness", [(0.0, 0.5), (1.0, 0.3), (0.0, 1.0), (0.4, 0.15)])
def test_pdf_integral_closed_form_at_normal_incidence(ob, mode, metallic, roughness):
    """Both models mix a cosine lobe (weight d = 0.5 (1 - metallic)) with a GGX half-vector lobe of alpha = roughness.  With V = N the
    reflected direction leaves the upper hemisphere exactly when theta_h > 45 deg, and the GGX half-vector CDF
    cos^2(theta_h) = (1 - xi) / (1 + (alpha^2 - 1) xi) gives P(theta_h < 45 deg) = 1 / (1 + alpha^2).  Hence
    int pdf dw = d + (1 - d) / (1 + alpha^2) -- not 1: the reference drops below-horizon samples instead of renormalising."""
    lib = ob.load()
    dirs, w = _hemisphere(300, 600)
    e = _eval(lib, mode, _state((0, 0, 1), metallic=metallic, roughness=roughness), dirs)
    d = 0.5 * (1 - metallic)
    want = d + (1 - d) / (1 + roughness ** 2)
    got = float((e[:, 3].astype(np.float64) * w).sum())
    assert abs(got - want) < 4e-3, (got, want)
    assert (e[:, 3] >= 0).all() and np.isfinite(e).all()


@pytest.mark.parametrize("theta", [0.0, 0.6, 1.2])
@pytest.mark.parametrize("metallic,roughness,albedo", [(0.0, 0.5, (1, 1, 1)), (1.0, 0.3, (1, 1, 1)), (0.0, 1.0, (0.8, 0.8, 0.8)), (0.5, 0.25, (0.9, 0.9, 0.9))])
def test_gltf_sample_matches_eval(ob, theta, metallic, roughness, albedo):
    """Directional albedo two ways: Monte-Carlo over PbrSample's own (f, pdf) and midpoint quadrature of PbrEval.  PbrSample returns
    one lobe's f over that lobe's pdf x selection probability, so its expectation is the integral of the summed lobes PbrEval
    returns (pbr_gltf.glsl:439-554 vs :365-434; no clearcoat: there the reference picks roughness and lobe with two independent
    draws and is not self-consistent)."""
    lib = ob.load()
    V = np.array([np.sin(theta), 0, np.cos(theta)], np.float32)
    st = _state(V, albedo=albedo, metallic=metallic, roughness=roughness)
    dirs, w = _hemisphere(200, 400)
    e = _eval(lib, 1, st, dirs)
    quad = float((e[:, 0].astype(np.float64) * dirs[:, 2] * w).sum())
    mc, err = _sample_albedo(lib, 1, st, 30000, 5)
    assert abs(mc - quad) < max(4 * err, 0.004) + 0.004 * quad, (mc, quad, err)
    assert 0.3 < quad < 1.2


def test_disney_specular_inconsistency_is_the_references(ob):
    """pbr_disney.glsl:85-94: ImportanceSampleGTR2_aniso returns an un-normalised half vector, which reflect() and GTR2_aniso then
    use as if it were a unit vector -- so for the Disney model Sample and Eval disagree *in the reference* (the metallic albedo by
    Monte-Carlo exceeds the quadrature).  The restatement must keep that: a 'fixed' oracle would no longer be the reference."""
    lib = ob.load()
    st = _state((0, 0, 1), metallic=1.0, roughness=0.3)
    dirs, w = _hemisphere(200, 400)
    quad = float((_eval(lib, 0, st, dirs)[:, 0].astype(np.float64) * dirs[:, 2] * w).sum())
    mc, err = _sample_albedo(lib, 0, st, 30000, 9)
    assert mc - quad > 0.05 and err < 0.01, (mc, quad, err)
    # the glTF model uses a normalised half vector for the same material and agrees with its quadrature
    quad1 = float((_eval(lib, 1, st, dirs)[:, 0].astype(np.float64) * dirs[:, 2] * w).sum())
    mc1, err1 = _sample_albedo(lib, 1, st, 30000, 9)
    assert abs(mc1 - quad1) < 4 * err1 + 0.004


@pytest.mark.parametrize("pbr,metallic,roughness", [(1, 0.0, 0.6), (1, 1.0, 0.4), (0, 0.0, 1.0)])
def test_furnace_radiance_follows_from_the_integrator_structure(ob, vk, pbr, metallic, roughness):
    """White-furnace with the reference's integrator, not an idealised one.  A large double-sided quad under a constant environment
    E = 1: a camera path hits the quad once; then (pathtrace.glsl:199-340)
      * DirectLight samples the environment (pdf p_l = 1/4pi for a constant map) and adds  w f cos E / p_l  with the power heuristic
        w = p_l^2 / (p_l^2 + p_b^2), p_b = Eval's pdf;
      * Sample picks the next direction; a miss adds E x throughput *without* MIS weight (pathtrace.glsl:227), RR is unbiased.
    So radiance = E[f cos / pdf]_Sample + int w f cos dw.  Both terms are computed here from the BSDF hooks alone (Monte-Carlo over
    Sample with other seeds, quadrature over Eval) and compared with the rendered mean.  The result is NOT 1 -- the reference
    double-counts the environment (and says so for its BSDF at pbr_disney.glsl:566) -- which is exactly why this pins the oracle's
    integrator: an implementation with textbook MIS would give a different number."""
    from vk_raytrace_b200 import scenes, structs as S
    from vk_raytrace_b200.context import SceneDesc
    lib = ob.load()
    theta = 0.5
    sc = SceneDesc(name="furnace")
    sc.meshes = [scenes.quad((-50, 0, 50), (50, 0, 50), (50, 0, -50), (-50, 0, -50), 0)]        # faces +Y
    sc.nodes = [(np.eye(4, dtype=np.float32), 0)]
    m = S.default_material(1)
    m["pbrMetallicFactor"], m["pbrRoughnessFactor"], m["doubleSided"] = metallic, roughness, 1
    sc.materials = m
    sc.env = scenes.constant_environment(64, 32, 1.0)
    eye = np.array([np.sin(theta), np.cos(theta), 0.0]) * 4.0
    sc.camera = dict(eye=tuple(eye), center=(0, 0, 0), up=(0, 0, -1), fov=1.5)
    W = H = 24
    o = ob.Oracle()
    vk.upload_scene(o, sc, W, H, packer=ob)
    frames = 96
    st = S.default_rtx_state(W, H, maxDepth=3, pbrMode=pbr, fireflyClampThreshold=1e9)
    o.render(st, frames)
    img = o.output_read()[..., 0].astype(np.float64)
    got = img.mean()
    # prediction: the shading frame of the oracle_bsdf hooks has N = +Z; V makes the angle theta with it
    stt = _state((np.sin(theta), 0, np.cos(theta)), metallic=metallic, roughness=roughness)
    dirs, w = _hemisphere(300, 600)
    e = _eval(lib, pbr, stt, dirs).astype(np.float64)
    pl = 1.0 / (4 * np.pi)
    wmis = pl * pl / (pl * pl + e[:, 3] ** 2)
    nee = float((wmis * e[:, 0] * dirs[:, 2] * w).sum())
    mc, err = _sample_albedo(lib, pbr, stt, 40000, 21)
    want = mc + nee
    # per-sample spread of the render is that of the BSDF estimator: bound both noises
    noise = 4 * err + 4 * img.std() / np.sqrt(W * H)
    assert abs(got - want) < noise + 0.01 * want, (got, want, mc, nee, noise)
    # the double-counted environment: a textbook integrator would return the directional albedo `mc` alone
    assert nee > 0.02 * mc and abs(got - mc) > noise


def test_blend_alpha_passes_rays_with_probability_one_minus_alpha(ob, vk):
    """traceray_rq.glsl:87-99: for BLEND (any mode but MASK) a candidate is ignored when rand > opacity, opacity = base alpha."""
    for alpha in (0.3, 0.75):
        o = ob.Oracle()
        vk.upload_scene(o, blend_probe_scene(alpha), 8, 8, packer=ob)
        rng = np.random.default_rng(5)
        n = 40000
        rays = np.zeros((n, 8), np.float32)
        rays[:, 0:2] = rng.uniform(-0.9, 0.9, (n, 2))
        rays[:, 2] = 1.0
        rays[:, 3] = 1e32
        rays[:, 6] = -1.0
        rays[:, 7] = rng.integers(0, 2 ** 32, n, dtype=np.uint32).view(np.float32)
        for closest in (True, False):
            hit = o.trace_rays(rays, closest)[:, 0] < 1e31
            assert abs(hit.mean() - alpha) < 4 * np.sqrt(alpha * (1 - alpha) / n) + 1e-3, (alpha, hit.mean())


@pytest.mark.parametrize("kind", ["blend", "dof", "directional", "unlit", "hdr0"])
def test_branch_scenes_reach_their_branch(ob, vk, kind):
    """The Cornell variants the GPU parity tests use (tests/test_gpu_parity.py::test_reference_branches) do change the image in the
    way the branch implies -- i.e. the branch is really taken by the oracle."""
    from vk_raytrace_b200 import scenes, structs as S
    W = H = 160 if kind == "dof" else 40      # the circle of confusion of aperture 0.1 is ~1/40 of the image width
    sc, kw = branch_scene(kind)
    base = scenes.cornell_box(with_light=(kind in ("hdr0", "blend")))
    imgs = []
    for s, k in ((sc, kw), (base, {})):
        o = ob.Oracle()
        vk.upload_scene(o, s, W, H, packer=ob)
        if kind == "dof":     # base-colour debug view: no shading noise, so the lens blur is the only difference
            k = dict(k, debugging_mode=1, maxDepth=1)
        st = S.default_rtx_state(W, H, maxDepth=k.pop("maxDepth", 4), **k)
        o.render(st, 32 if kind == "dof" else 6)
        imgs.append(o.output_read()[..., :3].astype(np.float64))
    a, b = imgs
    assert np.isfinite(a).all() and a.mean() > 0.01
    diff = np.abs(a - b).mean() / b.mean()
    assert diff > 0.01, (kind, diff)
    if kind == "unlit":
        # an unlit surface returns its albedo x throughput and ends the path: the red wall's first-hit pixels are exactly its colour
        assert np.isclose(a[H // 2, 1], (0.65, 0.05, 0.05), atol=1e-6).all()
    if kind == "dof":
        # the thin lens blurs: fewer hard edges than the pinhole image at equal sample count
        # (sum of squared differences: blurring an edge over k pixels divides it by ~k; the sum of absolute ones is conserved)
        gx = lambda im: (np.diff(im, axis=1) ** 2).sum()
        assert gx(a) < 0.6 * gx(b)


def _np_post_chain(hdr, tm, auto):
    """post.frag:98-147 + tonemapping.glsl:48-65 written again in float64 numpy from the GLSL semantics (mat3 constructors are
    column-major, so `RGB2XYZ * RGB` has XYZ.y = 0.3575761 R + 0.7151522 G + 0.1191920 B -- post.frag:58; fragment centres at
    (x + 0.5) / W; `textureLod(inImage, vec2(0.5), 20)` of a power-of-two image = the mean of all texels).  No dither."""
    H, W = hdr.shape[:2]
    rgb = hdr[..., :3].astype(np.float64)
    if auto:
        avg = rgb.reshape(-1, 3).mean(axis=0)
        log_avg = avg @ np.array([0.2126, 0.7152, 0.0722])
        y = 0.3575761 * rgb[..., 0] + 0.7151522 * rgb[..., 1] + 0.1191920 * rgb[..., 2]
        Y = (tm.key / log_avg) * y
        Yd = (Y * (1.0 + Y / (tm.Ywhite * tm.Ywhite))) / (1.0 + Y)
        rgb = rgb / y[..., None] * Yd[..., None]

    def u2(c):
        A, B, Cc, D, E, F = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
        return ((c * (A * c + Cc * B) + D * E) / (c * (A * c + B) + D * F)) - E / F

    c = u2(rgb * tm.avgLum * 2.0) * (1.0 / u2(np.float64(11.2)))
    c = np.power(c, 1.0 / 2.2)
    c = np.clip(0.5 + (c - 0.5) * tm.contrast, 0.0, 1.0)
    c = np.power(c, 1.0 / tm.brightness)
    i = c @ np.array([0.299, 0.587, 0.114])
    c = i[..., None] + (c - i[..., None]) * tm.saturation
    u = ((np.arange(W) + 0.5) / W * tm.renderingRatio[0] - 0.5) * 2.0
    v = ((np.arange(H) + 0.5) / H * tm.renderingRatio[1] - 0.5) * 2.0
    c = c * (1.0 - (u[None, :] ** 2 + v[:, None] ** 2) * tm.vignette)[..., None]
    return c


@pytest.mark.parametrize("auto", [0, 1])
def test_post_chain_against_an_independent_restatement(ob, auto):
    """Tonemap, contrast, brightness, saturation, vignette and the global auto-exposure operator with every knob off its default."""
    from vk_raytrace_b200 import structs as S
    rng = np.random.default_rng(5)
    H, W = 32, 64
    hdr = np.ones((H, W, 4), dtype=np.float32)
    hdr[..., :3] = (rng.random((H, W, 3)) ** 3 * 12.0 + 0.01).astype(np.float32)      # up to beyond the white point
    o = ob.Oracle()
    o.output_create(W, H)
    o.output_write(hdr)
    tm = S.default_tonemapper(brightness=1.3, contrast=1.2, saturation=0.8, vignette=0.4, avgLum=0.7, dither=0, autoExposure=auto, Ywhite=0.8, key=0.4)
    tm.renderingRatio[0], tm.renderingRatio[1] = 1.0, 0.75
    got = o.tonemap_float(tm)[..., :3].astype(np.float64)
    want = _np_post_chain(hdr, tm, auto)
    assert np.abs(got - want).max() < (2e-4 if auto else 2e-5), np.abs(got - want).max()
    assert np.abs(want).max() > 0.5 and want.std() > 0.1


def _np_bilinear(tex, uv, wrap):
    """VK_FILTER_LINEAR at lod 0 written from the Vulkan texel-addressing rules in float64: unnormalised coordinate u*W - 0.5, the
    two nearest texels weighted by its fraction; REPEAT (10497) = modulo, CLAMP_TO_EDGE (33071) = clamp, MIRRORED_REPEAT (33648).
    RGBA8 UNORM texels = k / 255.  (Hardware rounds the weights to 8 fractional bits; the oracle and the kernels use full fp32
    weights, DESIGN.md section 4 -- this pin is about addressing, wrap and orientation.)"""
    H, W = tex.shape[:2]

    def wrap_i(i, n):
        if wrap == 10497:
            return np.mod(i, n)
        if wrap == 33648:
            m = np.mod(i, 2 * n)
            return np.where(m >= n, 2 * n - 1 - m, m)
        return np.clip(i, 0, n - 1)

    x, y = uv[..., 0].astype(np.float64) * W - 0.5, uv[..., 1].astype(np.float64) * H - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
    t = tex.astype(np.float64) / 255.0
    c00, c10 = t[wrap_i(y0, H), wrap_i(x0, W)], t[wrap_i(y0, H), wrap_i(x0 + 1, W)]
    c01, c11 = t[wrap_i(y0 + 1, H), wrap_i(x0, W)], t[wrap_i(y0 + 1, H), wrap_i(x0 + 1, W)]
    return (c00 * (1 - fx) + c10 * fx) * (1 - fy) + (c01 * (1 - fx) + c11 * fx) * fy


@pytest.mark.parametrize("wrap", [10497, 33071, 33648])
def test_texture_addressing_against_an_independent_restatement(ob, vk, wrap):
    """A textured quad seen through debug views TEXCOORD and BASECOLOR (frame 0: rays through pixel centres).  The interpolated uv
    must be what a float64 ray / plane intersection gives, and the base colour must be the sRGB -> linear (pow 2.2,
    gltf_material.glsl:79-83) of a bilinear fetch at that uv times the factor -- for every wrap mode, with uv running past [0, 1]."""
    from vk_raytrace_b200 import SceneDesc, TextureDesc, scenes, structs as S
    rng = np.random.default_rng(3)
    tex = (rng.random((8, 16, 4)) * 255).astype(np.uint8)
    s = SceneDesc(name="tex_probe")
    s.meshes = [scenes.quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0), 0, uv_scale=2.5)]
    s.nodes = [(np.eye(4, dtype=np.float32), 0)]
    m = S.default_material(1)
    m["pbrBaseColorFactor"] = (0.9, 0.8, 0.7, 1.0)
    m["pbrBaseColorTexture"] = 0
    s.materials = m
    s.textures = [TextureDesc(tex, wrap_s=wrap, wrap_t=wrap)]
    s.env = scenes.constant_environment(8, 4, 1.0)
    s.camera = dict(eye=(0.3, 0.2, 3.0), center=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=30.0)
    W, H = 48, 40
    o = ob.Oracle()
    vk.upload_scene(o, s, W, H, packer=ob)
    out = {}
    for name, mode in (("uv", 7), ("base", 1)):
        o.render(S.default_rtx_state(W, H, maxDepth=2, debugging_mode=mode, fireflyClampThreshold=1e9), 1)   # frame 0 overwrites the image; debug views go through the firefly clamp too (pathtrace.glsl:380-384)
        out[name] = o.output_read()[..., :3].astype(np.float64)
    # independent uv: pinhole camera of glm::lookAt / perspectiveRH_ZO (fov is vertical; image row 0 is the top, proj[1][1] *= -1)
    eye = np.array(s.camera["eye"], dtype=np.float64)
    f = -eye / np.linalg.norm(eye)
    r = np.cross(f, np.array([0.0, 1.0, 0.0]))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    th = np.tan(np.radians(30.0) / 2)
    px = ((np.arange(W) + 0.5) / W * 2 - 1) * th * (W / H)
    py = -((np.arange(H) + 0.5) / H * 2 - 1) * th
    d = f[None, None, :] + px[None, :, None] * r[None, None, :] + py[:, None, None] * u[None, None, :]
    t = -eye[2] / d[..., 2]
    hit = eye[None, None, :] + t[..., None] * d
    inside = (np.abs(hit[..., 0]) < 0.98) & (np.abs(hit[..., 1]) < 0.98)
    assert inside.sum() > 0.3 * W * H
    uv = np.stack([(hit[..., 0] + 1) / 2 * 2.5, (hit[..., 1] + 1) / 2 * 2.5], axis=-1)     # quad(): uv (0,0) at (-1,-1), (2.5,2.5) at (1,1)
    assert np.abs(out["uv"][..., :2] - uv)[inside].max() < 2e-5
    assert uv[inside].max() > 2.0 and uv[inside].min() < 0.5                                # past one repeat
    texel = _np_bilinear(tex, out["uv"][..., :2], wrap)[..., :3]
    want = np.power(texel, 2.2) * np.array([0.9, 0.8, 0.7])
    assert np.abs(out["base"] - want)[inside].max() < 2e-5
