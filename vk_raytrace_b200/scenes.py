"""Procedural stand-ins for the configs of BASELINE.json (no scene assets exist offline).

Every generator is a deterministic function of its arguments (integer seeds feed
numpy's PCG64) and returns a `SceneDesc` holding exactly what the reference's glTF
importer hands to Scene::create*Buffer: per-primMesh attribute arrays, nodes with world
matrices, GltfShadeMaterial records, KHR_lights_punctual lights, RGBA8 textures, an
RGBA32F environment map and a camera.

  cornell_box          C1  ~36 triangles, opaque, single sided
  helmet_standin       C2  one displaced icosphere with 4 PBR textures
  sponza_standin       C3/C4  arcade of columns / arches / floor / alpha-masked drapes
  instanced_stress     C5  N instances of one mesh (10 M triangles at the default size)
"""
import math

import numpy as np

from . import structs as S
from .context import MeshDesc, SceneDesc, TextureDesc


# ------------------------------------------------------------------------------ helpers
def _normalize(v, axis=-1):
    n = np.linalg.norm(v, axis=axis, keepdims=True)
    return v / np.maximum(n, 1e-20)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def scale(x, y=None, z=None):
    y = x if y is None else y
    z = x if z is None else z
    return np.diag([x, y, z, 1.0]).astype(np.float32)


def rotate(axis, angle):
    axis = _normalize(np.asarray(axis, dtype=np.float64))
    x, y, z = axis
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    m = np.eye(4)
    m[:3, :3] = [[c + x * x * C, x * y * C - z * s, x * z * C + y * s], [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                 [z * x * C - y * s, z * y * C + x * s, c + z * z * C]]
    return m.astype(np.float32)


def finish_mesh(positions, indices, texcoords=None, material=0, smooth=True, colors=None):
    """Adds what the importer would generate when attributes are missing: normals from faces,
    tangents from the uv parametrisation (w = handedness), white vertex colours."""
    positions = np.asarray(positions, dtype=np.float32)
    indices = np.asarray(indices, dtype=np.uint32).reshape(-1)
    tri = indices.reshape(-1, 3)
    n = len(positions)
    if texcoords is None:
        texcoords = np.zeros((n, 2), dtype=np.float32)
    texcoords = np.asarray(texcoords, dtype=np.float32)
    p0, p1, p2 = positions[tri[:, 0]].astype(np.float64), positions[tri[:, 1]].astype(np.float64), positions[tri[:, 2]].astype(np.float64)
    fn = np.cross(p1 - p0, p2 - p0)
    if not smooth:
        # unshare vertices so that every face has its own normal
        positions = positions[indices]
        texcoords = texcoords[indices]
        if colors is not None:
            colors = np.asarray(colors, dtype=np.float32)[indices]
        n = len(positions)
        indices = np.arange(n, dtype=np.uint32)
        tri = indices.reshape(-1, 3)
        normals = np.repeat(_normalize(fn), 3, axis=0)
    else:
        normals = np.zeros((n, 3))
        for k in range(3):
            np.add.at(normals, tri[:, k], fn)
        normals = _normalize(normals)
    # tangents
    uv = texcoords.astype(np.float64)
    p = positions.astype(np.float64)
    e1, e2 = p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]]
    d1, d2 = uv[tri[:, 1]] - uv[tri[:, 0]], uv[tri[:, 2]] - uv[tri[:, 0]]
    det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
    ok = np.abs(det) > 1e-12
    r = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)[:, None]
    ft = (e1 * d2[:, 1:2] - e2 * d1[:, 1:2]) * r
    fb = (e2 * d1[:, 0:1] - e1 * d2[:, 0:1]) * r
    tan = np.zeros((n, 3))
    bit = np.zeros((n, 3))
    for k in range(3):
        np.add.at(tan, tri[:, k], ft)
        np.add.at(bit, tri[:, k], fb)
    # Gram-Schmidt, fall back to any perpendicular where the uv map is degenerate
    tan = tan - normals * np.sum(normals * tan, axis=1, keepdims=True)
    bad = np.linalg.norm(tan, axis=1) < 1e-10
    alt = np.where(np.abs(normals[:, 0:1]) < 0.9, [[1.0, 0.0, 0.0]], [[0.0, 1.0, 0.0]])
    alt = alt - normals * np.sum(normals * alt, axis=1, keepdims=True)
    tan = np.where(bad[:, None], alt, tan)
    tan = _normalize(tan)
    hand = np.where(np.sum(np.cross(normals, tan) * bit, axis=1) < 0.0, -1.0, 1.0)
    tangents = np.concatenate([tan, hand[:, None]], axis=1)
    if colors is None:
        colors = np.ones((n, 4), dtype=np.float32)
    return MeshDesc(positions.astype(np.float32), normals.astype(np.float32), tangents.astype(np.float32), texcoords.astype(np.float32),
                    np.asarray(colors, dtype=np.float32), indices.astype(np.uint32), material)


def quad(p0, p1, p2, p3, material=0, uv_scale=1.0):
    """CCW quad p0 p1 p2 p3 (front face towards the side from which the order is counter-clockwise)."""
    pos = np.array([p0, p1, p2, p3], dtype=np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32) * uv_scale
    return finish_mesh(pos, [0, 1, 2, 0, 2, 3], uv, material, smooth=False)


def box_mesh(lo, hi, material=0):
    """Axis-aligned box, outward-facing CCW faces, 12 triangles."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], dtype=np.float32)
    faces = [(4, 5, 6, 7), (1, 0, 3, 2), (0, 4, 7, 3), (5, 1, 2, 6), (3, 7, 6, 2), (0, 1, 5, 4)]
    pos, uv, idx = [], [], []
    for f in faces:
        b = len(pos)
        pos += [v[i] for i in f]
        uv += [[0, 0], [1, 0], [1, 1], [0, 1]]
        idx += [b, b + 1, b + 2, b, b + 2, b + 3]
    return finish_mesh(np.array(pos), idx, np.array(uv, dtype=np.float32), material, smooth=False)


def grid_mesh(nx, nz, size_x, size_z, material=0, uv_scale=1.0, height=None):
    """XZ grid facing +Y, (nx*nz*2) triangles; optional height field callback(x, z)."""
    xs = np.linspace(-size_x / 2, size_x / 2, nx + 1)
    zs = np.linspace(-size_z / 2, size_z / 2, nz + 1)
    X, Z = np.meshgrid(xs, zs, indexing="xy")
    Y = np.zeros_like(X) if height is None else height(X, Z)
    pos = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    uv = np.stack([(X / size_x + 0.5) * uv_scale, (Z / size_z + 0.5) * uv_scale], axis=-1).reshape(-1, 2)
    i = np.arange(nx)[None, :] + (nx + 1) * np.arange(nz)[:, None]
    a, b, c, d = i, i + 1, i + nx + 2, i + nx + 1
    idx = np.stack([a, d, c, a, c, b], axis=-1).reshape(-1)   # CCW seen from +Y
    return finish_mesh(pos, idx, uv, material, smooth=True)


def uv_sphere(nu, nv, radius=1.0, material=0, displace=None):
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v, indexing="xy")
    phi, theta = U * 2 * np.pi, V * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], axis=-1)
    r = radius if displace is None else radius * (1.0 + displace(d))[..., None]
    pos = (d * r).reshape(-1, 3)
    uv = np.stack([U, V], axis=-1).reshape(-1, 2)
    i = np.arange(nu)[None, :] + (nu + 1) * np.arange(nv)[:, None]
    a, b, c, dd = i, i + 1, i + nu + 2, i + nu + 1
    idx = np.stack([a, b, c, a, c, dd], axis=-1).reshape(-1)  # CCW seen from outside
    return finish_mesh(pos, idx, uv, material, smooth=True)


def cylinder(nseg, nrings, radius, height, material=0, profile=None):
    u = np.linspace(0.0, 1.0, nseg + 1)
    v = np.linspace(0.0, 1.0, nrings + 1)
    U, V = np.meshgrid(u, v, indexing="xy")
    phi = U * 2 * np.pi
    r = radius * (np.ones_like(V) if profile is None else profile(V, phi))
    pos = np.stack([r * np.cos(phi), V * height, r * np.sin(phi)], axis=-1).reshape(-1, 3)
    uv = np.stack([U * 2.0, V * 4.0], axis=-1).reshape(-1, 2)
    i = np.arange(nseg)[None, :] + (nseg + 1) * np.arange(nrings)[:, None]
    a, b, c, d = i, i + 1, i + nseg + 2, i + nseg + 1
    idx = np.stack([a, c, b, a, d, c], axis=-1).reshape(-1)   # CCW seen from outside
    return finish_mesh(pos, idx, uv, material, smooth=True)


def icosphere(level):
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1],
                  [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v = _normalize(v)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * (len(v) + 1) + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        first = np.zeros(len(uniq), dtype=np.int64)
        first[inv] = np.arange(len(key))
        mid = _normalize((v[es[first, 0]] + v[es[first, 1]]) * 0.5)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        m = (base + inv).reshape(3, -1).T  # midpoints of edges 01, 12, 20 per face
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        ab, bc, ca = m[:, 0], m[:, 1], m[:, 2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], axis=0)
    return v, f


def value_noise(shape, cells, rng):
    """Smooth tileable value noise in [0,1], shape (h,w)."""
    h, w = shape
    g = rng.random((cells + 1, cells + 1))
    g[-1, :] = g[0, :]
    g[:, -1] = g[:, 0]
    y = np.linspace(0, cells, h, endpoint=False)
    x = np.linspace(0, cells, w, endpoint=False)
    y0, x0 = np.floor(y).astype(int), np.floor(x).astype(int)
    fy, fx = y - y0, x - x0
    fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx)[None, :] + b * fx[None, :]) * (1 - fy)[:, None] + (c * (1 - fx)[None, :] + d * fx[None, :]) * fy[:, None]


def fbm(shape, rng, octaves=4, base=4):
    out = np.zeros(shape)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        out += amp * value_noise(shape, base << o, rng)
        tot += amp
        amp *= 0.5
    return out / tot


def to_rgba8(rgb, alpha=None):
    rgb = np.clip(np.asarray(rgb), 0.0, 1.0)
    h, w = rgb.shape[:2]
    a = np.ones((h, w, 1)) if alpha is None else np.clip(alpha, 0.0, 1.0).reshape(h, w, 1)
    return np.round(np.concatenate([rgb, a], axis=-1) * 255.0).astype(np.uint8)


def sky_environment(width=512, height=256, seed=7, sun=(0.35, 0.28), sun_power=600.0):
    """Procedural lat-long HDR: horizon-to-zenith gradient, darker ground, soft clouds, a sun blob."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    U, V = np.meshgrid(u, v, indexing="xy")
    up = np.cos(V * np.pi)                      # +1 zenith .. -1 nadir
    sky = np.stack([0.25 + 0.45 * (1 - np.clip(up, 0, 1)) ** 2, 0.45 + 0.35 * (1 - np.clip(up, 0, 1)) ** 2, 0.9 - 0.15 * (1 - np.clip(up, 0, 1))], axis=-1)
    ground = np.stack([0.18 * np.ones_like(up), 0.16 * np.ones_like(up), 0.13 * np.ones_like(up)], axis=-1)
    t = np.clip(up * 12.0 + 0.5, 0.0, 1.0)[..., None]
    img = sky * t + ground * (1 - t)
    clouds = np.clip(fbm((height, width), rng, 4, 4) * 1.8 - 0.7, 0.0, 1.0)[..., None] * np.clip(up, 0, 1)[..., None]
    img = img * (1 - 0.6 * clouds) + 0.9 * clouds
    du = np.minimum(np.abs(U - sun[0]), 1 - np.abs(U - sun[0])) * 2 * np.pi * np.sin(V * np.pi)
    dv = (V - sun[1]) * np.pi
    ang2 = du * du + dv * dv
    img = img + (np.exp(-ang2 / (2 * 0.02 ** 2)) * sun_power + np.exp(-ang2 / (2 * 0.12 ** 2)) * 4.0)[..., None] * np.array([1.0, 0.93, 0.82])
    out = np.concatenate([img, np.ones((height, width, 1))], axis=-1)
    return np.ascontiguousarray(out, dtype=np.float32)


def constant_environment(width=16, height=8, value=1.0):
    return np.full((height, width, 4), value, dtype=np.float32)


def point_light(position, color=(1, 1, 1), intensity=10.0, rng=0.0):
    l = np.zeros(1, dtype=S.Light)
    l["position"] = position
    l["direction"] = (0, 0, -1)
    l["color"] = color
    l["intensity"] = intensity
    l["range"] = rng
    l["innerConeCos"] = 1.0
    l["outerConeCos"] = math.cos(math.pi / 4)
    l["type"] = 1
    return l


# ------------------------------------------------------------------------------ C1
def cornell_box(with_light=False, env_value=1.0):
    """Cornell box: 5 walls + 2 boxes + emissive ceiling quad, 36 triangles, opaque, single sided.
    The open side faces +Z; the camera sits outside looking down -Z."""
    mats = S.default_material(6)
    mats["pbrMetallicFactor"] = 0.0
    mats["pbrRoughnessFactor"] = 1.0
    mats["pbrBaseColorFactor"][0] = (0.73, 0.73, 0.73, 1)       # white
    mats["pbrBaseColorFactor"][1] = (0.65, 0.05, 0.05, 1)       # red
    mats["pbrBaseColorFactor"][2] = (0.12, 0.45, 0.15, 1)       # green
    mats["pbrBaseColorFactor"][3] = (0.0, 0.0, 0.0, 1)          # light
    mats["emissiveFactor"][3] = (17.0, 12.0, 4.0)
    mats["pbrBaseColorFactor"][4] = (0.9, 0.85, 0.6, 1)         # brass-like short box
    mats["pbrMetallicFactor"][4] = 1.0
    mats["pbrRoughnessFactor"][4] = 0.25
    mats["pbrBaseColorFactor"][5] = (0.6, 0.7, 0.9, 1)          # glossy dielectric tall box
    mats["pbrRoughnessFactor"][5] = 0.35
    mats["clearcoatFactor"][5] = 0.6
    mats["clearcoatRoughness"][5] = 0.1
    s = SceneDesc(name="cornell")
    a = 1.0
    s.meshes = [
        quad((-a, -a, a), (a, -a, a), (a, -a, -a), (-a, -a, -a), 0),     # floor, faces +Y
        quad((-a, a, -a), (a, a, -a), (a, a, a), (-a, a, a), 0),         # ceiling, faces -Y
        quad((-a, -a, -a), (a, -a, -a), (a, a, -a), (-a, a, -a), 0),     # back wall, faces +Z
        quad((-a, -a, a), (-a, -a, -a), (-a, a, -a), (-a, a, a), 1),     # left (red), faces +X
        quad((a, -a, -a), (a, -a, a), (a, a, a), (a, a, -a), 2),         # right (green), faces -X
        quad((-0.3, a - 0.01, -0.3), (0.3, a - 0.01, -0.3), (0.3, a - 0.01, 0.3), (-0.3, a - 0.01, 0.3), 3),  # light, faces -Y
        box_mesh((-0.3, -0.3, -0.3), (0.3, 0.3, 0.3), 4),
        box_mesh((-0.3, -0.6, -0.3), (0.3, 0.6, 0.3), 5),
    ]
    s.nodes = [(np.eye(4, dtype=np.float32), i) for i in range(6)]
    s.nodes.append((translate(0.35, -0.7, 0.3) @ rotate((0, 1, 0), -0.3), 6))
    s.nodes.append((translate(-0.35, -0.4, -0.3) @ rotate((0, 1, 0), 0.35), 7))
    s.materials = mats
    s.lights = point_light((0.0, 0.5, 0.2), (1.0, 0.9, 0.8), 3.0) if with_light else None
    s.env = constant_environment(16, 8, env_value)
    s.camera = dict(eye=(0.0, 0.0, 3.6), center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=40.0)
    return s


# ------------------------------------------------------------------------------ C2
def helmet_standin(level=6, tex_size=2048, env_size=(2048, 1024), seed=1234):
    """Single PBR mesh stand-in for DamagedHelmet: displaced icosphere (20*4^level triangles) with
    base-colour, metallic-roughness, normal and emissive textures (PCG-seeded noise)."""
    rng = np.random.default_rng(seed)
    v, f = icosphere(level)
    # spherical uv; duplicate nothing (seam triangles stretch, which is fine for a stress asset)
    u = np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi) + 0.5
    w = np.arccos(np.clip(v[:, 1], -1, 1)) / np.pi
    bump = 0.06 * np.sin(v[:, 0] * 9.0) * np.sin(v[:, 1] * 7.0 + 1.3) * np.sin(v[:, 2] * 8.0 + 0.4) + 0.03 * np.sin(v[:, 1] * 23.0)
    pos = v * (1.0 + bump)[:, None]
    mesh = finish_mesh(pos, f.reshape(-1), np.stack([u * 2.0, w * 2.0], axis=1), 0, smooth=True)

    T = tex_size
    n1, n2, n3 = fbm((T, T), rng, 5, 4), fbm((T, T), rng, 4, 8), fbm((T, T), rng, 3, 16)
    base = np.stack([0.55 + 0.4 * n1, 0.35 + 0.4 * n2, 0.25 + 0.35 * n3], axis=-1)
    rough = 0.15 + 0.7 * n2
    metal = (n1 > 0.5).astype(np.float64) * 0.9 + 0.05
    mr = np.stack([np.ones_like(rough), rough, metal], axis=-1)
    gy, gx = np.gradient(n1 * 0.6 + n3 * 0.4)
    nrm = _normalize(np.stack([-gx * T * 0.02, -gy * T * 0.02, np.ones_like(gx)], axis=-1)) * 0.5 + 0.5
    emis = np.clip(n3 * 4.0 - 3.0, 0, 1)[..., None] * np.array([1.0, 0.45, 0.1])
    s = SceneDesc(name="helmet_standin")
    s.meshes = [mesh]
    s.nodes = [(rotate((0, 1, 0), 0.6) @ rotate((1, 0, 0), -0.2), 0)]
    s.textures = [TextureDesc(to_rgba8(base)), TextureDesc(to_rgba8(mr)), TextureDesc(to_rgba8(nrm)), TextureDesc(to_rgba8(emis))]
    m = S.default_material(1)
    m["pbrBaseColorTexture"] = 0
    m["pbrMetallicRoughnessTexture"] = 1
    m["normalTexture"] = 2
    m["emissiveTexture"] = 3
    m["emissiveFactor"] = (2.0, 2.0, 2.0)
    s.materials = m
    s.env = sky_environment(env_size[0], env_size[1], seed + 1)
    s.camera = dict(eye=(0.0, 0.4, 3.2), center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=45.0)
    return s


# ------------------------------------------------------------------------------ C3 / C4
def sponza_standin(detail=1.0, tex_size=512, env_size=(1024, 512), seed=99):
    """Arcade hall: tessellated floor, two rows of fluted columns with arches, side walls, a coffered
    ceiling with openings, alpha-masked drapes and foliage cards.  detail=1.0 gives ~262k triangles
    in ~100 primMeshes over 25 materials; detail scales the tessellation (tests use ~0.05)."""
    rng = np.random.default_rng(seed)
    T = tex_size
    d = max(detail, 0.01) * 0.3265   # detail = 1.0 -> ~262k triangles (Sponza's count)

    def res(n, lo=2):
        return max(lo, int(round(n * math.sqrt(d))))

    s = SceneDesc(name="sponza_standin")
    # --- textures: 0..5 colour, 6 mr, 7 normal, 8 drape alpha, 9 leaf alpha
    texs = []
    for k in range(6):
        n1, n2 = fbm((T, T), rng, 4, 4 + k), fbm((T, T), rng, 3, 16)
        tint = np.array([[0.75, 0.7, 0.62], [0.6, 0.55, 0.5], [0.7, 0.45, 0.35], [0.5, 0.55, 0.6], [0.8, 0.78, 0.7], [0.45, 0.4, 0.38]][k])
        texs.append(TextureDesc(to_rgba8((0.6 + 0.4 * n1)[..., None] * tint * (0.85 + 0.15 * n2)[..., None])))
    nr = fbm((T, T), rng, 4, 8)
    texs.append(TextureDesc(to_rgba8(np.stack([np.ones((T, T)), 0.35 + 0.6 * nr, 0.02 + 0.0 * nr], axis=-1))))
    gy, gx = np.gradient(fbm((T, T), rng, 5, 8))
    texs.append(TextureDesc(to_rgba8(_normalize(np.stack([-gx * T * 0.03, -gy * T * 0.03, np.ones_like(gx)], axis=-1)) * 0.5 + 0.5)))
    yy, xx = np.meshgrid(np.linspace(0, 1, T), np.linspace(0, 1, T), indexing="ij")
    stripes = (np.sin(xx * 40 * np.pi) > -0.3) & (np.sin(yy * 6 * np.pi + xx * 3) > -0.8)
    drape_rgb = np.stack([0.7 + 0.2 * np.sin(xx * 20), 0.1 + 0.1 * yy, 0.12 + 0 * xx], axis=-1)
    texs.append(TextureDesc(to_rgba8(drape_rgb, stripes.astype(np.float64)), wrap_s=10497, wrap_t=33071))
    leaf = (fbm((T, T), rng, 4, 16) > 0.52) & (((xx - 0.5) ** 2 + (yy - 0.5) ** 2) < 0.24)
    texs.append(TextureDesc(to_rgba8(np.stack([0.15 + 0.2 * yy, 0.45 + 0.3 * xx, 0.12 + 0 * xx], axis=-1), leaf.astype(np.float64))))
    s.textures = texs

    # --- materials (25)
    M = S.default_material(25)
    M["pbrMetallicFactor"] = 0.0
    for i in range(12):                      # stone / plaster variants
        M["pbrBaseColorTexture"][i] = i % 6
        M["pbrMetallicRoughnessTexture"][i] = 6
        M["normalTexture"][i] = 7
        M["pbrMetallicFactor"][i] = 1.0
        M["pbrRoughnessFactor"][i] = 0.6 + 0.03 * i
    for i in range(12, 18):                  # painted / glossy
        M["pbrBaseColorFactor"][i] = (0.3 + 0.1 * (i - 12), 0.5, 0.7 - 0.08 * (i - 12), 1.0)
        M["pbrRoughnessFactor"][i] = 0.2 + 0.1 * (i - 12)
        M["clearcoatFactor"][i] = 0.3 * ((i - 12) % 3)
        M["clearcoatRoughness"][i] = 0.1
    M["pbrBaseColorFactor"][18] = (0.95, 0.8, 0.4, 1)  # metal trims
    M["pbrMetallicFactor"][18] = 1.0
    M["pbrRoughnessFactor"][18] = 0.2
    M["pbrBaseColorFactor"][19] = (0.9, 0.9, 0.92, 1)
    M["pbrMetallicFactor"][19] = 1.0
    M["pbrRoughnessFactor"][19] = 0.05
    M["anisotropy"][19] = 0.6
    for i in (20, 21):                       # drapes: alpha MASK, double sided
        M["pbrBaseColorTexture"][i] = 8
        M["alphaMode"][i] = 1
        M["alphaCutoff"][i] = 0.5
        M["doubleSided"][i] = 1
        M["pbrRoughnessFactor"][i] = 0.9
        M["sheen"][i] = 0x80FFFFFF if i == 21 else 0
    M["pbrBaseColorTexture"][22] = 9         # foliage cards: alpha MASK
    M["alphaMode"][22] = 1
    M["doubleSided"][22] = 1
    M["pbrRoughnessFactor"][22] = 0.7
    M["emissiveFactor"][23] = (6.0, 4.5, 2.5)  # lanterns
    M["pbrBaseColorFactor"][23] = (0.1, 0.1, 0.1, 1)
    M["pbrBaseColorFactor"][24] = (1, 1, 1, 1)  # glass orbs
    M["transmissionFactor"][24] = 1.0
    M["pbrRoughnessFactor"][24] = 0.02
    M["ior"][24] = 1.45
    M["thicknessFactor"][24] = 0.1
    M["attenuationColor"][24] = (0.8, 0.95, 0.9)
    M["attenuationDistance"][24] = 0.5
    s.materials = M

    L, Wd, Hh = 30.0, 12.0, 9.0
    meshes, nodes = [], []

    def add(mesh, mat=np.eye(4, dtype=np.float32)):
        meshes.append(mesh)
        nodes.append((np.asarray(mat, dtype=np.float32), len(meshes) - 1))
        return len(meshes) - 1

    def inst(idx, mat):
        nodes.append((np.asarray(mat, dtype=np.float32), idx))

    # floor: 2 x 6 patches with a gentle height field
    def bumps(X, Z):
        return 0.015 * np.sin(X * 3.1) * np.sin(Z * 2.7)
    for ix in range(6):
        for iz in range(2):
            g = grid_mesh(res(96), res(96), L / 6, Wd / 2, (ix + iz) % 6, 3.0, bumps)
            add(g, translate(-L / 2 + (ix + 0.5) * L / 6, 0.0, -Wd / 4 + iz * Wd / 2))
    # side walls (facing inward) and end walls
    wall = grid_mesh(res(160), res(48), L, Hh, 6, 6.0, lambda X, Z: 0.03 * np.sin(X * 5) * np.cos(Z * 4))
    add(wall, translate(0, Hh / 2, -Wd / 2) @ rotate((1, 0, 0), math.pi / 2))
    inst(len(meshes) - 1, translate(0, Hh / 2, Wd / 2) @ rotate((1, 0, 0), -math.pi / 2))
    endw = grid_mesh(res(64), res(48), Wd, Hh, 7, 4.0)
    add(endw, translate(-L / 2, Hh / 2, 0) @ rotate((0, 0, 1), -math.pi / 2))
    inst(len(meshes) - 1, translate(L / 2, Hh / 2, 0) @ rotate((0, 0, 1), math.pi / 2))
    # ceiling with 5 openings = 6 strips
    for k in range(6):
        c = grid_mesh(res(40), res(64), L / 6 - 1.2, Wd, 8 + (k % 2), 3.0)
        add(c, translate(-L / 2 + (k + 0.5) * L / 6, Hh, 0) @ rotate((1, 0, 0), math.pi))
    # columns: 2 rows x 8, each its own primMesh variant pair (fluted shaft + capital), instanced
    shaft = cylinder(res(96, 8), res(80, 4), 0.42, 5.2, 10, lambda V, phi: 1.0 - 0.12 * V + 0.035 * np.cos(phi * 20))
    cap = cylinder(res(64, 8), res(24, 2), 0.62, 0.5, 11, lambda V, phi: 0.75 + 0.35 * V * V)
    base_blk = box_mesh((-0.6, 0.0, -0.6), (0.6, 0.35, 0.6), 9)
    i_shaft = i_cap = i_base = None
    for row, z in enumerate((-3.2, 3.2)):
        for k in range(8):
            x = -L / 2 + 2.2 + k * (L - 4.4) / 7
            if i_shaft is None:
                i_base = add(base_blk, translate(x, 0, z))
                i_shaft = add(shaft, translate(x, 0.35, z))
                i_cap = add(cap, translate(x, 5.55, z))
            else:
                inst(i_base, translate(x, 0, z))
                inst(i_shaft, translate(x, 0.35, z) @ rotate((0, 1, 0), 0.37 * k))
                inst(i_cap, translate(x, 5.55, z))
    # arches between neighbouring columns: half tori as displaced cylinders
    def arch_profile(V, phi):
        return 1.0 + 0.05 * np.cos(phi * 8)
    arch = cylinder(res(48, 6), res(64, 4), 0.28, math.pi, 3, arch_profile)
    # bend the straight cylinder (height = angle) into a half ring of radius R in the XY plane
    R = ((L - 4.4) / 7) / 2
    ang = arch.positions[:, 1].copy()
    rad = R + arch.positions[:, 0]
    bent = np.stack([-np.cos(ang) * rad, np.sin(ang) * rad, arch.positions[:, 2]], axis=1)
    arch = finish_mesh(bent, arch.indices, arch.texcoords, 3, smooth=True)
    i_arch = None
    for z in (-3.2, 3.2):
        for k in range(7):
            x = -L / 2 + 2.2 + (k + 0.5) * (L - 4.4) / 7
            if i_arch is None:
                i_arch = add(arch, translate(x, 6.05, z))
            else:
                inst(i_arch, translate(x, 6.05, z))
    # drapes: wavy alpha-masked cloth hanging between columns (non-opaque, double sided)
    def wave(X, Z):
        return 0.25 * np.sin(X * 2.4) * (0.5 + Z / 6.0)
    for k in range(6):
        dr = grid_mesh(res(64), res(64), 3.2, 3.6, 20 + (k % 2), 1.0, wave)
        add(dr, translate(-L / 2 + 4.0 + k * 4.4, 4.4, (-1) ** k * 3.2) @ rotate((1, 0, 0), math.pi / 2))
    # foliage: clusters of alpha-masked cards in planters
    card = grid_mesh(1, 1, 0.9, 0.9, 22, 1.0)
    i_card = None
    for p in range(10):
        px, pz = -L / 2 + 3.0 + p * 2.7, (-1) ** p * 1.4
        add(box_mesh((-0.5, 0.0, -0.5), (0.5, 0.6, 0.5), 12 + p % 6), translate(px, 0, pz))
        for c in range(max(4, int(40 * d))):
            m = translate(px + rng.uniform(-0.4, 0.4), 0.9 + rng.uniform(0, 1.2), pz + rng.uniform(-0.4, 0.4)) @ rotate(rng.normal(size=3), rng.uniform(0, 6.28))
            if i_card is None:
                i_card = add(card, m)
            else:
                inst(i_card, m)
    # lanterns (emissive) + glass orbs + metal trims
    lantern = uv_sphere(res(32, 8), res(16, 4), 0.22, 23)
    orb = uv_sphere(res(64, 8), res(32, 4), 0.45, 24)
    i_l = i_o = None
    for k in range(7):
        x = -L / 2 + 2.2 + (k + 0.5) * (L - 4.4) / 7
        if i_l is None:
            i_l = add(lantern, translate(x, 5.2, 0))
            i_o = add(orb, translate(x, 0.47, 0.0))
        else:
            inst(i_l, translate(x, 5.2, 0))
            inst(i_o, translate(x, 0.47, (-1) ** k * 0.8))
    trim = cylinder(res(24, 6), res(200, 4), 0.06, L - 2, 18)
    add(trim, translate(-L / 2 + 1, 6.9, -3.2) @ rotate((0, 0, 1), -math.pi / 2))
    inst(len(meshes) - 1, translate(-L / 2 + 1, 6.9, 3.2) @ rotate((0, 0, 1), -math.pi / 2))
    rail = cylinder(res(24, 6), res(200, 4), 0.05, L - 2, 19)
    add(rail, translate(-L / 2 + 1, 1.1, -5.2) @ rotate((0, 0, 1), -math.pi / 2))
    inst(len(meshes) - 1, translate(-L / 2 + 1, 1.1, 5.2) @ rotate((0, 0, 1), -math.pi / 2))

    s.meshes, s.nodes = meshes, nodes
    lights = np.concatenate([point_light((-6.0, 4.0, 0.0), (1.0, 0.85, 0.7), 40.0), point_light((7.0, 3.0, 1.0), (0.7, 0.8, 1.0), 30.0)])
    lights["type"][1] = 2                    # spot
    lights["direction"][1] = _normalize(np.array([-0.3, -1.0, 0.1]))
    lights["innerConeCos"][1] = math.cos(0.3)
    lights["outerConeCos"][1] = math.cos(0.6)
    s.lights = lights
    s.env = sky_environment(env_size[0], env_size[1], seed + 1)
    s.camera = dict(eye=(-12.5, 2.2, 0.6), center=(4.0, 3.0, -0.4), up=(0.0, 1.0, 0.0), fov=60.0)
    return s


# ------------------------------------------------------------------------------ C5
def instanced_stress(instances_per_axis=10, nu=72, nv=72, seed=5):
    """N^3 instances of one bumpy sphere (nu*nv*2 triangles each); 10^3 x 10,368 = 10.4 M triangles."""
    rng = np.random.default_rng(seed)
    mesh = uv_sphere(nu, nv, 0.42, 0, lambda dd: 0.08 * np.sin(dd[..., 0] * 11) * np.sin(dd[..., 1] * 9) * np.sin(dd[..., 2] * 10))
    s = SceneDesc(name="instanced_stress")
    s.meshes = [mesh, grid_mesh(8, 8, 4.0 * instances_per_axis, 4.0 * instances_per_axis, 1, 8.0)]
    n = instances_per_axis
    for i in range(n):
        for j in range(n):
            for k in range(n):
                m = translate((i - (n - 1) / 2) * 1.1, (j + 0.5) * 1.1, (k - (n - 1) / 2) * 1.1) @ rotate(rng.normal(size=3), rng.uniform(0, 6.28)) @ scale(rng.uniform(0.8, 1.1))
                s.nodes.append((m, 0))
    s.nodes.append((np.eye(4, dtype=np.float32), 1))
    M = S.default_material(2)
    M["pbrBaseColorFactor"][0] = (0.8, 0.6, 0.3, 1)
    M["pbrMetallicFactor"][0] = 0.3
    M["pbrRoughnessFactor"][0] = 0.4
    M["pbrBaseColorFactor"][1] = (0.5, 0.5, 0.5, 1)
    M["pbrMetallicFactor"][1] = 0.0
    s.materials = M
    s.env = sky_environment(1024, 512, seed + 1)
    ext = n * 1.1
    s.camera = dict(eye=(ext * 0.9, ext * 1.1, ext * 1.3), center=(0.0, ext * 0.45, 0.0), up=(0.0, 1.0, 0.0), fov=50.0)
    return s
