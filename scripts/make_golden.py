"""Generates tests/golden/* from the CPU oracle (the reference ships no golden vectors and cannot run here,
SURVEY.md 8c).  Re-run only when the oracle changes on purpose:  python scripts/make_golden.py
The integer KATs in rng_kat.json are additionally cross-checked in tests/test_oracle.py against an independent
pure-Python restatement of the published TEA / PCG-RXS-M-XS algorithms."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from vk_raytrace_b200 import scenes, structs as S, upload_scene  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)
L = ob.load()

# ---- RNG known answers -------------------------------------------------------------------------
kat = {"tea": [], "pcg": [], "rand": [], "pcg3d": [], "alpha_draw": []}
for a, b in [(0, 0), (1, 0), (0, 1), (123456, 789), (0xFFFFFFFF, 0xFFFFFFFF), (256 * 100 + 17, 5), (1920 * 1079 + 1919, 255)]:
    kat["tea"].append([a, b, int(L.oracle_tea(a, b))])
for seed in (0, 1, 0xDEADBEEF, 2891336453):
    st = C.c_uint32(seed)
    outs = [int(L.oracle_pcg(C.byref(st))) for _ in range(4)]
    kat["pcg"].append([seed, outs, int(st.value)])
    st = C.c_uint32(seed)
    kat["rand"].append([seed, [float(L.oracle_rand(C.byref(st))) for _ in range(4)]])
for v in [(0, 0, 0), (1, 2, 0), (1919, 1079, 0), (77, 300, 0)]:
    arr = (C.c_uint32 * 3)(*v)
    L.oracle_pcg3d(arr)
    kat["pcg3d"].append([list(v), [int(x) for x in arr]])
for seed, tri in [(0, 0), (12345, 678), (0xFFFFFFFF, 262143)]:
    kat["alpha_draw"].append([seed, tri, float(L.oracle_alpha_draw(seed, tri))])
json.dump(kat, open(os.path.join(G, "rng_kat.json"), "w"), indent=1)

# ---- oct encoding --------------------------------------------------------------------------------
rng = np.random.default_rng(2024)
v = rng.normal(size=(64, 3))
v /= np.linalg.norm(v, axis=1, keepdims=True)
v = np.concatenate([v, [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [0.5, -0.5, -0.70710678]]]).astype(np.float32)
enc = np.array([L.oracle_compress_unit_vec(*map(float, x)) for x in v], dtype=np.uint32)
dec = np.zeros((len(v), 3), dtype=np.float32)
for i, e in enumerate(enc):
    L.oracle_decompress_unit_vec(int(e), dec[i].ctypes.data_as(C.c_void_p))
np.savez(os.path.join(G, "oct.npz"), v=v, enc=enc, dec=dec)

# ---- BSDF tables --------------------------------------------------------------------------------
def bsdf_inputs(n, seed):
    r = np.random.default_rng(seed)
    rows = []
    for _ in range(n):
        N = np.array([0, 0, 1.0])
        T, B = np.array([1.0, 0, 0]), np.array([0, 1.0, 0])
        def hemi(allow_below=False):
            d = r.normal(size=3)
            d /= np.linalg.norm(d)
            if not allow_below:
                d[2] = abs(d[2]) * 0.98 + 0.02
                d /= np.linalg.norm(d)
            return d
        V, Ld = hemi(), hemi(allow_below=r.random() < 0.3)
        albedo = r.uniform(0.05, 0.95, 3)
        metallic, rough = r.choice([0.0, 0.3, 1.0]), r.uniform(0.05, 1.0)
        trans = r.choice([0.0, 0.0, 0.6, 1.0])
        ior = r.uniform(1.1, 1.8)
        eta = 1.0 / ior if r.random() < 0.7 else ior
        clearcoat, ccr = r.choice([0.0, 0.5, 1.0]), r.uniform(0.01, 0.5)
        aniso = r.choice([0.0, 0.0, 0.5])
        sheen, sheenTint = r.choice([0.0, 0.7]), r.uniform(0, 1, 3)
        thin = float(r.random() < 0.3)
        rows.append(np.concatenate([V, N, Ld, T, B, albedo, [metallic, rough, trans, ior, eta, clearcoat, ccr, aniso, sheen], sheenTint, [thin, 0.0]]))
    return np.array(rows, dtype=np.float32)

inp = bsdf_inputs(96, 7)
ev = np.zeros((2, len(inp), 4), dtype=np.float32)
sm = np.zeros((2, len(inp), 8), dtype=np.float32)
for mode in (0, 1):
    for i, row in enumerate(inp):
        L.oracle_bsdf_eval(mode, row.ctypes.data_as(C.c_void_p), ev[mode, i].ctypes.data_as(C.c_void_p))
        L.oracle_bsdf_sample(mode, row.ctypes.data_as(C.c_void_p), 1000 + i, sm[mode, i].ctypes.data_as(C.c_void_p))
np.savez(os.path.join(G, "bsdf.npz"), inputs=inp, eval=ev, sample=sm)

# ---- sun & sky ------------------------------------------------------------------------------------
ss = S.default_sunsky(in_use=1)
dirs = rng.normal(size=(48, 3))
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
dirs = np.concatenate([dirs, [[0.0, 0.78, 0.62] / np.linalg.norm([0.0, 0.78, 0.62]), [0, 1, 0], [0, -1, 0], [1, 0, 0]]]).astype(np.float32)
sky = np.zeros((len(dirs), 3), dtype=np.float32)
for i, d in enumerate(dirs):
    L.oracle_sun_and_sky(C.byref(ss), d.ctypes.data_as(C.c_void_p), sky[i].ctypes.data_as(C.c_void_p))
np.savez(os.path.join(G, "sunsky.npz"), dirs=dirs, rgb=sky)

# ---- C1: Cornell 64x64 images (frame 0 .. 3, depth 4) and debug views (depth 1) ----------------------------
sc = scenes.cornell_box(with_light=True)
W = H = 64
o = ob.Oracle(threads=1)
upload_scene(o, sc, W, H)
imgs = {}
for pbr in (0, 1):
    st = S.default_rtx_state(W, H, maxDepth=4, pbrMode=pbr)
    o.output_create(W, H)
    o.render(st, 4)
    imgs[f"radiance_pbr{pbr}"] = o.output_read()[..., :3].copy()
for dbg in range(1, 12):
    st = S.default_rtx_state(W, H, maxDepth=1 if dbg < 9 else 3, debugging_mode=dbg)
    o.output_create(W, H)
    o.render(st, 1)
    imgs[f"debug{dbg}"] = o.output_read()[..., :3].copy()
tm = S.default_tonemapper()
st = S.default_rtx_state(W, H, maxDepth=4)
o.output_create(W, H)
o.render(st, 4)
imgs["tonemapped_rgba8"] = o.tonemap(tm)
imgs["tonemapped_auto1_rgba8"] = o.tonemap(S.default_tonemapper(autoExposure=1))   # post.frag:64-70 through the mip pyramid
imgs["tonemapped_auto3_rgba8"] = o.tonemap(S.default_tonemapper(autoExposure=3))   # post.frag:72-95 (local exposure)
rays = np.zeros((256, 8), dtype=np.float32)
r2 = np.random.default_rng(5)
rays[:, :3] = r2.uniform(-0.9, 0.9, (256, 3))
d = r2.normal(size=(256, 3))
rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
rays[:, 3] = 1e32
imgs["probe_rays"] = rays
imgs["probe_hits_closest"] = o.trace_rays(rays, True)
imgs["probe_hits_any"] = o.trace_rays(rays, False)
np.savez_compressed(os.path.join(G, "cornell64.npz"), **imgs)
print("golden vectors written to", G, {k: os.path.getsize(os.path.join(G, k)) for k in os.listdir(G)})
