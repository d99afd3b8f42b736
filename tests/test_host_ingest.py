"""CPU tests of the glTF / image ingest path (vk_raytrace_b200/host/): the host half of Scene::load (src/scene.cpp:56-118)
and HdrSampling::loadEnvironment's decoder (src/hdr_sampling.cpp:64), restated without tinygltf / stb_image."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import small_sponza

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "vk_raytrace_b200", "host")


@pytest.fixture(scope="module")
def selftest(built):
    exe = os.path.join(HOST, "vkrt_host_selftest")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    return exe


def decode_png(selftest, path, tmp_path):
    raw = str(tmp_path / "out.raw")
    out = subprocess.run([selftest, "png", path, raw], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    w, h = map(int, out.stdout.split())
    return np.fromfile(raw, dtype=np.uint8).reshape(h, w, 4)


def test_png_decoder_matches_pillow(selftest, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    rgba = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    cases = {
        "rgba": Image.fromarray(rgba, "RGBA"), "rgb": Image.fromarray(rgba[..., :3].copy(), "RGB"), "gray": Image.fromarray(rgba[..., 0].copy(), "L"),
        "la": Image.fromarray(rgba[..., [0, 3]].copy(), "LA"), "pal": Image.fromarray(rgba[..., :3].copy(), "RGB").quantize(32),
        "big": Image.fromarray(np.tile(rgba, (8, 6, 1)), "RGBA"),
    }
    for name, im in cases.items():
        for opts in ({}, {"optimize": True}, {"compress_level": 0}):
            p = str(tmp_path / f"{name}.png")
            im.save(p, **opts)
            got = decode_png(selftest, p, tmp_path)
            want = np.asarray(Image.open(p).convert("RGBA"))
            assert np.array_equal(got, want), (name, opts)
    # 16-bit: the decoder keeps the high byte (what stb_image's 8-bit path does)
    g16 = (rng.integers(0, 65536, (9, 11), dtype=np.uint16))
    p = str(tmp_path / "g16.png")
    Image.fromarray(g16, "I;16").save(p)
    got = decode_png(selftest, p, tmp_path)
    assert np.array_equal(got[..., 0], (g16 >> 8).astype(np.uint8)) and np.all(got[..., 3] == 255)
    # our own writer (used by the glTF exporter)
    from vk_raytrace_b200.gltf_export import write_png
    p = str(tmp_path / "own.png")
    write_png(p, rgba)
    assert np.array_equal(decode_png(selftest, p, tmp_path), rgba)


def test_png_adam7_interlaced(selftest, tmp_path):
    """Hand-assembled Adam7 PNG (Pillow cannot write interlaced files)."""
    import zlib
    rng = np.random.default_rng(5)
    h, w = 13, 10
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    raw = b""
    for x0, y0, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
        sub = img[y0::dy, x0::dx]
        if sub.size:
            raw += b"".join(b"\x00" + sub[y].tobytes() for y in range(sub.shape[0]))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    p = str(tmp_path / "a7.png")
    open(p, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 1)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
    got = decode_png(selftest, p, tmp_path)
    assert np.array_equal(got[..., :3], img) and np.all(got[..., 3] == 255)


def test_hdr_and_pfm_readers(selftest, tmp_path):
    from vk_raytrace_b200 import scenes
    from vk_raytrace_b200.gltf_export import write_pfm
    env = scenes.sky_environment(64, 32, 9)
    p = str(tmp_path / "e.pfm")
    write_pfm(p, env)
    raw = str(tmp_path / "e.raw")
    out = subprocess.run([selftest, "hdr", p, raw], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["64", "32"]
    got = np.fromfile(raw, dtype=np.float32).reshape(32, 64, 4)
    assert np.array_equal(got[..., :3], env[..., :3]) and np.all(got[..., 3] == 1.0)
    # Radiance RGBE, flat and run-length encoded scanlines; decode = mantissa * 2^(e-136) (stb_image's hdr_convert)
    rgb = env[..., :3].astype(np.float64)
    m = rgb.max(axis=2)
    e = np.where(m > 1e-32, np.ceil(np.log2(np.maximum(m, 1e-38)) + 1e-12), -128).astype(int)
    rgbe = np.zeros((32, 64, 4), dtype=np.uint8)
    scale = np.where(m > 1e-32, 256.0 / (2.0 ** e), 0.0)
    rgbe[..., :3] = np.clip(np.floor(rgb * scale[..., None]), 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    want = np.where(rgbe[..., 3:4] > 0, rgbe[..., :3].astype(np.float32) * np.ldexp(1.0, rgbe[..., 3:4].astype(int) - 136).astype(np.float32), 0).astype(np.float32)
    for mode in ("flat", "rle"):
        body = b""
        for y in range(32):
            if mode == "flat":
                body += rgbe[y].tobytes()
            else:
                body += bytes([2, 2, 0, 64])
                for c in range(4):
                    row = rgbe[y, :, c]
                    x = 0
                    while x < 64:
                        n = min(64 - x, 100)
                        if x % 2 == 0 and np.all(row[x:x + n] == row[x]) and n > 2:
                            body += bytes([128 + n, row[x]])
                        else:
                            body += bytes([n]) + row[x:x + n].tobytes()
                        x += n
        p = str(tmp_path / f"{mode}.hdr")
        open(p, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 32 +X 64\n" + body)
        out = subprocess.run([selftest, "hdr", p, raw], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = np.fromfile(raw, dtype=np.float32).reshape(32, 64, 4)
        assert np.array_equal(got[..., :3], want), mode


def test_gltf_import_matches_exported_scene(selftest, tmp_path):
    """Export the Sponza stand-in as glTF (.gltf + .bin + PNG) and read it back through the C++ importer."""
    from vk_raytrace_b200.gltf_export import export_gltf
    sc = small_sponza()
    gltf, env = export_gltf(sc, str(tmp_path), "s")
    out = subprocess.run([selftest, "gltf", gltf], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    kv = lines[0].split()
    summary = dict(zip(kv[0::2][:7], kv[1::2][:7]))
    assert int(summary["primMeshes"]) == len(sc.meshes) and int(summary["nodes"]) == len(sc.nodes)
    assert int(summary["triangles"]) == sc.num_triangles() and int(summary["materials"]) == len(sc.materials)
    assert int(summary["textures"]) == len(sc.textures) and int(summary["lights"]) == len(sc.lights) and summary["camera"] == "1"
    sums = dict(zip(lines[1].split()[0::2], map(float, lines[1].split()[1::2])))
    want_pos = sum(float(m.positions.astype(np.float64).sum()) for m in sc.meshes)
    want_idx = sum(float(m.indices.astype(np.float64).sum()) for m in sc.meshes)
    M = sc.materials
    want_mat = float((M["pbrRoughnessFactor"].astype(np.float64) + M["pbrMetallicFactor"] + M["pbrBaseColorFactor"][:, 0] + M["alphaMode"] + M["doubleSided"]
                      + M["clearcoatFactor"] + M["transmissionFactor"] + M["ior"] + M["anisotropy"] + M["sheen"].astype(np.float64)).sum())
    assert abs(sums["sumPos"] - want_pos) < 1e-3 * max(1, abs(want_pos)) and sums["sumIdx"] == want_idx
    assert abs(sums["sumMat"] - want_mat) < 1e-6 * want_mat
    spot = [l for l in lines if l.startswith("light type 2")][0].split()
    assert np.allclose([float(v) for v in spot[8:11]], sc.lights["direction"][1], atol=1e-6)
    assert abs(float(spot[-2]) - float(sc.lights["innerConeCos"][1])) < 1e-6
    # .glb container: same JSON + BIN in one file
    import json
    doc = json.load(open(gltf))
    binb = open(os.path.join(str(tmp_path), "s.bin"), "rb").read()
    del doc["buffers"][0]["uri"]
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    binb += b"\0" * (-len(binb) % 4)
    glb = str(tmp_path / "s.glb")
    open(glb, "wb").write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(binb)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(binb), 0x004E4942) + binb)
    out2 = subprocess.run([selftest, "gltf", glb], capture_output=True, text=True)
    assert out2.returncode == 0 and out2.stdout.splitlines()[1] == lines[1]


def decode_image(selftest, path, tmp_path):
    raw = str(tmp_path / "out.raw")
    out = subprocess.run([selftest, "jpg", path, raw], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    w, h = map(int, out.stdout.split())
    return np.fromfile(raw, dtype=np.uint8).reshape(h, w, 4)


def _photo(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0), 128 + 90 * np.sin((xx + yy) / 11.0), 128 + 80 * np.cos(xx / 5.0 + yy / 13.0)], -1)
    return np.clip(img + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)


def test_jpeg_decoder_matches_pillow_exactly(selftest, tmp_path):
    """Baseline / progressive, 4:4:4 / 4:2:2 / 4:2:0 / grey, optimised tables, restart intervals, odd sizes: the decoder implements the IJG
    reference arithmetic (islow IDCT, triangle upsampling, 16-bit YCbCr conversion), so it must equal Pillow (libjpeg-turbo) bit for bit."""
    from PIL import Image
    n = 0
    for (h, w) in ((64, 64), (37, 53), (100, 131), (8, 8), (1, 1), (3, 2), (17, 5), (250, 16)):
        img = _photo(h, w, h * 131 + w)
        for name, kw in (("444", dict(subsampling=0)), ("422", dict(subsampling=1)), ("420", dict(subsampling=2)),
                         ("prog420", dict(subsampling=2, progressive=True)), ("prog444", dict(subsampling=0, progressive=True)), ("grey", dict()),
                         ("q30", dict(quality=30, subsampling=2)), ("q100", dict(quality=100, subsampling=0)), ("opt", dict(optimize=True, subsampling=2)),
                         ("rst", dict(subsampling=2, restart_marker_blocks=3)), ("rstprog", dict(subsampling=1, progressive=True, restart_marker_rows=1))):
            p = str(tmp_path / f"{name}.jpg")
            kw = dict(kw)
            Image.fromarray(img if name != "grey" else img[..., 0].copy()).save(p, quality=kw.pop("quality", 90), **kw)
            got = decode_image(selftest, p, tmp_path)
            want = np.asarray(Image.open(p).convert("RGBA"))
            assert np.array_equal(got, want), (h, w, name, int(np.abs(got.astype(int) - want.astype(int)).max()))
            n += 1
    assert n > 80
    # garbage and truncation are errors, not crashes
    blob = open(str(tmp_path / "420.jpg"), "rb").read()
    for bad in (blob[:100], blob[:len(blob) // 2], b"\xff\xd8\xff" + bytes(200), blob[:20] + bytes(64) + blob[84:]):
        p = str(tmp_path / "bad.jpg")
        open(p, "wb").write(bad)
        out = subprocess.run([selftest, "jpg", p, str(tmp_path / "o.raw")], capture_output=True, text=True)
        assert out.returncode in (0, 2), out.returncode      # decoded partially or rejected; never a signal


def test_gltf_jpeg_textures_and_sparse_accessors(selftest, tmp_path):
    """A hand-written glTF with what the real Khronos assets use and round 1 lacked: JPEG images (external file and base64 data URI)
    and sparse accessors (with and without a base bufferView, glTF 2.0 section 3.6.2.3)."""
    import base64
    import io
    import json
    from PIL import Image
    tex_a, tex_b = _photo(32, 48, 1), _photo(16, 16, 2)
    Image.fromarray(tex_a).save(str(tmp_path / "a.jpg"), quality=85, subsampling=2)
    bio = io.BytesIO()
    Image.fromarray(tex_b).save(bio, "JPEG", quality=92, subsampling=0, progressive=True)
    want_tex = 0.0
    for src in (str(tmp_path / "a.jpg"), bio):
        if not isinstance(src, str):
            src.seek(0)
        im = np.asarray(Image.open(src).convert("RGBA")).astype(np.float64)
        want_tex += im.sum() + 1000003.0 * im.shape[1] + 7919.0 * im.shape[0]
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint16)
    sp_idx = np.array([1, 3], np.uint8)                         # replace vertices 1 and 3
    sp_val = np.array([[2, 0, 0.5], [2, 2, -1]], np.float32)
    sp2_idx = np.array([0, 2], np.uint16)                       # second mesh: zero base, two vertices set
    sp2_val = np.array([[0, 0, 1], [0, 3, 1]], np.float32)
    blob = b""
    views = []

    def add(data):
        nonlocal blob
        blob += b"\0" * (-len(blob) % 4)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)})
        blob += data
        return len(views) - 1
    v_base, v_idx, v_si, v_sv, v_s2i, v_s2v = (add(a.tobytes()) for a in (base, idx, sp_idx, sp_val, sp2_idx, sp2_val))
    open(str(tmp_path / "m.bin"), "wb").write(blob)
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 1]}],
        "nodes": [{"mesh": 0}, {"mesh": 1, "translation": [0, 0, 2]}],
        "buffers": [{"uri": "m.bin", "byteLength": len(blob)}], "bufferViews": views,
        "accessors": [
            {"bufferView": v_base, "componentType": 5126, "count": 4, "type": "VEC3",
             "sparse": {"count": 2, "indices": {"bufferView": v_si, "componentType": 5121}, "values": {"bufferView": v_sv}}},
            {"bufferView": v_idx, "componentType": 5123, "count": 6, "type": "SCALAR"},
            {"componentType": 5126, "count": 3, "type": "VEC3",
             "sparse": {"count": 2, "indices": {"bufferView": v_s2i, "componentType": 5123}, "values": {"bufferView": v_s2v}}}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "material": 0}]},
                   {"primitives": [{"attributes": {"POSITION": 2}, "material": 0}]}],
        "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}}}],
        "textures": [{"source": 0}, {"source": 1}],
        "images": [{"uri": "a.jpg"}, {"uri": "data:image/jpeg;base64," + base64.b64encode(bio.getvalue()).decode()}],
    }
    p = str(tmp_path / "m.gltf")
    json.dump(doc, open(p, "w"))
    out = subprocess.run([selftest, "gltf", p], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    sums = dict(zip(lines[1].split()[0::2], map(float, lines[1].split()[1::2])))
    pos0 = base.copy()
    pos0[sp_idx] = sp_val
    pos1 = np.zeros((3, 3), np.float32)
    pos1[sp2_idx] = sp2_val
    assert abs(sums["sumPos"] - float(pos0.sum() + pos1.sum())) < 1e-5
    assert sums["sumTex"] == want_tex
    assert "primMeshes 2 nodes 2 triangles 3" in lines[0] and "textures 2" in lines[0]
    # malformed files fail with a message instead of reading out of bounds or looping
    bad = json.loads(json.dumps(doc))
    bad["accessors"].append({"bufferView": v_base, "componentType": 5126, "count": 3, "type": "VEC3"})
    bad["meshes"][0]["primitives"][0]["attributes"]["NORMAL"] = 3          # 3 normals for 4 positions
    json.dump(bad, open(str(tmp_path / "bad1.gltf"), "w"))
    out = subprocess.run([selftest, "gltf", str(tmp_path / "bad1.gltf")], capture_output=True, text=True)
    assert out.returncode == 2 and "different count" in out.stderr
    bad = json.loads(json.dumps(doc))
    bad["nodes"] = [{"children": [1]}, {"children": [0], "mesh": 0}]       # cycle
    json.dump(bad, open(str(tmp_path / "bad2.gltf"), "w"))
    out = subprocess.run([selftest, "gltf", str(tmp_path / "bad2.gltf")], capture_output=True, text=True, timeout=20)
    assert out.returncode == 2 and "cyclic" in out.stderr
    bad = json.loads(json.dumps(doc))
    bad["accessors"][0]["sparse"]["indices"]["componentType"] = 5126       # float indices are not allowed
    json.dump(bad, open(str(tmp_path / "bad3.gltf"), "w"))
    assert subprocess.run([selftest, "gltf", str(tmp_path / "bad3.gltf")], capture_output=True, text=True).returncode == 2


def test_png_header_validation(selftest, tmp_path):
    """crafted IHDR fields (bad length, illegal bit depth, absurd size, data that does not match the header) are errors."""
    import zlib

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    def png(ihdr, raw, ihdr_extra=b""):
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr + ihdr_extra) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    good_raw = b"".join(b"\x00" + bytes([10, 20, 30] * 4) for _ in range(3))
    cases = {
        "ok": (png(struct.pack(">IIBBBBB", 4, 3, 8, 2, 0, 0, 0), good_raw), 0),
        "depth0": (png(struct.pack(">IIBBBBB", 4, 3, 0, 0, 0, 0, 0), good_raw), 2),
        "depth3": (png(struct.pack(">IIBBBBB", 4, 3, 3, 3, 0, 0, 0), good_raw), 2),
        "depth4_rgb": (png(struct.pack(">IIBBBBB", 4, 3, 4, 2, 0, 0, 0), good_raw), 2),
        "huge": (png(struct.pack(">IIBBBBB", 0xFFFFFFFF, 0x40000001, 1, 0, 0, 0, 0), good_raw), 2),
        "long_ihdr": (png(struct.pack(">IIBBBBB", 4, 3, 8, 2, 0, 0, 0), good_raw, b"\0\0"), 2),
        "short_data": (png(struct.pack(">IIBBBBB", 4, 30, 8, 2, 0, 0, 0), good_raw), 2),
        "long_data": (png(struct.pack(">IIBBBBB", 4, 2, 8, 2, 0, 0, 0), good_raw), 2),
        "filter_method": (png(struct.pack(">IIBBBBB", 4, 3, 8, 2, 0, 1, 0), good_raw), 2),
    }
    for name, (blob, want) in cases.items():
        p = str(tmp_path / f"{name}.png")
        open(p, "wb").write(blob)
        out = subprocess.run([selftest, "png", p, str(tmp_path / "o.raw")], capture_output=True, text=True)
        assert out.returncode == want, (name, out.returncode, out.stderr)
    # Radiance header with negative / absurd dimensions
    for dims in ("-Y -5 +X 64", "-Y 5 +X 99999999"):
        p = str(tmp_path / "bad.hdr")
        open(p, "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + dims.encode() + b"\n" + bytes(64))
        assert subprocess.run([selftest, "hdr", p, str(tmp_path / "o.raw")], capture_output=True, text=True).returncode == 2


def test_decoders_survive_mutated_files(built, tmp_path):
    """Mutation fuzzing of every file format the ingest reads (PNG colour types, baseline / progressive / restart-interval JPEG, Radiance
    HDR flat + RLE, PFM, glTF JSON with sparse accessors and JPEG textures, GLB), with the tool built under AddressSanitizer +
    UBSan: a crafted texture or scene may be rejected or decoded to garbage, never read or write out of bounds (ADVICE round 1:
    the decoders trusted file headers).  Seeded, so a failure reproduces with `vkrt_host_selftest fuzz <file> <seed> <n>`."""
    from PIL import Image
    from vk_raytrace_b200 import scenes
    from vk_raytrace_b200.gltf_export import export_gltf, write_pfm
    exe = str(tmp_path / "selftest_asan")
    src = ["host_selftest.cpp", "gltf_scene.cpp", "image_io.cpp", "jpeg_decode.cpp", "cuda_renderer.cpp"]
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] + src +
                        ["-I../../include", "-L..", "-lvkrt_b200", "-Wl,-rpath," + os.path.dirname(HOST), "-o", exe], cwd=HOST, capture_output=True, text=True)
    if cc.returncode != 0 and "sanitize" in cc.stderr:
        pytest.skip("no sanitizer runtime on this image")
    assert cc.returncode == 0, cc.stderr
    rng = np.random.default_rng(7)
    a = (rng.random((24, 20, 4)) * 255).astype(np.uint8)
    g = np.linspace(0, 255, 48 * 40 * 3).reshape(48, 40, 3).astype(np.uint8)
    d = str(tmp_path)
    Image.fromarray(a, "RGBA").save(d + "/rgba.png")
    Image.fromarray(a[..., 0].copy(), "L").save(d + "/gray.png")
    Image.fromarray(a[..., :3].copy(), "RGB").convert("P").save(d + "/pal.png")
    Image.fromarray(g, "RGB").save(d + "/base.jpg", quality=85, subsampling=2)
    Image.fromarray(g, "RGB").save(d + "/prog.jpg", quality=70, progressive=True, subsampling=1)
    Image.fromarray(g, "RGB").save(d + "/rst.jpg", quality=85, subsampling=0, restart_marker_blocks=2)
    env = scenes.sky_environment(32, 16, 9)
    write_pfm(d + "/e.pfm", env)
    rgbe = np.zeros((16, 32, 4), dtype=np.uint8)
    rgbe[..., :3] = (rng.random((16, 32, 3)) * 255).astype(np.uint8)
    rgbe[..., 3] = 128
    rgbe[:, 8:20] = rgbe[:, 8:9]                                   # runs for the RLE writer
    body = b""
    for y in range(16):
        body += bytes([2, 2, 0, 32])
        for c in range(4):
            row, x = rgbe[y, :, c], 0
            while x < 32:
                n = 1
                while x + n < 32 and row[x + n] == row[x] and n < 127:
                    n += 1
                if n > 2:
                    body += bytes([128 + n, row[x]])
                else:
                    n = 1
                    body += bytes([1, row[x]])
                x += n
    open(d + "/rle.hdr", "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 16 +X 32\n" + body)
    open(d + "/flat.hdr", "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 16 +X 32\n" + rgbe.tobytes())
    export_gltf(small_sponza(), d, "s")                              # s.gltf + s.bin + PNG textures
    env_san = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    # the unmutated seeds decode (n = 0 mutations would not prove that: use the ordinary modes of the same binary)
    for f, mode in (("rgba.png", "png"), ("prog.jpg", "jpg"), ("rle.hdr", "hdr"), ("e.pfm", "hdr")):
        r = subprocess.run([exe, mode, d + "/" + f, d + "/o.raw"], capture_output=True, text=True, env=env_san)
        assert r.returncode == 0, (f, r.stderr)
    total_ok = total_bad = 0
    for f, n in (("rgba.png", 400), ("gray.png", 300), ("pal.png", 300), ("base.jpg", 400), ("prog.jpg", 400), ("rst.jpg", 300),
                 ("e.pfm", 200), ("flat.hdr", 200), ("rle.hdr", 400), ("s.gltf", 120)):
        r = subprocess.run([exe, "fuzz", d + "/" + f, "17", str(n)], capture_output=True, text=True, env=env_san, timeout=600)
        assert r.returncode == 0, (f, r.stdout[-300:], r.stderr[-2000:])      # a sanitizer report exits non-zero
        w = r.stdout.split()
        assert w[0] == "ok" and w[2] == "rejected" and int(w[1]) + int(w[3]) == n, r.stdout
        total_ok += int(w[1])
        total_bad += int(w[3])
    assert total_ok > 100 and total_bad > 100                        # the mutations exercise both the accepting and the rejecting paths
