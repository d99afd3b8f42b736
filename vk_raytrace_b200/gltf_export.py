"""Writes a `SceneDesc` as a standard glTF 2.0 asset (.gltf + .bin + PNG textures) and its environment as a PFM file,
so that the procedural stand-ins of BASELINE.json's configs exist as real glTF files and the C++ ingest path
(vk_raytrace_b200/host/gltf_scene.cpp, the reference's Scene::load) can be exercised end to end."""
import json
import math
import os
import struct
import zlib

import numpy as np


def write_png(path, rgba):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def write_pfm(path, rgba32f):
    img = np.ascontiguousarray(rgba32f[..., :3], dtype="<f4")
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(f"PF\n{w} {h}\n-1.0\n".encode())
        f.write(img[::-1].tobytes())     # PFM rows run bottom-up


def _f(x):
    return float(np.float32(x))


def _fl(a):
    return [_f(v) for v in np.asarray(a, dtype=np.float32).reshape(-1)]


def export_gltf(scene, directory, name="scene"):
    """Returns (gltf_path, env_path)."""
    os.makedirs(directory, exist_ok=True)
    blob = bytearray()
    views, accessors = [], []

    def add(arr, comp, typ, target=None, minmax=False):
        arr = np.ascontiguousarray(arr)
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": arr.nbytes}
        if target:
            v["target"] = target
        views.append(v)
        blob.extend(arr.tobytes())
        a = {"bufferView": len(views) - 1, "componentType": comp, "count": int(arr.shape[0]), "type": typ}
        if minmax:
            a["min"], a["max"] = _fl(arr.min(axis=0)), _fl(arr.max(axis=0))
        accessors.append(a)
        return len(accessors) - 1

    meshes = []
    for m in scene.meshes:
        prim = {"attributes": {"POSITION": add(m.positions.astype("<f4"), 5126, "VEC3", 34962, True), "NORMAL": add(m.normals.astype("<f4"), 5126, "VEC3", 34962),
                               "TANGENT": add(m.tangents.astype("<f4"), 5126, "VEC4", 34962), "TEXCOORD_0": add(m.texcoords.astype("<f4"), 5126, "VEC2", 34962),
                               "COLOR_0": add(m.colors.astype("<f4"), 5126, "VEC4", 34962)},
                "indices": add(m.indices.astype("<u4"), 5125, "SCALAR", 34963), "material": int(m.material), "mode": 4}
        meshes.append({"primitives": [prim]})

    nodes = [{"mesh": int(mi), "matrix": _fl(np.asarray(mat, dtype=np.float32).T)} for mat, mi in scene.nodes]

    materials = []
    for m in scene.materials:
        def tex(i):
            return {"index": int(i)}
        pbr = {"baseColorFactor": _fl(m["pbrBaseColorFactor"]), "metallicFactor": _f(m["pbrMetallicFactor"]), "roughnessFactor": _f(m["pbrRoughnessFactor"])}
        if m["pbrBaseColorTexture"] >= 0:
            pbr["baseColorTexture"] = tex(m["pbrBaseColorTexture"])
        if m["pbrMetallicRoughnessTexture"] >= 0:
            pbr["metallicRoughnessTexture"] = tex(m["pbrMetallicRoughnessTexture"])
        g = {"pbrMetallicRoughness": pbr, "emissiveFactor": _fl(m["emissiveFactor"]), "alphaMode": ["OPAQUE", "MASK", "BLEND"][int(m["alphaMode"])],
             "alphaCutoff": _f(m["alphaCutoff"]), "doubleSided": bool(m["doubleSided"])}
        if m["emissiveTexture"] >= 0:
            g["emissiveTexture"] = tex(m["emissiveTexture"])
        if m["normalTexture"] >= 0:
            g["normalTexture"] = {"index": int(m["normalTexture"]), "scale": _f(m["normalTextureScale"])}
        ext = {}
        if m["unlit"]:
            ext["KHR_materials_unlit"] = {}
        if m["transmissionFactor"] != 0 or m["transmissionTexture"] >= 0:
            ext["KHR_materials_transmission"] = {"transmissionFactor": _f(m["transmissionFactor"])}
            if m["transmissionTexture"] >= 0:
                ext["KHR_materials_transmission"]["transmissionTexture"] = tex(m["transmissionTexture"])
        if m["ior"] != np.float32(1.5):
            ext["KHR_materials_ior"] = {"ior": _f(m["ior"])}
        if m["anisotropy"] != 0:
            ext["KHR_materials_anisotropy"] = {"anisotropyStrength": _f(m["anisotropy"]),
                                               "anisotropyRotation": float(math.atan2(float(m["anisotropyDirection"][0]), float(m["anisotropyDirection"][1])))}
        vol = {}
        if m["thicknessFactor"] != 0:
            vol["thicknessFactor"] = _f(m["thicknessFactor"])
        if tuple(m["attenuationColor"]) != (1.0, 1.0, 1.0):
            vol["attenuationColor"] = _fl(m["attenuationColor"])
        if m["attenuationDistance"] != np.finfo(np.float32).max:
            vol["attenuationDistance"] = _f(m["attenuationDistance"])
        if vol:
            ext["KHR_materials_volume"] = vol
        if m["clearcoatFactor"] != 0 or m["clearcoatRoughness"] != 0:
            ext["KHR_materials_clearcoat"] = {"clearcoatFactor": _f(m["clearcoatFactor"]), "clearcoatRoughnessFactor": _f(m["clearcoatRoughness"])}
        if m["sheen"] != 0:
            s = int(m["sheen"])
            ext["KHR_materials_sheen"] = {"sheenColorFactor": [(s & 255) / 255.0, ((s >> 8) & 255) / 255.0, ((s >> 16) & 255) / 255.0], "sheenRoughnessFactor": (s >> 24) / 255.0}
        if ext:
            g["extensions"] = ext
        materials.append(g)

    images, textures, samplers = [], [], []
    for i, t in enumerate(scene.textures):
        fn = f"{name}_tex{i}.png"
        write_png(os.path.join(directory, fn), t.rgba)
        images.append({"uri": fn})
        samplers.append({"magFilter": int(t.mag_filter), "minFilter": int(t.min_filter), "wrapS": int(t.wrap_s), "wrapT": int(t.wrap_t)})
        textures.append({"source": i, "sampler": i})

    root_nodes = list(range(len(nodes)))
    doc = {"asset": {"version": "2.0", "generator": "vk_raytrace_b200.gltf_export"}, "extensionsUsed": []}
    lights = []
    if scene.lights is not None:
        for l in scene.lights:
            typ = ["directional", "point", "spot"][int(l["type"])]
            jl = {"type": typ, "color": _fl(l["color"]), "intensity": _f(l["intensity"])}
            if l["range"] > 0:
                jl["range"] = _f(l["range"])
            if typ == "spot":
                jl["spot"] = {"innerConeAngle": float(math.acos(min(1.0, float(l["innerConeCos"])))), "outerConeAngle": float(math.acos(min(1.0, float(l["outerConeCos"]))))}
            lights.append(jl)
            d = np.asarray(l["direction"], dtype=np.float64)
            z = -d / max(np.linalg.norm(d), 1e-20)              # the light looks down its local -Z
            x = np.cross([0.0, 1.0, 0.0], z)
            if np.linalg.norm(x) < 1e-6:
                x = np.cross([1.0, 0.0, 0.0], z)
            x /= np.linalg.norm(x)
            y = np.cross(z, x)
            M = np.eye(4, dtype=np.float32)
            M[:3, 0], M[:3, 1] = x, y
            M[:3, 2] = -np.asarray(l["direction"], dtype=np.float32)   # so that M * (0,0,-1,0) is the stored direction, bit for bit
            M[:3, 3] = l["position"]
            nodes.append({"matrix": _fl(M.T), "extensions": {"KHR_lights_punctual": {"light": len(lights) - 1}}})
            root_nodes.append(len(nodes) - 1)
        doc["extensions"] = {"KHR_lights_punctual": {"lights": lights}}
        doc["extensionsUsed"].append("KHR_lights_punctual")
    cam = scene.camera
    if cam:
        eye, ctr, up = (np.asarray(cam[k], dtype=np.float64) for k in ("eye", "center", "up"))
        f = (ctr - eye) / np.linalg.norm(ctr - eye)
        s_ = np.cross(f, up)
        s_ /= np.linalg.norm(s_)
        u = np.cross(s_, f)
        M = np.eye(4)
        M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = s_, u, -f, eye
        nodes.append({"matrix": _fl(M.T.astype(np.float32)), "camera": 0})
        root_nodes.append(len(nodes) - 1)
        doc["cameras"] = [{"type": "perspective", "perspective": {"yfov": math.radians(cam["fov"]), "znear": 0.001, "zfar": 100000.0}}]
    for g in materials:
        for e in g.get("extensions", {}):
            if e not in doc["extensionsUsed"]:
                doc["extensionsUsed"].append(e)
    doc.update({"scene": 0, "scenes": [{"nodes": root_nodes}], "nodes": nodes, "meshes": meshes, "materials": materials, "accessors": accessors,
                "bufferViews": views, "buffers": [{"uri": f"{name}.bin", "byteLength": len(blob)}]})
    if textures:
        doc.update({"textures": textures, "images": images, "samplers": samplers})
    if not doc["extensionsUsed"]:
        del doc["extensionsUsed"]
    with open(os.path.join(directory, f"{name}.bin"), "wb") as f:
        f.write(bytes(blob))
    gltf_path = os.path.join(directory, f"{name}.gltf")
    with open(gltf_path, "w") as f:
        json.dump(doc, f)
    env_path = os.path.join(directory, f"{name}_env.pfm")
    write_pfm(env_path, scene.env)
    return gltf_path, env_path
