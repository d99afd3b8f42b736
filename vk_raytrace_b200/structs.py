"""ctypes / numpy mirrors of include/vkrt_host_device.h (reference shaders/host_device.h:107-282).

Sizes are asserted against SURVEY.md Appendix A at import time.
"""
import ctypes as C
import numpy as np

f32, i32, u32 = C.c_float, C.c_int32, C.c_uint32


class SceneCamera(C.Structure):
    _pack_ = 4
    _fields_ = [("viewInverse", f32 * 16), ("projInverse", f32 * 16), ("focalDist", f32), ("aperture", f32), ("nbLights", i32)]


class RtxState(C.Structure):
    _pack_ = 4
    _fields_ = [("frame", i32), ("maxDepth", i32), ("maxSamples", i32), ("fireflyClampThreshold", f32), ("hdrMultiplier", f32),
                ("debugging_mode", i32), ("pbrMode", i32), ("_pad0", i32), ("size", i32 * 2), ("minHeatmap", i32), ("maxHeatmap", i32)]


class Tonemapper(C.Structure):
    _pack_ = 4
    _fields_ = [("brightness", f32), ("contrast", f32), ("saturation", f32), ("vignette", f32), ("avgLum", f32), ("zoom", f32),
                ("renderingRatio", f32 * 2), ("autoExposure", i32), ("Ywhite", f32), ("key", f32), ("dither", i32)]


class SunAndSky(C.Structure):
    _pack_ = 4
    _fields_ = [("rgb_unit_conversion", f32 * 3), ("multiplier", f32), ("haze", f32), ("redblueshift", f32), ("saturation", f32),
                ("horizon_height", f32), ("ground_color", f32 * 3), ("horizon_blur", f32), ("night_color", f32 * 3),
                ("sun_disk_intensity", f32), ("sun_direction", f32 * 3), ("sun_disk_scale", f32), ("sun_glow_intensity", f32),
                ("y_is_up", i32), ("physically_scaled_sun", i32), ("in_use", i32)]


class RenderOpts(C.Structure):
    _fields_ = [("seed_mode", i32), ("frames_in_flight", i32), ("tile_size", i32), ("shard_rank", i32), ("shard_count", i32),
                ("count_traversal", i32), ("sort_rays", i32), ("profile", i32), ("precise_shading", i32), ("tail_from", i32), ("reserved", i32 * 6)]


class Stats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64),
                ("frames_rendered", C.c_uint64), ("kernel_launches", C.c_uint64), ("ms_raygen", C.c_double),
                ("ms_trace_closest", C.c_double), ("ms_trace_shadow", C.c_double), ("ms_shade", C.c_double), ("ms_sort", C.c_double),
                ("ms_accumulate", C.c_double), ("ms_tonemap", C.c_double), ("ms_build", C.c_double), ("num_triangles", u32),
                ("num_bvh_nodes", u32), ("num_instances", u32), ("has_non_opaque", u32), ("shadow_nodes_visited", C.c_uint64),
                ("shadow_tris_tested", C.c_uint64), ("bvh_max_depth", u32), ("node_bytes", u32), ("bvh_width", u32), ("_pad", u32), ("ms_trace_closest_primary", C.c_double),
                ("ms_trace_shadow_primary", C.c_double), ("shadow_rays_primary", C.c_uint64), ("ms_tail", C.c_double)]


# numpy dtypes for the bulk arrays (packed, identical to the C structs)
VertexAttributes = np.dtype([("position", "<f4", 3), ("normal", "<u4"), ("texcoord", "<f4", 2), ("tangent", "<u4"), ("color", "<u4")])

GltfShadeMaterial = np.dtype([
    ("pbrBaseColorFactor", "<f4", 4), ("pbrBaseColorTexture", "<i4"), ("pbrMetallicFactor", "<f4"), ("pbrRoughnessFactor", "<f4"),
    ("pbrMetallicRoughnessTexture", "<i4"), ("emissiveTexture", "<i4"), ("_pad0", "<i4"), ("emissiveFactor", "<f4", 3),
    ("alphaMode", "<i4"), ("alphaCutoff", "<f4"), ("doubleSided", "<i4"), ("normalTexture", "<i4"), ("normalTextureScale", "<f4"),
    ("uvTransform", "<f4", 16), ("unlit", "<i4"), ("transmissionFactor", "<f4"), ("transmissionTexture", "<i4"), ("ior", "<f4"),
    ("anisotropyDirection", "<f4", 3), ("anisotropy", "<f4"), ("attenuationColor", "<f4", 3), ("thicknessFactor", "<f4"),
    ("thicknessTexture", "<i4"), ("attenuationDistance", "<f4"), ("clearcoatFactor", "<f4"), ("clearcoatRoughness", "<f4"),
    ("clearcoatTexture", "<i4"), ("clearcoatRoughnessTexture", "<i4"), ("sheen", "<u4"), ("_pad1", "<i4")])

Light = np.dtype([("direction", "<f4", 3), ("range", "<f4"), ("color", "<f4", 3), ("intensity", "<f4"), ("position", "<f4", 3),
                  ("innerConeCos", "<f4"), ("outerConeCos", "<f4"), ("type", "<i4"), ("padding", "<f4", 2)])

EnvAccel = np.dtype([("alias", "<u4"), ("q", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])

assert C.sizeof(SceneCamera) == 140 and C.sizeof(RtxState) == 48 and C.sizeof(Tonemapper) == 48 and C.sizeof(SunAndSky) == 96
assert VertexAttributes.itemsize == 32 and GltfShadeMaterial.itemsize == 216 and Light.itemsize == 64 and EnvAccel.itemsize == 16
assert GltfShadeMaterial.fields["emissiveFactor"][1] == 40 and GltfShadeMaterial.fields["uvTransform"][1] == 72
assert GltfShadeMaterial.fields["unlit"][1] == 136 and GltfShadeMaterial.fields["anisotropyDirection"][1] == 152
assert GltfShadeMaterial.fields["attenuationColor"][1] == 168 and GltfShadeMaterial.fields["clearcoatFactor"][1] == 192
assert GltfShadeMaterial.fields["sheen"][1] == 208


def default_material(n=1):
    """Defaults of the importer the reference relies on."""
    m = np.zeros(n, dtype=GltfShadeMaterial)
    m["pbrBaseColorFactor"] = (1, 1, 1, 1)
    m["pbrBaseColorTexture"] = -1
    m["pbrMetallicFactor"] = 1.0
    m["pbrRoughnessFactor"] = 1.0
    m["pbrMetallicRoughnessTexture"] = -1
    m["emissiveTexture"] = -1
    m["alphaMode"] = 0
    m["alphaCutoff"] = 0.5
    m["normalTexture"] = -1
    m["normalTextureScale"] = 1.0
    m["uvTransform"] = np.eye(4, dtype=np.float32).reshape(16)
    m["transmissionTexture"] = -1
    m["ior"] = 1.5
    m["anisotropyDirection"] = (0.0, 1.0, 0.0)  # (sin 0, cos 0, 0): src/scene.cpp:365
    m["attenuationColor"] = (1, 1, 1)
    m["thicknessTexture"] = -1
    m["attenuationDistance"] = np.finfo(np.float32).max
    m["clearcoatTexture"] = -1
    m["clearcoatRoughnessTexture"] = -1
    return m


def default_rtx_state(width, height, **kw):
    """src/sample_example.hpp:162-174."""
    s = RtxState(0, 10, 1, 1.0, 1.0, 0, 0, 0, (i32 * 2)(width, height), 0, 65000)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def default_tonemapper(**kw):
    """src/render_output.hpp:37-49."""
    t = Tonemapper(1.0, 1.0, 1.0, 0.0, 1.0, 1.0, (f32 * 2)(1.0, 1.0), 0, 0.5, 0.5, 1)
    for k, v in kw.items():
        setattr(t, k, v)
    return t


def default_sunsky(**kw):
    """src/sample_example.hpp:176-193."""
    s = SunAndSky((f32 * 3)(1, 1, 1), 0.0000101320, 0.0, 0.0, 1.0, 0.0, (f32 * 3)(0.4, 0.4, 0.4), 0.1, (f32 * 3)(0.0, 0.0, 0.01), 0.8,
                  (f32 * 3)(0.0, 0.78, 0.62), 5.0, 1.0, 1, 1, 0)
    for k, v in kw.items():
        setattr(s, k, v)
    return s
