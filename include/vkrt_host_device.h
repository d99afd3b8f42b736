/*
 * vkrt_host_device.h -- byte-exact restatement of the host/device data contract
 * of nvpro-samples/vk_raytrace (reference: shaders/host_device.h:107-282).
 *
 * The reference shares these structs between C++ (glm types, packed) and GLSL
 * (scalar block layout).  Here they are plain C PODs with float arrays so that
 * they compile as C, C++ and CUDA without glm.  Every struct is static-asserted
 * against the reference size.  Matrices are column-major
 * like glm::mat4: m[col*4 + row].
 */
#ifndef VKRT_HOST_DEVICE_H
#define VKRT_HOST_DEVICE_H

#include <stdint.h>

#ifdef __cplusplus
#define VKRT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define VKRT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* DebugMode -- reference shaders/host_device.h:88-102 */
enum VkrtDebugMode {
  VKRT_DBG_NONE      = 0,
  VKRT_DBG_BASECOLOR = 1,
  VKRT_DBG_NORMAL    = 2,
  VKRT_DBG_METALLIC  = 3,
  VKRT_DBG_EMISSIVE  = 4,
  VKRT_DBG_ALPHA     = 5,
  VKRT_DBG_ROUGHNESS = 6,
  VKRT_DBG_TEXCOORD  = 7,
  VKRT_DBG_TANGENT   = 8,
  VKRT_DBG_RADIANCE  = 9,
  VKRT_DBG_WEIGHT    = 10,
  VKRT_DBG_RAYDIR    = 11,
  VKRT_DBG_HEATMAP   = 12  /* per-pixel cost through temperature(): nanoseconds (%globaltimer) the kernels held the path, like clockRealtimeEXT (pathtrace.comp:108-117) */
};

/* alpha modes -- reference shaders/host_device.h:129-131 */
#define VKRT_ALPHA_OPAQUE 0
#define VKRT_ALPHA_MASK 1
#define VKRT_ALPHA_BLEND 2

/* light types -- reference shaders/host_device.h:211-213 */
#define VKRT_LIGHT_DIRECTIONAL 0
#define VKRT_LIGHT_POINT 1
#define VKRT_LIGHT_SPOT 2

/* Inside the reference tree the reference's own shaders/host_device.h (include guard COMMON_HOST_DEVICE, glm types) has
 * already defined these structs: use them as they are -- same names, same bytes -- and only assert the sizes. */
#ifndef COMMON_HOST_DEVICE

#pragma pack(push, 4)

/* reference shaders/host_device.h:107-115 (140 bytes) */
typedef struct SceneCamera {
  float   viewInverse[16];
  float   projInverse[16];
  float   focalDist;
  float   aperture;
  int32_t nbLights;
} SceneCamera;

/* reference shaders/host_device.h:117-124 (32 bytes) */
typedef struct VertexAttributes {
  float    position[3];
  uint32_t normal;      /* oct-compressed */
  float    texcoord[2]; /* tangent handedness in the LSB of texcoord[1] */
  uint32_t tangent;     /* oct-compressed */
  uint32_t color;       /* RGBA8 unorm */
} VertexAttributes;

/* reference shaders/host_device.h:133-179 (216 bytes) */
typedef struct GltfShadeMaterial {
  float    pbrBaseColorFactor[4];       /*   0 */
  int32_t  pbrBaseColorTexture;         /*  16 */
  float    pbrMetallicFactor;           /*  20 */
  float    pbrRoughnessFactor;          /*  24 */
  int32_t  pbrMetallicRoughnessTexture; /*  28 */
  int32_t  emissiveTexture;             /*  32 */
  int32_t  _pad0;                       /*  36 */
  float    emissiveFactor[3];           /*  40 */
  int32_t  alphaMode;                   /*  52 */
  float    alphaCutoff;                 /*  56 */
  int32_t  doubleSided;                 /*  60 */
  int32_t  normalTexture;               /*  64 */
  float    normalTextureScale;          /*  68 */
  float    uvTransform[16];             /*  72 */
  int32_t  unlit;                       /* 136 */
  float    transmissionFactor;          /* 140 */
  int32_t  transmissionTexture;         /* 144 */
  float    ior;                         /* 148 */
  float    anisotropyDirection[3];      /* 152 */
  float    anisotropy;                  /* 164 */
  float    attenuationColor[3];         /* 168 */
  float    thicknessFactor;             /* 180 */
  int32_t  thicknessTexture;            /* 184 */
  float    attenuationDistance;         /* 188 */
  float    clearcoatFactor;             /* 192 */
  float    clearcoatRoughness;          /* 196 */
  int32_t  clearcoatTexture;            /* 200 */
  int32_t  clearcoatRoughnessTexture;   /* 204 */
  uint32_t sheen;                       /* 208 RGBA8: rgb = colour, a = roughness */
  int32_t  _pad1;                       /* 212 */
} GltfShadeMaterial;

/* reference shaders/host_device.h:183-196 (48 bytes, push constant) */
typedef struct RtxState {
  int32_t frame;
  int32_t maxDepth;
  int32_t maxSamples;
  float   fireflyClampThreshold;
  float   hdrMultiplier;
  int32_t debugging_mode;
  int32_t pbrMode; /* 0 Disney, 1 glTF */
  int32_t _pad0;
  int32_t size[2];
  int32_t minHeatmap;
  int32_t maxHeatmap;
} RtxState;

/* reference shaders/host_device.h:215-230 (64 bytes) */
typedef struct Light {
  float   direction[3];
  float   range;
  float   color[3];
  float   intensity;
  float   position[3];
  float   innerConeCos;
  float   outerConeCos;
  int32_t type;
  float   padding[2];
} Light;

/* reference shaders/host_device.h:233-239 (16 bytes) */
typedef struct EnvAccel {
  uint32_t alias;
  float    q;
  float    pdf;
  float    aliasPdf;
} EnvAccel;

/* reference shaders/host_device.h:242-256 (48 bytes) */
typedef struct Tonemapper {
  float   brightness;
  float   contrast;
  float   saturation;
  float   vignette;
  float   avgLum;
  float   zoom;
  float   renderingRatio[2];
  int32_t autoExposure;
  float   Ywhite;
  float   key;
  int32_t dither;
} Tonemapper;

/* reference shaders/host_device.h:259-281 (96 bytes) */
typedef struct SunAndSky {
  float   rgb_unit_conversion[3];
  float   multiplier;
  float   haze;
  float   redblueshift;
  float   saturation;
  float   horizon_height;
  float   ground_color[3];
  float   horizon_blur;
  float   night_color[3];
  float   sun_disk_intensity;
  float   sun_direction[3];
  float   sun_disk_scale;
  float   sun_glow_intensity;
  int32_t y_is_up;
  int32_t physically_scaled_sun;
  int32_t in_use;
} SunAndSky;

#pragma pack(pop)

/* reference shaders/host_device.h:200-205 (24 bytes, 8-byte aligned).
 * In Vulkan these are buffer device addresses; in this framework they are
 * CUDA device pointers filled in by the library at scene commit. */
typedef struct InstanceData {
  uint64_t vertexAddress;
  uint64_t indexAddress;
  int32_t  materialIndex;
  int32_t  _pad;
} InstanceData;

#endif /* !COMMON_HOST_DEVICE */

VKRT_STATIC_ASSERT(sizeof(SceneCamera) == 140, "SceneCamera layout");
VKRT_STATIC_ASSERT(sizeof(VertexAttributes) == 32, "VertexAttributes layout");
VKRT_STATIC_ASSERT(sizeof(GltfShadeMaterial) == 216, "GltfShadeMaterial layout");
VKRT_STATIC_ASSERT(sizeof(RtxState) == 48, "RtxState layout");
VKRT_STATIC_ASSERT(sizeof(InstanceData) == 24, "InstanceData layout");
VKRT_STATIC_ASSERT(sizeof(Light) == 64, "Light layout");
VKRT_STATIC_ASSERT(sizeof(EnvAccel) == 16, "EnvAccel layout");
VKRT_STATIC_ASSERT(sizeof(Tonemapper) == 48, "Tonemapper layout");
VKRT_STATIC_ASSERT(sizeof(SunAndSky) == 96, "SunAndSky layout");

#endif /* VKRT_HOST_DEVICE_H */
