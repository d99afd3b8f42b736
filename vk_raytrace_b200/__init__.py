"""vk_raytrace_b200 -- B200-native wavefront path tracer behind the vk_raytrace Renderer interface.

The product is the CUDA library ``libvkrt_b200.so`` (csrc/, C ABI in include/vkrt_b200.h).
This package is the thin Python harness used by tests and bench.py: ctypes bindings
(`lib`), a convenience wrapper (`Context`) and procedural scene generators (`scenes`).
"""
from . import structs  # noqa: F401
from .lib import load_library, library_path, LibraryMissing  # noqa: F401
from .context import Context, SceneDesc, MeshDesc, TextureDesc, upload_scene  # noqa: F401
