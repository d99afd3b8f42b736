import sys, os, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from helpers import small_sponza, rel_l2
from vk_raytrace_b200 import structs as S, upload_scene, Context
from vk_raytrace_b200.context import build_env_accel, make_camera
from vk_raytrace_b200.gltf_export import export_gltf
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "vk_raytrace_b200", "host", "vkrt_gltf_render")
sc = small_sponza(); d = tempfile.mkdtemp()
gltf, env = export_gltf(sc, d, "s")
W, H = 96, 64
_, integral, _ = build_env_accel(sc.env)
def offfrac(a, b): return float((np.abs(a - b) > 2e-3 * (1 + np.abs(b))).any(axis=2).mean())
g = Context(0); upload_scene(g, sc, W, H)
for frames, depth in ((1, 5), (4, 1), (4, 2), (4, 5)):
    out = subprocess.run([exe, gltf, env, str(W), str(H), str(frames), str(depth), d + "/o", "0"], capture_output=True, text=True)
    got = np.fromfile(d + "/o.f32", dtype=np.float32).reshape(H, W, 4)[..., :3]
    st = S.default_rtx_state(W, H, maxDepth=depth, fireflyClampThreshold=integral * 4.0)
    cam = sc.camera
    g.set_camera(make_camera(cam["eye"], cam["center"], cam["up"], cam["fov"], W / H, 0.0, 2))
    g.output_clear(); g.render(st, frames); a = g.output_read()[..., :3]
    # perturbed camera: what the glTF importer reconstructs
    eye, ctr = np.array(cam["eye"], np.float64), np.array(cam["center"], np.float64)
    f = (ctr - eye) / np.linalg.norm(ctr - eye); s = np.cross(f, [0, 1, 0]); s /= np.linalg.norm(s); u = np.cross(s, f)
    g.set_camera(make_camera(eye.astype(np.float32), (eye + f.astype(np.float32) * 7.3).astype(np.float32), u.astype(np.float32), cam["fov"], W / H, 0.0, 2))
    g.output_clear(); g.render(st, frames); b = g.output_read()[..., :3]
    print(f"frames {frames} depth {depth}: gltf-vs-direct {offfrac(got, a):.4f}  perturbedcam-vs-direct {offfrac(b, a):.4f}  gltf-vs-perturbed {offfrac(got, b):.4f}", flush=True)
