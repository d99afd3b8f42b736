"""GPU box: whole-image per-channel relative L2 between the CUDA path and the oracle as a function of the number of accumulated
frames (no outlier mask), for both shade kernels -- the measurement behind tests/test_gpu_parity.py::test_converged_whole_image."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import small_sponza  # noqa: E402
from oracle import binding as ob  # noqa: E402
from vk_raytrace_b200 import Context, scenes, structs as S, upload_scene  # noqa: E402


def pcl2(a, b):
    return max(float(np.sqrt(((a[..., c] - b[..., c]) ** 2).sum() / max((b[..., c] ** 2).sum(), 1e-300))) for c in range(3))


out = {}
gpu = Context(0)
for name, sc, W, H, depth in (("cornell", scenes.cornell_box(with_light=True), 48, 48, 4), ("small_sponza", small_sponza(), 64, 40, 6)):
    for pbr in (0, 1):
        o = ob.Oracle()
        upload_scene(o, sc, W, H, packer=ob)
        refs, done = {}, 0
        checkpoints = (16, 64, 256, 1024, 4096)
        for n in checkpoints:
            st = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=pbr, frame=done)
            o.render(st, n - done)
            done = n
            refs[n] = o.output_read()[..., :3].astype(np.float64)
        for precise in (True, False):
            upload_scene(gpu, sc, W, H, packer=ob)
            done = 0
            for n in checkpoints:
                st = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=pbr, frame=done)
                gpu.render(st, n - done, precise=precise)
                done = n
                a = gpu.output_read()[..., :3].astype(np.float64)
                key = f"{name} pbr{pbr} {'precise' if precise else 'fast'} N={n}"
                out[key] = pcl2(a, refs[n])
                print(key, f"{out[key]:.3e}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/converge_r02.json", "w"), indent=1)
