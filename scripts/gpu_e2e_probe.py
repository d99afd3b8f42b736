"""GPU box: where does an interactive frame (one vkrt_render call per frame) spend its time?  ms per frame of the sponza_1080p workload
for several depths, with and without the tonemap / read-back, single and dual wavefront slots (VKRT_DUAL_SLOT is read per process: run twice)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vk_raytrace_b200 import Context, structs as S, upload_scene  # noqa: E402

cfg = bench.WORKLOADS["sponza_1080p"]
W, H = cfg["w"], cfg["h"]
scene = bench.make_scene(cfg)
ctx = Context(0)
upload_scene(ctx, scene, W, H, build_mode=2)
tm = S.default_tonemapper()
bufs = [ctx.host_alloc((H, W, 4)), ctx.host_alloc((H, W, 4))]


def loop(depth, frames, tonemap, batch=1):
    st = S.default_rtx_state(W, H, maxDepth=depth)
    for rep in range(2):
        ctx.reset_stats()
        ctx.sync()
        t0 = time.time()
        for f in range(frames):
            st.frame = f * batch
            ctx.render(st, batch, None, sync=False)
            if tonemap:
                ctx.tonemap_wait(f & 1)
                ctx.tonemap_async(tm, bufs[f & 1], f & 1)
        ctx.tonemap_wait(0)
        ctx.tonemap_wait(1)
        ctx.sync()
        dt = time.time() - t0
    s = ctx.stats()
    rays = s.closest_rays + s.shadow_rays
    return dt / (frames * batch) * 1e3, rays / dt / 1e6, s.kernel_launches / frames


print("dual_slot env:", os.environ.get("VKRT_DUAL_SLOT", "default(1)"))
for depth in (1, 2, 3, 4, 8):
    for tonemap in (False, True):
        ms, mr, launches = loop(depth, 24, tonemap)
        print(f"depth {depth} tonemap {int(tonemap)}: {ms:.3f} ms/frame  {mr:.0f} Mrays/s  {launches:.0f} launches/frame", flush=True)
for batch in (2, 4, 8):
    ms, mr, launches = loop(8, 8, False, batch)
    print(f"depth 8 batch {batch}: {ms:.3f} ms/frame {mr:.0f} Mrays/s", flush=True)
