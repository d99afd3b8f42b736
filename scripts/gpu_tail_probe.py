"""GPU box: from which bounce should the fused tail kernel take over in *batched* rendering (32 frames in flight, 2 slots)?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vk_raytrace_b200 import Context, structs as S, upload_scene  # noqa: E402

for wl in ("sponza_1080p", "helmet_1080p"):
    cfg = bench.WORKLOADS[wl]
    W, H = cfg["w"], cfg["h"]
    scene = bench.make_scene(cfg)
    ctx = Context(0)
    upload_scene(ctx, scene, W, H, build_mode=4)
    for tail in (-1, 3, 4, 5, 6):
        st = S.default_rtx_state(W, H, maxDepth=cfg["depth"])
        o = S.RenderOpts()
        o.tail_from, o.frames_in_flight = tail, 32
        best = 1e9
        for rep in range(3):
            st.frame = 0
            ctx.render(st, 4, o)
            ctx.reset_stats()
            ctx.mark(0)
            st.frame = 4
            ctx.render(st, 32, o, sync=False)
            ctx.mark(1)
            best = min(best, ctx.elapsed_ms())
        s = ctx.stats()
        rays = s.closest_rays + s.shadow_rays
        print(f"{wl} tail_from {tail:2d}: {best:7.3f} ms / 32 frames  {rays / best / 1e3:8.1f} Mrays/s  launches {s.kernel_launches}", flush=True)
    ctx.close()
