"""GPU box (1 GPU): what a rank of an N-GPU strong-scaling run does -- 20 frames of 1/N of the 1080p pixels -- timed alone, for the
scheduling knobs: wavefront slots (VKRT_DUAL_SLOT, per process), fused tail, frames per batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vk_raytrace_b200 import Context, structs as S, upload_scene  # noqa: E402

cfg = bench.WORKLOADS["sponza_1080p"]
W, H = cfg["w"], cfg["h"]
scene = bench.make_scene(cfg)
ctx = Context(0)
upload_scene(ctx, scene, W, H, build_mode=4)
print("dual_slot env:", os.environ.get("VKRT_DUAL_SLOT", "default(1)"), "slots env:", os.environ.get("VKRT_SLOTS", "default(2)"))
for shards in [int(x) for x in os.environ.get("SHARD_PROBE_SHARDS", "8,4,1").split(",")]:
    for tail in [int(x) for x in os.environ.get("SHARD_PROBE_TAILS", "0,-1,2").split(",")]:
        for fif in [int(x) for x in os.environ.get("SHARD_PROBE_FIFS", "20,10,40").split(",")]:
            st = S.default_rtx_state(W, H, maxDepth=8)
            o = S.RenderOpts()
            o.shard_rank, o.shard_count, o.tile_size, o.tail_from, o.frames_in_flight = 0, shards, 64, tail, fif
            best = 1e9
            for rep in range(3):
                st.frame = 0
                ctx.render(st, 3, o)
                ctx.reset_stats()
                ctx.mark(0)
                st.frame = 3
                ctx.render(st, 20, o, sync=False)
                ctx.mark(1)
                best = min(best, ctx.elapsed_ms())
            s = ctx.stats()
            rays = s.closest_rays + s.shadow_rays
            print(f"shards {shards} tail_from {tail:2d} frames_in_flight {fif:2d}: {best:7.3f} ms / 20 frames  {rays / best / 1e3:8.1f} Mrays/s  launches {s.kernel_launches}", flush=True)
