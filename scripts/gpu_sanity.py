"""Ad-hoc GPU bring-up check (not part of the test-suite): traversal + render parity on small scenes, then a timing probe."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vk_raytrace_b200 import Context, scenes, structs as S, upload_scene
from oracle import binding as ob


def rel_l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30)))


def random_rays(n, seed, ext=3.0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-ext, ext, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, :3] = o; rays[:, 3] = 1e32; rays[:, 4:7] = d.astype(np.float32)
    rays[:, 7] = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32)
    return rays


def main():
    g = Context(0)
    for name, sc in (("cornell", scenes.cornell_box(with_light=True)), ("sponza_small", scenes.sponza_standin(detail=0.06, tex_size=64, env_size=(64, 32)))):
        W, H = 96, 64
        o = ob.Oracle()
        upload_scene(g, sc, W, H); upload_scene(o, sc, W, H)
        st_ = g.stats()
        print(name, "tris", st_.num_triangles, "nodes", st_.num_bvh_nodes, "build ms", st_.ms_build, "nonopaque", st_.has_non_opaque, flush=True)
        rays = random_rays(20000, 1, 2.0 if name == "cornell" else 12.0)
        for closest in (True, False):
            a = g.trace_rays(rays, closest); b = o.trace_rays(rays, closest)
            same = np.array_equal(a.view(np.uint32), b.view(np.uint32))
            nhit = int((b[:, 0] < 1e31).sum())
            print(f"  trace closest={closest}: bit-identical={same} hits={nhit}", flush=True)
            if not same:
                bad = np.where((a.view(np.uint32) != b.view(np.uint32)).any(axis=1))[0]
                print("   mismatches", len(bad), a[bad[:5]], b[bad[:5]])
        for pbr in (0, 1):
            for dbg in (0, 1, 2):
                st = S.default_rtx_state(W, H, maxDepth=5 if dbg == 0 else 1, pbrMode=pbr, debugging_mode=dbg)
                g.output_clear(); o.output_create(W, H)
                g.render(st, 3); o.render(st, 3)
                a, b = g.output_read()[..., :3], o.output_read()[..., :3]
                nbad = int((np.abs(a - b) > 1e-4 * (1 + np.abs(b))).any(axis=2).sum())
                print(f"  render pbr={pbr} dbg={dbg}: relL2={rel_l2(a, b):.3e} pixels off={nbad}/{W*H} nan={int(np.isnan(a).sum())}", flush=True)
        s = g.stats(); c = o.counters()
        print("  rays gpu", s.closest_rays, s.shadow_rays, "oracle", c["closest_rays"], c["shadow_rays"], flush=True)
        tm = S.default_tonemapper()
        ta, tb = g.tonemap(tm), o.tonemap(tm)
        print("  tonemap max abs diff (8-bit)", int(np.abs(ta.astype(int) - tb.astype(int)).max()), flush=True)
        g.reset_stats()
    # timing probe
    sc = scenes.sponza_standin(detail=1.0, tex_size=256, env_size=(512, 256))
    W, H = 1920, 1080
    t = time.time(); upload_scene(g, sc, W, H); print("upload+build s", time.time() - t)
    s = g.stats(); print("tris", s.num_triangles, "nodes", s.num_bvh_nodes, "build ms", s.ms_build, flush=True)
    st = S.default_rtx_state(W, H, maxDepth=8)
    opts = S.RenderOpts(); opts.profile = 1
    g.render(st, 4, opts)
    g.reset_stats()
    st.frame = 4
    t = time.time(); g.render(st, 16, opts); dt = time.time() - t
    s = g.stats()
    rays = s.closest_rays + s.shadow_rays
    print(f"16 frames {dt*1e3:.1f} ms  rays {rays/1e6:.1f}M  -> {rays/dt/1e6:.1f} Mrays/s", flush=True)
    print(f"  ms: raygen {s.ms_raygen:.2f} closest {s.ms_trace_closest:.2f} shadow {s.ms_trace_shadow:.2f} shade {s.ms_shade:.2f} accum {s.ms_accumulate:.2f}", flush=True)
    opts.count_traversal = 1
    g.reset_stats(); st.frame = 4; g.render(st, 4, opts); s = g.stats()
    r = s.closest_rays + s.shadow_rays
    print(f"  per ray: nodes {s.nodes_visited/r:.1f} tris {s.tris_tested/r:.1f}", flush=True)


if __name__ == "__main__":
    main()
