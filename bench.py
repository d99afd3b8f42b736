#!/usr/bin/env python
"""bench.py -- Mrays/s of the wavefront path tracer on the BASELINE.json workload.

A *step* is one frame (1 sample per pixel, all bounces) of the workload image:
    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]
N = 1: the `sponza_1080p` workload (configs[2] of BASELINE.json: Sponza stand-in, ~262k triangles,
1920x1080, depth 8, Disney BSDF, HDR environment) -- the configuration the metric is quoted on.
N > 1 (under torchrun): the same image, tiles dealt round-robin to the ranks (strong scaling), one NCCL
reduce of the RGBA32F accumulation image at the end of the timed region.

Prints ONE JSON line (rank 0).  value = rays traced by all ranks (ClosestHit + AnyHit calls, device
counters) / device time (CUDA events on the library stream, max over ranks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene factory kwargs, width, height, maxDepth)
    "sponza_1080p": dict(scene="sponza", detail=1.0, tex=512, env=(1024, 512), w=1920, h=1080, depth=8),
    "sponza_4k": dict(scene="sponza", detail=1.0, tex=512, env=(1024, 512), w=3840, h=2160, depth=8),
    "helmet_1080p": dict(scene="helmet", level=6, tex=2048, env=(2048, 1024), w=1920, h=1080, depth=6),
    "instanced_10m": dict(scene="instanced", n=10, w=1920, h=1080, depth=8),
    "cornell_256": dict(scene="cornell", w=256, h=256, depth=1),
    "sponza_small": dict(scene="sponza", detail=0.1, tex=128, env=(256, 128), w=640, h=360, depth=8),
}


def make_scene(cfg):
    from vk_raytrace_b200 import scenes
    k = cfg["scene"]
    if k == "sponza":
        return scenes.sponza_standin(detail=cfg["detail"], tex_size=cfg["tex"], env_size=cfg["env"])
    if k == "helmet":
        return scenes.helmet_standin(level=cfg["level"], tex_size=cfg["tex"], env_size=cfg["env"])
    if k == "instanced":
        return scenes.instanced_stress(instances_per_axis=cfg["n"])
    return scenes.cornell_box()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None

    def _run_nvml(self):
        """NVML polling every 10 ms (the timed region of the default run is ~0.1 s; nvidia-smi takes that long per call)."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        try:
            reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons
        except AttributeError:
            reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            mask = int(reasons(h))
            for n, b in bits.items():
                if mask & b:
                    self.reasons.add(n)
            self.stop_flag.wait(0.01)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def algorithmic_bytes(rays, nodes, tris, node_bytes=64):
    """per ray 32 B ray + 16 B hit + node_bytes per node + 48 B per triangle (single-level BVH: no instance term).
    node_bytes is what the traversal kernel actually fetches per node: 32 (16-bit grid boxes, default) or 64 (fp32 boxes)."""
    return 48 * rays + node_bytes * nodes + 48 * tris


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference(cfg, steps, warmup, threads=0, budget_s=20.0):
    """The reference has no CPU path and its Vulkan paths cannot run here; the timed CPU arm is the oracle port
    (oracle/, std::thread over scanlines) on a bounded sample of the same workload: a centred band of rows."""
    from vk_raytrace_b200 import structs as S, upload_scene
    from oracle import binding as ob
    scene = make_scene(cfg)
    W, H = cfg["w"], cfg["h"]
    o = ob.Oracle(threads=threads)
    upload_scene(o, scene, W, H)
    st = S.default_rtx_state(W, H, maxDepth=cfg["depth"])
    cores = threads or os.cpu_count()
    # calibrate the band so that the whole run stays bounded
    rows = max(2, H // 90)
    r0 = (H - rows) // 2
    t = time.time()
    o.render(st, 1, rows=(r0, r0 + rows))
    per_row = (time.time() - t) / rows
    total = max(1, steps + warmup)
    rows = int(max(2, min(H, budget_s / max(per_row, 1e-6) / total)))
    r0 = (H - rows) // 2
    for f in range(warmup):
        st.frame = f
        o.render(st, 1, rows=(r0, r0 + rows))
    o.reset_counters()
    t = time.time()
    for f in range(steps):
        st.frame = warmup + f
        o.render(st, 1, rows=(r0, r0 + rows))
    dt = time.time() - t
    c = o.counters()
    rays = c["closest_rays"] + c["shadow_rays"]
    return dict(value=rays / dt / 1e6, unit="Mrays/s", cores=cores, kind="port",
                sample=f"{steps} frames x rows [{r0},{r0 + rows}) of {W}x{H} ({rays} rays in {dt:.2f} s, oracle/liboracle.so, {cores} threads)"), dt / max(steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sponza_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--pbr-mode", type=int, default=0)
    ap.add_argument("--build-mode", type=int, default=-1,
                    help="-1 auto (default): the best-tracing builder that is still quick at load time, as the reference asks its driver for "
                         "(PREFER_FAST_TRACE, src/accelstruct.cpp:125): host binned SAH up to 2 M triangles, GPU LBVH + leaf collapse above; "
                         "0 GPU LBVH, 1 GPU LBVH + SAH leaf collapse, 2 host binned SAH, 3 GPU PLOC + SAH leaf collapse")
    ap.add_argument("--frames-in-flight", type=int, default=0)
    ap.add_argument("--sort-rays", action="store_true", help="vkrt_render_opts.sort_rays: order each bounce's queue by octant + origin Morton key")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = WORKLOADS[args.workload]
    W, H = cfg["w"], cfg["h"]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        cb, s_per_step = cpu_reference(cfg, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scene": "procedural Sponza stand-in" if cfg["scene"] == "sponza" else cfg["scene"],
                           "resolution": [W, H], "max_depth": cfg["depth"], "spp_per_step": 1,
                           "note": "reference Vulkan RTX / ray-query paths not runnable on this box (no Vulkan ICD, glslang, nvpro_core); "
                                   "CPU oracle port timed instead"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import numpy as np
    import torch
    from vk_raytrace_b200 import Context, structs as S, upload_scene

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if use_dist else 0
    torch.cuda.set_device(device)

    scene = make_scene(cfg)
    if args.build_mode < 0:
        args.build_mode = 2 if scene.num_triangles() <= 2_000_000 else 1
    ctx = Context(device)
    upload_scene(ctx, scene, W, H, build_mode=args.build_mode)
    # the accumulation image lives in a torch tensor so that torch.distributed can reduce it over NVLink
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=f"cuda:{device}")
    ctx.output_create(W, H, accum.data_ptr())
    _, integral, _ = ctx.build_env_accel(scene.env)
    st = S.default_rtx_state(W, H, maxDepth=cfg["depth"], pbrMode=args.pbr_mode, fireflyClampThreshold=integral * 4.0)  # src/sample_example.cpp:110
    opts = S.RenderOpts()
    opts.shard_rank, opts.shard_count = (rank, world) if use_dist else (0, 1)
    opts.frames_in_flight = args.frames_in_flight
    opts.sort_rays = 1 if args.sort_rays else 0
    opts.profile = 0    # timed region: no per-kernel events, so shadow(b) overlaps closest(b+1) on the library's second stream

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (untimed)
    st.frame = 0
    ctx.render(st, args.warmup, opts)
    if use_dist:
        dist.reduce(accum.clone(), 0)   # warm the communicator without touching the image
    barrier()
    ctx.reset_stats()

    sampler = ClockSampler(device)
    sampler.start()
    barrier()
    t_wall = time.time()
    ctx.mark(0)
    st.frame = args.warmup
    ctx.render(st, args.steps, opts, sync=False)
    ctx.mark(1)
    ms_render = ctx.elapsed_ms()
    ms_reduce = 0.0
    if use_dist:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce(accum, 0)       # disjoint tiles + zeros: the sum is an exact gather
        e1.record()
        torch.cuda.synchronize()
        ms_reduce = e0.elapsed_time(e1)
    barrier()
    wall_ms = (time.time() - t_wall) * 1e3
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    s = ctx.stats()
    my_ms = ms_render + ms_reduce
    rays = int(s.closest_rays + s.shadow_rays)
    if use_dist:
        t = torch.tensor([my_ms], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r = torch.tensor([rays], dtype=torch.int64, device=f"cuda:{device}")
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        total_ms, total_rays = float(t.item()), int(r.item())
    else:
        total_ms, total_rays = my_ms, rays
    launches = int(s.kernel_launches)
    # per-kernel device time: the same frames again with per-kernel CUDA events (kernels serialised, untimed pass)
    opts_p = S.RenderOpts()
    opts_p.shard_rank, opts_p.shard_count, opts_p.profile = opts.shard_rank, opts.shard_count, 1
    opts_p.frames_in_flight = args.frames_in_flight
    opts_p.sort_rays = opts.sort_rays
    ctx.reset_stats()
    st.frame = args.warmup
    ctx.render(st, args.steps, opts_p)
    s = ctx.stats()
    kernel_ms = dict(raygen=s.ms_raygen, trace_closest=s.ms_trace_closest, trace_shadow=s.ms_trace_shadow, shade=s.ms_shade,
                     accumulate=s.ms_accumulate, sort=s.ms_sort)

    # ---- roofline of the dominant kernel: instrumented re-run of the same frames (untimed) gives the visit counts
    peak, peak_src = measured_peaks()
    opts_c = S.RenderOpts()
    opts_c.shard_rank, opts_c.shard_count, opts_c.count_traversal = opts.shard_rank, opts.shard_count, 1
    opts_c.frames_in_flight = args.frames_in_flight
    opts_c.sort_rays = opts.sort_rays
    ctx.reset_stats()
    st.frame = args.warmup
    ctx.render(st, args.steps, opts_c)
    sc = ctx.stats()
    c_nodes, c_tris = sc.nodes_visited - sc.shadow_nodes_visited, sc.tris_tested - sc.shadow_tris_tested
    node_bytes = int(sc.node_bytes) or 64
    bytes_closest = algorithmic_bytes(sc.closest_rays, c_nodes, c_tris, node_bytes)
    bytes_shadow = algorithmic_bytes(sc.shadow_rays, sc.shadow_nodes_visited, sc.shadow_tris_tested, node_bytes)
    ach_closest = bytes_closest / max(kernel_ms["trace_closest"], 1e-9) / 1e6   # GB/s
    ach_shadow = bytes_shadow / max(kernel_ms["trace_shadow"], 1e-9) / 1e6
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof) and args.workload == "sponza_1080p":   # the capture is of this workload's bounce-0 launch
        try:
            traffic = json.load(open(prof)).get("traceClosestKernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "traceKernel<CLOSEST>", "bound": "hbm", "achieved": ach_closest, "peak": peak, "unit": "GB/s", "frac": ach_closest / peak,
                "traffic": traffic, "peak_source": peak_src,
                "per_ray": {"node_bytes": node_bytes, "nodes": c_nodes / max(1, sc.closest_rays), "tris": c_tris / max(1, sc.closest_rays),
                            "bytes": bytes_closest / max(1, sc.closest_rays)},
                "shadow_kernel": {"achieved": ach_shadow, "frac": ach_shadow / peak,
                                  "nodes_per_ray": sc.shadow_nodes_visited / max(1, sc.shadow_rays),
                                  "tris_per_ray": sc.shadow_tris_tested / max(1, sc.shadow_rays)},
                "note": "algorithmic bytes (SURVEY 8d) / summed CUDA-event kernel time of the serialised pass.  For L2-resident scenes "
                        "(this one: 30 MB BVH) the algorithmic rate exceeds the HBM peak because node traffic is served by L1/L2; DRAM "
                        "traffic (`traffic`, ncu, bounce-0 launch) is the ray / hit streams only.  What bounds the kernel is the latency of "
                        "the dependent node fetches at 32 warps per SM (profiles/README.md); the HBM-resident case is instanced_10m."}

    # ---- e2e: host buffers in, host image out, every step through the C ABI (rank-local; N=1 only)
    e2e = None
    if not args.no_e2e and not use_dist:
        cam = scene.camera
        tm = S.default_tonemapper()
        out8 = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()   # pinned host frame buffer
        opts_e = S.RenderOpts()
        k = max(4, min(args.steps, 16))
        for f in range(2):
            st.frame = f
            ctx.render(st, 1, opts_e)
            ctx.tonemap(tm, out8)
        ctx.reset_stats()
        torch.cuda.synchronize()
        t0 = time.time()
        for f in range(k):
            camera = ctx.make_camera(cam["eye"], cam["center"], cam["up"], cam["fov"], W / H, 0.0, len(scene.lights) if scene.lights is not None else 0)
            ctx.set_camera(camera)          # H2D: SceneCamera (140 B) + RtxState (48 B) travel as kernel parameters
            st.frame = 2 + f
            ctx.render(st, 1, opts_e, sync=False)
            ctx.tonemap(tm, out8)           # D2H: the displayed RGBA8 frame
        dt = time.time() - t0
        se = ctx.stats()
        e2e = {"value": (se.closest_rays + se.shadow_rays) / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 140 + 48,
               "d2h_bytes_per_step": W * H * 4, "steps": k, "ms_per_step": dt / k * 1e3,
               "what": "per step: vkrt_set_camera(host SceneCamera) + vkrt_render(1 frame) + vkrt_tonemap -> pinned host RGBA8 "
                       "(interactive use: one frame per call, so no multi-frame batching)"}
    elif use_dist:
        # per step host traffic is the same tiny parameter upload; the image comes back once after the reduce
        t0 = time.time()
        host = accum.cpu() if rank == 0 else None
        torch.cuda.synchronize()
        d2h_ms = (time.time() - t0) * 1e3
        e2e = {"value": total_rays / ((total_ms + d2h_ms) * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 140 + 48,
               "d2h_bytes_per_step": int(W * H * 16 / max(1, args.steps)),
               "what": "timed region + D2H of the reduced RGBA32F image on rank 0 (amortised over the steps)"}
        del host

    cb = None
    if rank == 0 and not args.no_cpu_baseline and not use_dist:
        cb, _ = cpu_reference(cfg, 8, 1, budget_s=25.0)   # ~10-25 s of host work, bounded by shrinking the band of rows

    if rank == 0:
        line = {"metric": "Mrays/s", "value": total_rays / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scene": scene.name, "triangles": int(s.num_triangles), "bvh_nodes": int(s.num_bvh_nodes),
                           "bvh_build": {"mode": {0: "lbvh", 1: "lbvh+leaf-collapse", 2: "host-sah", 3: "ploc+leaf-collapse"}[args.build_mode],
                                         "ms": float(s.ms_build), "bvh_width": int(s.bvh_width), "max_depth": int(s.bvh_max_depth),
                                         "what": "vkrt_scene_commit: instance flattening + BVH build, outside the timed region (SURVEY 8d)"},
                           "resolution": [W, H], "max_depth": cfg["depth"], "spp_per_step": 1, "pbr_mode": args.pbr_mode,
                           "sharding": f"{world} x interleaved 64x64 tiles" if use_dist else "none",
                           "l2": "per-step wavefront state (~170 B x %d paths) exceeds the 126 MB L2; the BVH/geometry is L2-resident by design" % (W * H),
                           "rays_counted": "ClosestHit + AnyHit calls (device counters)"},
                "rays": total_rays, "gpu_launches": launches, "kernel_ms": kernel_ms, "reduce_ms": ms_reduce, "wall_ms": wall_ms,
                "clocks": sampler.summary(), "roofline": roofline, "e2e": e2e}
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
