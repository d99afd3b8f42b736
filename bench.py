#!/usr/bin/env python
"""bench.py -- Mrays/s of the wavefront path tracer on the BASELINE.json workload.

A *step* is one frame (1 sample per pixel, all bounces) of the workload image:
    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]
N = 1: the `sponza_1080p` workload (configs[2] of BASELINE.json: Sponza stand-in, ~262k triangles,
1920x1080, depth 8, Disney BSDF, HDR environment) -- the configuration the metric is quoted on.
N > 1 (under torchrun): the same image, tiles dealt round-robin to the ranks (strong scaling).  Rank 0 gets the other
ranks' pixels inside the timed region: their accumulate kernels store them into its image over NVLink (CUDA IPC peer
mapping; class Exchange), and the run checks that image against an unsharded render (`config.exchange_verified`).

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


def default_build_mode(num_triangles):
    """The builder bench.py uses when --build-mode is not given (also what tests/test_gpu_parity.py validates against the oracle): the GPU
    binned-SAH builder, the analogue of the PREFER_FAST_TRACE build the reference requests from its driver (src/accelstruct.cpp:125).
    Round 1 timed a tree built by one CPU thread (mode 2); mode 4 traces the same (40.15 nodes + 5.90 triangles per closest ray on the
    headline scene, identical counts) and builds on the GPU."""
    return 4


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None

    _BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def prepare(self):
        """NVML handle ahead of the timed region (nvmlInit takes longer than a multi-GPU timed region lasts)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            try:
                self._reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons
            except AttributeError:
                self._reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            self._h = None

    def sample_now(self):
        """One sample from the calling thread: called right after the timed work has been enqueued, while the GPU executes it (the
        polling thread alone gets no sample out of a 5 ms region)."""
        if getattr(self, "_h", None) is None:
            return
        try:
            self.samples.append(float(self._nvml.nvmlDeviceGetClockInfo(self._h, self._nvml.NVML_CLOCK_SM)))
            mask = int(self._reasons(self._h))
            for n, b in self._BITS.items():
                if mask & b:
                    self.reasons.add(n)
        except Exception:
            pass

    def _run_nvml(self):
        """NVML polling every 10 ms (the timed region of the default run is ~0.1 s; nvidia-smi takes that long per call)."""
        if not hasattr(self, "_h"):
            self.prepare()
        if self._h is None:
            raise RuntimeError("no NVML")
        while not self.stop_flag.is_set():
            self.sample_now()
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


def host_cores():
    """Threads the CPU arm may really use on this box: the scheduler affinity of this process, capped by the cgroup CPU quota
    (os.cpu_count() reports the machine, not the container: round 1's CPU arm swung 5.8x between two boxes with '128 cores')."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                if q > 0:
                    quota = q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            break
        except Exception:
            continue
    if quota:
        n = max(1, min(n, int(quota + 0.5)))
    return n


def vulkan_probe():
    """is the reference's own Vulkan path runnable on this box?  Looked up at run time, not assumed."""
    import glob
    import shutil
    icds = sorted(glob.glob("/usr/share/vulkan/icd.d/*.json") + glob.glob("/etc/vulkan/icd.d/*.json"))
    info = shutil.which("vulkaninfo")
    summary = None
    if info:
        try:
            r = subprocess.run([info, "--summary"], capture_output=True, text=True, timeout=20)
            summary = (r.stdout or r.stderr).strip().splitlines()[:3]
        except Exception as e:   # noqa: BLE001
            summary = [f"vulkaninfo failed: {e}"]
    glslang = shutil.which("glslangValidator") or shutil.which("glslc")
    runnable = bool(icds) and info is not None and glslang is not None
    return {"icd_files": icds, "vulkaninfo": info, "vulkaninfo_summary": summary, "glsl_compiler": glslang,
            "driver_capabilities": os.environ.get("NVIDIA_DRIVER_CAPABILITIES"), "reference_vulkan_runnable": runnable,
            "note": "the reference needs a Vulkan ICD with VK_KHR_ray_query, glslang and the un-vendored nvpro_core; "
                    + ("present" if runnable else "not all present on this box, so its RTX / ray-query paths cannot be timed here")}


def cpu_reference(cfg, steps, warmup, threads=0, budget_s=20.0):
    """The reference has no CPU path and its Vulkan paths cannot run here; the timed CPU arm is the oracle port
    (oracle/, std::thread over scanlines) on a bounded sample of the same workload: a centred band of rows."""
    from vk_raytrace_b200 import structs as S, upload_scene
    from oracle import binding as ob
    scene = make_scene(cfg)
    W, H = cfg["w"], cfg["h"]
    cores = threads or host_cores()
    o = ob.Oracle(threads=cores)
    upload_scene(o, scene, W, H, packer=ob)      # the oracle's own vertex packer and camera: no product code on this arm
    st = S.default_rtx_state(W, H, maxDepth=cfg["depth"])
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
                    help="-1 (default): the GPU binned-SAH builder (4), the PREFER_FAST_TRACE analogue (src/accelstruct.cpp:125); "
                         "0 GPU LBVH, 1 GPU LBVH + SAH leaf collapse, 2 host binned SAH, 3 GPU PLOC + SAH leaf collapse, 4 GPU binned SAH")
    ap.add_argument("--frames-in-flight", type=int, default=0)
    ap.add_argument("--sort-rays", action="store_true", help="vkrt_render_opts.sort_rays: order each bounce's queue by octant + origin Morton key")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", choices=["peer", "gather"], default="peer", help="N > 1: how rank 0 gets the other ranks' pixels (see class Exchange)")
    ap.add_argument("--no-verify-exchange", action="store_true", help="N > 1: skip comparing the exchanged image with an unsharded render on rank 0")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hbm-record", action="store_true", help="skip the HBM-resident roofline record (instanced_10m sub-run, N=1)")
    ap.add_argument("--scene-file", default=None, help="render a glTF 2.0 file (.gltf / .glb: the real Sponza / DamagedHelmet when present) through the "
                    "C++ ingest (libvkrt_host.so) instead of a procedural workload; resolution / depth from --width --height --depth")
    ap.add_argument("--env-file", default=None, help="Radiance .hdr or .pfm environment for --scene-file (default: the procedural sky)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--tile", type=int, default=0, help="edge of the sharding tiles in pixels (default 64)")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the 3840x2160 sub-record (BASELINE.json configs[3])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = WORKLOADS[args.workload]
    if args.scene_file:
        cfg = dict(scene="file", w=args.width, h=args.height, depth=args.depth)
        args.workload = "file:" + os.path.basename(args.scene_file)
        args.no_cpu_baseline = True      # the CPU oracle has no glTF ingest of its own: it is fed from in-memory scenes only
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the CPU oracle has no glTF file ingest; --scene-file runs are GPU-arm only"}))
            return
    W, H = cfg["w"], cfg["h"]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        cb, s_per_step = cpu_reference(cfg, args.steps, args.warmup)
        vk_probe = vulkan_probe()
        line = {"impl": "reference", "metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scene": "procedural Sponza stand-in" if cfg["scene"] == "sponza" else cfg["scene"],
                           "resolution": [W, H], "max_depth": cfg["depth"], "spp_per_step": 1,
                           "vulkan": vk_probe,
                           "note": "the reference has no CPU path; its Vulkan RTX / ray-query paths are probed above; the CPU oracle port "
                                   "(oracle/, fed by its own packer and camera) is what is timed"},
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
    dev = f"cuda:{device}"

    ctx = Context(device)
    if args.scene_file:
        # Scene::load through the C++ ingest (vk_raytrace_b200/host: glTF + PNG / JPEG), environment from a file or the procedural sky
        from types import SimpleNamespace
        from vk_raytrace_b200 import scenes as _scenes
        if args.build_mode < 0:
            args.build_mode = default_build_mode(0)
        info = ctx.load_gltf(args.scene_file, args.build_mode)
        cold_build_ms = float(ctx.stats().ms_build)
        if args.env_file:
            integral, _, _ = ctx.load_environment(args.env_file)
        else:
            env = _scenes.sky_environment(1024, 512, 100)
            _, integral, _ = ctx.build_env_accel(env)
            ctx.set_environment(env)
        scene = SimpleNamespace(name=args.workload, env=None, lights=[None] * int(info.lights),
                                camera=dict(eye=tuple(info.eye), center=tuple(info.center), up=tuple(info.up), fov=float(info.yfov_degrees)),
                                num_triangles=lambda: int(info.triangles))
        ctx.set_camera(ctx.make_camera(scene.camera["eye"], scene.camera["center"], scene.camera["up"], scene.camera["fov"], W / H, 0.0, int(info.lights)))
        ctx.output_create(W, H)
    else:
        scene = make_scene(cfg)
        if args.build_mode < 0:
            args.build_mode = default_build_mode(scene.num_triangles())
        upload_scene(ctx, scene, W, H, build_mode=args.build_mode)
        cold_build_ms = float(ctx.stats().ms_build)
        ctx.commit(args.build_mode)            # again, warm: the first build pays CUDA's lazy kernel loading and first allocations
        _, integral, _ = ctx.build_env_accel(scene.env)
    TILE = args.tile if args.tile > 0 else 64
    shard = (rank, world) if use_dist else (0, 1)

    def make_opts(**kw):
        o = S.RenderOpts()
        o.shard_rank, o.shard_count, o.tile_size = shard[0], shard[1], TILE
        o.sort_rays = 1 if args.sort_rays else 0
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    class Exchange:
        """Image exchange of a sharded render, two ways.
        "peer" (default): the exchange is fused into the accumulate kernel -- every non-root rank maps rank 0's accumulation image
        over CUDA IPC (vkrt_output_export / vkrt_peer_open) and its accumulate kernel stores each pixel it owns there as well, over
        NVLink, while it runs (vkrt_output_set_mirror).  What is left is a 4-byte all-reduce between two vkrt_output_fence calls:
        a cross-process fence, no pack / gather / unpack launches and no staging buffers.
        "gather" (--exchange gather, or when the IPC mapping is refused): every rank packs the pixels it owns (RGBA32F, 1/N of the
        image) and rank 0 gathers and scatters them.  Both give the exact image of a single-GPU render."""

        def __init__(self, w, h):
            self.n = [ctx.shard_size(TILE, r, world) for r in range(world)]
            self.cap = max(self.n)
            self.bytes = self.cap * 16
            self.mode, self.peer_ptr, self.why = args.exchange, None, ""
            if self.mode == "peer":
                ok = 1
                try:
                    hs = [ctx.output_export() if rank == 0 else None]
                except Exception as e:           # the broadcast below must still be entered by every rank
                    hs, ok, self.why = [None], 0, str(e)
                dist.broadcast_object_list(hs, src=0)
                if rank != 0:
                    try:
                        self.peer_ptr = ctx.peer_open(hs[0])
                        ctx.set_mirror(self.peer_ptr)
                    except Exception as e:
                        ok, self.why = 0, str(e)
                t = torch.tensor([ok], dtype=torch.int32, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 0:
                    self.close()
                    self.mode = "gather"
            self.token = torch.zeros(1, dtype=torch.float32, device=dev)
            if self.mode == "gather":
                self.mine = torch.zeros((self.cap, 4), dtype=torch.float32, device=dev)
                self.all = [torch.zeros((self.cap, 4), dtype=torch.float32, device=dev) for _ in range(world)] if rank == 0 else None

        def _fence(self):
            ctx.output_fence()              # the collective waits for this rank's image operations (accumulate: its peer stores included)
            dist.all_reduce(self.token)     # 4 bytes: every rank has got this far
            ctx.output_fence()              # image operations enqueued from here on wait for the collective

        def run(self):
            """After a render: the root's image holds every rank's pixels once this is through.  Everything is enqueued on the
            library's image stream (torch.cuda.ExternalStream); no host synchronisation."""
            with torch.cuda.stream(lib_stream):
                if self.mode == "peer":
                    self._fence()
                    return
                ctx.pack_shard(TILE, rank, world, self.mine.data_ptr())
                dist.gather(self.mine, self.all, dst=0)
                if rank == 0:
                    for r in range(1, world):               # rank 0's own pixels are already in place
                        ctx.unpack_shard(TILE, r, world, self.all[r].data_ptr())

        def release(self):
            """After the root has enqueued its use of the frame (tonemap, read): the other ranks' next accumulate must not store
            into the root's image before that use is through.  The gather path orders this on the root alone."""
            if self.mode == "peer":
                with torch.cuda.stream(lib_stream):
                    self._fence()

        def close(self):
            if self.peer_ptr is not None:
                ctx.set_mirror(None)
                ctx.peer_close(self.peer_ptr)    # synchronises this context first
                self.peer_ptr = None

    live = {"ex": None}

    def measure(width, height, steps, warmup, fif):
        """One timed region at the given resolution: `warmup` untimed frames, then `steps` frames between two stream barriers of the
        library (CUDA events), plus the shard exchange when sharded.  Returns per-rank numbers."""
        if live["ex"] is not None:           # mappings of the previous image go before rank 0 frees it
            live["ex"].close()
            live["ex"] = None
            barrier()
        if use_dist:
            accum = None
            ctx.output_create(width, height)     # library-owned: rank 0's can be exported to the other processes
        else:
            accum = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)   # caller-owned buffer (a torch tensor)
            ctx.output_create(width, height, accum.data_ptr())
        ctx.set_camera(ctx.make_camera(scene.camera["eye"], scene.camera["center"], scene.camera["up"], scene.camera["fov"], width / height,
                                       scene.camera.get("aperture", 0.0), 0 if scene.lights is None else len(scene.lights)))
        st = S.default_rtx_state(width, height, maxDepth=cfg["depth"], pbrMode=args.pbr_mode, fireflyClampThreshold=integral * 4.0)  # src/sample_example.cpp:110
        n_local = ctx.shard_size(TILE, shard[0], shard[1])
        if fif <= 0:   # the library's automatic rule, made explicit so that the warm-up call already reserves the wavefront of the timed one
            fif = max(1, min(64, (64 << 20) // max(n_local, 1), steps))
        opts = make_opts(frames_in_flight=fif, profile=0)   # no per-kernel events: the second stream and the second slot overlap
        ex = Exchange(width, height) if use_dist else None
        live["ex"] = ex
        st.frame = 0
        ctx.render(st, warmup, opts)
        if ex:
            ex.run()
        barrier()
        ctx.reset_stats()
        sampler = ClockSampler(device)
        sampler.prepare()
        sampler.start()
        barrier()
        t_wall = time.time()
        ctx.mark(0)
        st.frame = warmup
        ctx.render(st, steps, opts, sync=False)
        ms_exchange = 0.0
        if ex:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.mark(1)                      # joins the render streams into the image stream, so e0 starts after the last kernel
            e0.record(lib_stream)
            ex.run()
            e1.record(lib_stream)
        ctx.mark(1)                          # timed region = render + exchange, CUDA events on the library's stream
        sampler.sample_now()                 # the GPU is executing the timed work now
        my_ms = ctx.elapsed_ms()
        if ex:
            torch.cuda.synchronize()
            ms_exchange = e0.elapsed_time(e1)
        barrier()
        wall_ms = (time.time() - t_wall) * 1e3
        sampler.stop_flag.set()
        sampler.join(timeout=3)
        s = ctx.stats()
        rays = int(s.closest_rays + s.shadow_rays)
        if use_dist:
            t = torch.tensor([my_ms, ms_exchange], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r = torch.tensor([rays], dtype=torch.int64, device=dev)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
            total_ms, ms_exchange, total_rays = float(t[0].item()), float(t[1].item()), int(r.item())
        else:
            total_ms, total_rays = my_ms, rays
        return dict(st=st, opts=opts, fif=fif, accum=accum, ex=ex, total_ms=total_ms, total_rays=total_rays, exchange_ms=ms_exchange,
                    exchange_bytes=(ex.bytes if ex else 0), exchange_mode=(ex.mode if ex else "none"), exchange_why=(ex.why if ex else ""), wall_ms=wall_ms, launches=int(s.kernel_launches), clocks=sampler.summary(),
                    stats=s, n_local=n_local)

    m = measure(W, H, args.steps, args.warmup, args.frames_in_flight)
    st, opts, accum = m["st"], m["opts"], m["accum"]

    # ---- N > 1: the exchanged image on rank 0 must be, bit for bit, what one GPU renders unsharded (same frames)
    verified = None
    if use_dist and not args.no_verify_exchange and not args.scene_file:
        ok = 1
        if rank == 0:
            got = ctx.output_read()
            c1 = Context(device)
            upload_scene(c1, scene, W, H, build_mode=args.build_mode)
            c1.build_env_accel(scene.env)
            c1.set_camera(c1.make_camera(scene.camera["eye"], scene.camera["center"], scene.camera["up"], scene.camera["fov"], W / H,
                                         scene.camera.get("aperture", 0.0), 0 if scene.lights is None else len(scene.lights)))
            s1 = S.default_rtx_state(W, H, maxDepth=cfg["depth"], pbrMode=args.pbr_mode, fireflyClampThreshold=integral * 4.0)
            o1 = S.RenderOpts()
            o1.sort_rays = opts.sort_rays
            s1.frame = 0
            c1.render(s1, args.warmup + args.steps, o1)
            want = c1.output_read()
            c1.close()
            ok = 1 if np.array_equal(got, want) else 0
            if not ok:
                sys.stderr.write("exchange verification FAILED: %d of %d pixels differ\n" % (int((got != want).any(axis=-1).sum()), W * H))
        t = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)    # also the barrier that keeps the other ranks off rank 0's image meanwhile
        verified = bool(int(t.item()))
    s = m["stats"]
    total_ms, total_rays = m["total_ms"], m["total_rays"]

    def kernel_profile(c, state, steps, first_frame, base_opts):
        """Per-kernel device time (the same frames again with per-kernel CUDA events: kernels serialised, untimed pass) and the
        visit counters of an instrumented pass."""
        op = make_opts(frames_in_flight=base_opts.frames_in_flight, profile=1)
        c.reset_stats()
        state.frame = first_frame
        c.render(state, steps, op)
        sp = c.stats()
        oc = make_opts(frames_in_flight=base_opts.frames_in_flight, count_traversal=1)
        c.reset_stats()
        state.frame = first_frame
        c.render(state, steps, oc)
        sc = c.stats()
        return sp, sc

    sp, sc = kernel_profile(ctx, st, args.steps, args.warmup, opts)
    kernel_ms = dict(raygen=sp.ms_raygen, trace_closest=sp.ms_trace_closest, trace_shadow=sp.ms_trace_shadow, shade=sp.ms_shade,
                     accumulate=sp.ms_accumulate, sort=sp.ms_sort)
    # bounce classes: primary rays are coherent, everything after the first hit is not
    prim_c = m["n_local"] * args.steps
    sec_c = int(sp.closest_rays) - prim_c
    prim_s, sec_s = int(sp.shadow_rays_primary), int(sp.shadow_rays) - int(sp.shadow_rays_primary)

    def rate(n, ms):
        return (n / (ms * 1e-3) / 1e6) if ms > 0 else None

    bounce_classes = {
        "closest_primary": {"rays": prim_c, "ms": sp.ms_trace_closest_primary, "mrays_s": rate(prim_c, sp.ms_trace_closest_primary)},
        "closest_secondary": {"rays": sec_c, "ms": sp.ms_trace_closest - sp.ms_trace_closest_primary,
                              "mrays_s": rate(sec_c, sp.ms_trace_closest - sp.ms_trace_closest_primary)},
        "shadow_from_primary_hits": {"rays": prim_s, "ms": sp.ms_trace_shadow_primary, "mrays_s": rate(prim_s, sp.ms_trace_shadow_primary)},
        "shadow_from_secondary_hits": {"rays": sec_s, "ms": sp.ms_trace_shadow - sp.ms_trace_shadow_primary,
                                       "mrays_s": rate(sec_s, sp.ms_trace_shadow - sp.ms_trace_shadow_primary)},
        "what": "traversal kernels only, serialised profile pass of this rank (per-kernel CUDA events)"}

    peak, peak_src = measured_peaks()

    def trace_record(stats_p, stats_c):
        c_nodes, c_tris = stats_c.nodes_visited - stats_c.shadow_nodes_visited, stats_c.tris_tested - stats_c.shadow_tris_tested
        nb = int(stats_c.node_bytes) or 64
        bytes_c = algorithmic_bytes(stats_c.closest_rays, c_nodes, c_tris, nb)
        bytes_s = algorithmic_bytes(stats_c.shadow_rays, stats_c.shadow_nodes_visited, stats_c.shadow_tris_tested, nb)
        return {"achieved": bytes_c / max(stats_p.ms_trace_closest, 1e-9) / 1e6, "bytes_per_launch_set": bytes_c,
                "per_ray": {"node_bytes": nb, "nodes": c_nodes / max(1, stats_c.closest_rays), "tris": c_tris / max(1, stats_c.closest_rays),
                            "bytes": bytes_c / max(1, stats_c.closest_rays)},
                "shadow_kernel": {"achieved": bytes_s / max(stats_p.ms_trace_shadow, 1e-9) / 1e6,
                                  "nodes_per_ray": stats_c.shadow_nodes_visited / max(1, stats_c.shadow_rays),
                                  "tris_per_ray": stats_c.shadow_tris_tested / max(1, stats_c.shadow_rays)},
                "ms_trace_closest": stats_p.ms_trace_closest, "closest_rays": int(stats_c.closest_rays), "bvh_width": int(stats_c.bvh_width)}

    def profile_traffic(key):
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        try:
            return json.load(open(prof)).get(key)
        except Exception:
            return None

    head = trace_record(sp, sc)
    bvh_bytes = int(s.num_bvh_nodes) * int(sc.node_bytes or 64) + int(s.num_triangles) * 48
    l2_resident = bvh_bytes < 100e6
    headline_kernel = {"kernel": "traceKernelWide<CLOSEST>" if head["bvh_width"] == 4 else "traceKernel<CLOSEST>", "workload": args.workload,
                       "bound": "latency+divergence (L2-resident BVH)" if l2_resident else "hbm",
                       "bvh_bytes": bvh_bytes, "algorithmic_gbs": head["achieved"], "per_ray": head["per_ray"], "shadow_kernel": head["shadow_kernel"],
                       "dram_bytes_per_launch_ncu": profile_traffic(f"{args.workload}_closest_dram_bytes_per_launch"),
                       "dram_gbs_ncu": profile_traffic(f"{args.workload}_closest_dram_gbs"),
                       "lanes_per_instruction_ncu": profile_traffic(f"{args.workload}_closest_lanes_per_instruction"),
                       "note": "the BVH of this workload lives in the 126 MB L2: node bytes never cross HBM, so an HBM fraction is not defined for it "
                               "(round 1 printed 1.69).  ncu (profiles/): DRAM traffic is the ray and hit streams only; the kernel is bound by the "
                               "latency of dependent node fetches and by SIMT divergence of incoherent rays.  `algorithmic_gbs` is SURVEY 8d's "
                               "bytes / summed CUDA-event time, reported for continuity, not as a fraction of any peak."}

    # ---- e2e: the same metric through the C ABI as an interactive host drives it, one definition for every N: per step a host
    #      SceneCamera goes in, ONE frame is rendered (sharded when N > 1, then exchanged), and the tonemapped RGBA8 frame comes
    #      back to page-locked host memory on rank 0.  The read-back of frame f overlaps the rendering of frame f+1
    #      (vkrt_tonemap_async, a ring of staging buffers), as a display loop does.
    e2e = None
    if not args.no_e2e:
        cam = scene.camera
        tm = S.default_tonemapper()
        nbuf = 4    # staging ring: the host may run this many frames ahead of the copies (the library has four staging slots)
        bufs = [ctx.host_alloc((H, W, 4)) for _ in range(nbuf)] if rank == 0 else None
        opts_e = make_opts()
        k = max(4, min(args.steps, 32))
        ex = m["ex"]
        nl = 0 if scene.lights is None else len(scene.lights)

        def step(f):
            camera = ctx.make_camera(cam["eye"], cam["center"], cam["up"], cam["fov"], W / H, 0.0, nl)
            ctx.set_camera(camera)          # H2D: SceneCamera (140 B) + RtxState (48 B) travel as kernel parameters
            st.frame = f
            ctx.render(st, 1, opts_e, sync=False)
            if ex:
                ex.run()
            if rank == 0:
                ctx.tonemap_wait(f % nbuf)     # the copy that last used this staging slot (frame f - nbuf)
                ctx.tonemap_async(tm, bufs[f % nbuf], f % nbuf)   # D2H: the displayed RGBA8 frame
            if ex:
                ex.release()                   # the other ranks may store the next frame into rank 0's image once its tonemap has read this one

        def drain():
            if rank == 0:
                for b in range(nbuf):
                    ctx.tonemap_wait(b)

        for f in range(nbuf):
            step(f)
        drain()
        ctx.reset_stats()
        barrier()
        t0 = time.time()
        for f in range(k):
            step(nbuf + f)
        drain()
        ctx.sync()
        barrier()
        dt = time.time() - t0
        se = ctx.stats()
        er = int(se.closest_rays + se.shadow_rays)
        if use_dist:
            r = torch.tensor([er], dtype=torch.int64, device=dev)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
            er = int(r.item())
        e2e = {"value": er / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 140 + 48, "d2h_bytes_per_step": W * H * 4, "steps": k,
               "ms_per_step": dt / k * 1e3, "gpu_launches_per_step": int(se.kernel_launches) / k,
               "what": "per step: vkrt_set_camera(host SceneCamera) + vkrt_render(1 frame%s) + vkrt_tonemap_async -> page-locked host RGBA8 on "
                       "rank 0; frame f's read-back overlaps the frames after it (ring of four staging buffers, rotating wavefront slots); wall clock over %d steps "
                       "including the final waits" % (", tile shard + gather to rank 0" if use_dist else "", k)}
        if rank == 0:
            for b in bufs:
                ctx.host_free(b)

    # ---- HBM-resident roofline record (the one `frac` is defined for): the 10.4 M-triangle scene, 1.2 GB of nodes + triangles,
    #      measured in this same run for the 4-wide tree (the default above 1.5 M triangles) and the binary tree
    roofline = None
    if not use_dist and not args.no_hbm_record:
        def hbm_record(width_flag, label):
            c2 = ctx if args.workload == "instanced_10m" and width_flag == 0 else Context(device)
            cfg2 = WORKLOADS["instanced_10m"]
            if c2 is not ctx:
                sc2 = make_scene(cfg2)
                upload_scene(c2, sc2, cfg2["w"], cfg2["h"], build_mode=default_build_mode(sc2.num_triangles()) | width_flag)
            st2 = S.default_rtx_state(cfg2["w"], cfg2["h"], maxDepth=cfg2["depth"], fireflyClampThreshold=4.0)
            k = 8
            o2 = make_opts(frames_in_flight=k)
            c2.render(st2, k, o2)                         # warm-up
            p2, n2 = kernel_profile(c2, st2, k, k, o2)
            rec = trace_record(p2, n2)
            tri = int(c2.stats().num_triangles)
            nodes = int(c2.stats().num_bvh_nodes)
            if c2 is not ctx:
                c2.close()
            rec.update({"tree": label, "triangles": tri, "bvh_bytes": nodes * 64 + tri * 48, "frames": k})
            return rec
        r4 = hbm_record(0x200, "bvh4")
        r2 = hbm_record(0x100, "bvh2")
        traffic = profile_traffic("instanced_10m_closest_dram_bytes_per_launch")
        roofline = {"kernel": "traceKernelWide<CLOSEST>", "workload": "instanced_10m (HBM resident: %.2f GB of nodes + triangles)" % (r4["bvh_bytes"] / 1e9),
                    "bound": "hbm", "achieved": r4["achieved"], "peak": peak, "unit": "GB/s", "frac": r4["achieved"] / peak,
                    "traffic": traffic, "peak_source": peak_src, "per_ray": r4["per_ray"], "ms_trace_closest": r4["ms_trace_closest"],
                    "closest_rays": r4["closest_rays"], "frames": r4["frames"],
                    "bvh2": {"achieved": r2["achieved"], "frac": r2["achieved"] / peak, "per_ray": r2["per_ray"], "ms_trace_closest": r2["ms_trace_closest"]},
                    "headline_kernel": headline_kernel,
                    "note": "achieved = algorithmic bytes (ray 32 + hit 16 + 64 per node visited + 48 per triangle tested, visit counters of an "
                            "instrumented pass over the same frames) / summed CUDA-event time of the closest-hit launches of those frames; both "
                            "trees fetch 64-byte nodes.  `traffic` = ncu dram bytes of one closest-hit launch (profiles/).  The headline workload's "
                            "BVH is L2 resident: see headline_kernel."}
    else:
        roofline = {"kernel": headline_kernel["kernel"], "workload": args.workload, "bound": headline_kernel["bound"], "achieved": head["achieved"],
                    "peak": peak, "unit": "GB/s", "frac": (head["achieved"] / peak) if not l2_resident else None, "traffic": None,
                    "peak_source": peak_src, "per_ray": head["per_ray"], "headline_kernel": headline_kernel,
                    "note": "multi-GPU / --no-hbm-record run: the HBM-resident record is measured by the N=1 default run"}

    # ---- C4 of BASELINE.json: the same scene at 3840x2160, tile-sharded over the ranks, in the same run (N > 1)
    c4 = None
    if use_dist and cfg["scene"] == "sponza" and not args.no_c4:
        m4 = measure(3840, 2160, min(args.steps, 32), 3, 0)
        c4 = {"workload": "sponza_4k (BASELINE.json configs[3]: 3840x2160, tile-sharded, exchange of the RGBA32F accumulation buffer)",
              "value": m4["total_rays"] / (m4["total_ms"] * 1e-3) / 1e6, "unit": "Mrays/s", "steps": min(args.steps, 32),
              "ms_per_step": m4["total_ms"] / min(args.steps, 32), "exchange_ms": m4["exchange_ms"], "exchange_bytes_per_rank": m4["exchange_bytes"],
              "note": "1024 spp of the config bounded to these steps; Mrays/s does not depend on the sample count"}

    cb = None
    if rank == 0 and not args.no_cpu_baseline and not use_dist:
        cb, _ = cpu_reference(cfg, 8, 1, budget_s=25.0)   # ~10-25 s of host work, bounded by shrinking the band of rows

    if rank == 0:
        line = {"metric": "Mrays/s", "value": total_rays / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scene": scene.name, "triangles": int(s.num_triangles), "bvh_nodes": int(s.num_bvh_nodes),
                           "bvh_build": {"mode": {0: "lbvh", 1: "lbvh+leaf-collapse", 2: "host-sah", 3: "ploc+leaf-collapse", 4: "gpu-binned-sah"}[args.build_mode],
                                         "ms": float(s.ms_build), "ms_first_build_in_process": cold_build_ms, "bvh_width": int(s.bvh_width), "max_depth": int(s.bvh_max_depth),
                                         "what": "vkrt_scene_commit: instance flattening + BVH build, outside the timed region (SURVEY 8d)"},
                           "resolution": [W, H], "max_depth": cfg["depth"], "spp_per_step": 1, "pbr_mode": args.pbr_mode,
                           "frames_in_flight": m["fif"],
                           "sharding": (f"{world} x interleaved {TILE}x{TILE} tiles; exchange: " + {"peer": "accumulate kernels store owned pixels into rank 0's image over NVLink (CUDA IPC mapping) + a 4-byte all-reduce as fence", "gather": "pack + NCCL gather of owned pixels to rank 0 + unpack"}[m["exchange_mode"]] + (f" (peer mapping refused: {m['exchange_why']})" if m["exchange_why"] else "")) if use_dist else "none",
                           "exchange_verified": verified,
                           "l2": "per-step wavefront state (~170 B x %d paths) exceeds the 126 MB L2; the BVH/geometry is L2-resident by design" % (W * H),
                           "rays_counted": "ClosestHit + AnyHit calls (device counters)",
                           "vulkan": vulkan_probe()},
                "rays": total_rays, "gpu_launches": m["launches"], "kernel_ms": kernel_ms, "bounce_classes": bounce_classes,
                "reduce_ms": m["exchange_ms"], "exchange_bytes_per_rank": m["exchange_bytes"], "wall_ms": m["wall_ms"],
                "clocks": m["clocks"], "roofline": roofline, "e2e": e2e}
        if c4 is not None:
            line["c4"] = c4
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if live["ex"] is not None:
        live["ex"].close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
