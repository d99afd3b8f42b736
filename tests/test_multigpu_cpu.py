"""world_size-2 gloo test (CPU) of the N > 1 path: tiles dealt round-robin to the ranks, each rank renders its shard
into a zero image, one reduce(sum) to rank 0 -- which must equal the single-process image bit for bit, because the
shards are disjoint and the RNG is keyed on the global pixel index.  The renderer here is the CPU oracle
(this test has no GPU); tests/test_gpu_parity.py checks the same property for the CUDA path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    from vk_raytrace_b200 import scenes, structs as S, upload_scene
    W, H = 96, 64
    o = ob.Oracle(threads=2)
    upload_scene(o, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=3)
    o.render_tiles(st, 3, tile=16, shard_rank=rank, shard_count=world)
    img = torch.from_numpy(o.output_read())
    rays = torch.tensor([sum(o.counters()[k] for k in ("closest_rays", "shadow_rays"))], dtype=torch.int64)
    dist.reduce(img, 0)
    dist.all_reduce(rays)
    if rank == 0:
        np.savez(out_path, img=img.numpy(), rays=rays.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_shards_reduce_to_the_whole_image(tmp_path, ob):
    from vk_raytrace_b200 import scenes, structs as S, upload_scene
    out = str(tmp_path / "reduced.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    W, H = 96, 64
    o = ob.Oracle()
    upload_scene(o, scenes.cornell_box(with_light=True), W, H)
    st = S.default_rtx_state(W, H, maxDepth=3)
    o.render(st, 3)
    assert np.array_equal(got["img"], o.output_read())
    c = o.counters()
    assert int(got["rays"][0]) == c["closest_rays"] + c["shadow_rays"]
