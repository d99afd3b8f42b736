"""world_size-2 gloo test (CPU) of the N > 1 path: tiles dealt round-robin to the ranks, each rank renders its shard
into a zero image, one reduce(sum) to rank 0 -- which must equal the single-process image bit for bit, because the
shards are disjoint and the RNG is keyed on the global pixel index -- and the two exchanges bench.py uses: the
gather of owned pixels, and the peer-image protocol (stores into the root's mapped image + two fences per frame).  The renderer here is the CPU oracle
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
    # exchange 2 (round 2, what bench.py does on NCCL): every rank sends only the pixels it owns, in shard order, padded to the
    # largest shard; the root scatters them.  1 / world of the bytes of the zero-padded reduce below, same image.
    from vk_raytrace_b200.context import shard_pixel_list
    lists = [shard_pixel_list(W, H, 16, r, world) for r in range(world)]
    cap = max(len(l) for l in lists)
    mine = torch.zeros((cap, 4), dtype=torch.float32)
    mine[: len(lists[rank])] = img.reshape(-1, 4)[torch.from_numpy(lists[rank].astype(np.int64))]
    gathered = [torch.zeros((cap, 4), dtype=torch.float32) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    rebuilt = None
    if rank == 0:
        rebuilt = img.clone().reshape(-1, 4)
        for r in range(1, world):
            rebuilt[torch.from_numpy(lists[r].astype(np.int64))] = gathered[r][: len(lists[r])]
        rebuilt = rebuilt.reshape(H, W, 4)
    # exchange 3 (the default of bench.py on a GPU box): the root publishes its image (there: a CUDA IPC handle, here: a named
    # shared-memory block), the other ranks map it and store the pixels they own straight into it after every frame (there: the
    # accumulate kernel's mirror stores); a 1-element all-reduce is the fence before the root reads a frame, a second one the
    # fence before anybody stores the next.  The root must see whole frames: the running mean after 1, 2, 3 frames.
    from multiprocessing import shared_memory
    shm = shared_memory.SharedMemory(create=True, size=W * H * 16) if rank == 0 else None
    names = [shm.name if rank == 0 else None]
    dist.broadcast_object_list(names, src=0)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=names[0])
    root_img = np.ndarray((H * W, 4), dtype=np.float32, buffer=shm.buf)
    o2 = ob.Oracle(threads=2)
    upload_scene(o2, scenes.cornell_box(with_light=True), W, H)
    token, seen = torch.zeros(1), []
    own = lists[rank].astype(np.int64)
    for f in range(3):
        o2.render_tiles(S.default_rtx_state(W, H, maxDepth=3, frame=f), 1, tile=16, shard_rank=rank, shard_count=world)
        root_img[own] = o2.output_read().reshape(-1, 4)[own]     # the mirror stores of this rank's accumulate
        dist.all_reduce(token)                                   # fence: every rank's pixels of frame f are in place
        if rank == 0:
            seen.append(root_img.reshape(H, W, 4).copy())        # the root's use of the frame (there: tonemap + read-back)
        dist.all_reduce(token)                                   # release: the root has read frame f
    dist.reduce(img, 0)
    dist.all_reduce(rays)
    if rank == 0:
        np.savez(out_path, img=img.numpy(), rays=rays.numpy(), rebuilt=rebuilt.numpy(), shard_sizes=np.array([len(l) for l in lists]),
                 peer_frames=np.stack(seen))
    del root_img
    shm.close()
    dist.barrier()
    if rank == 0:
        shm.unlink()
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
    assert np.array_equal(got["rebuilt"], o.output_read())          # gather of owned pixels == reduce of zero-padded images
    assert int(got["shard_sizes"].sum()) == W * H
    c = o.counters()
    assert int(got["rays"][0]) == c["closest_rays"] + c["shadow_rays"]
    # the peer exchange: the frames the root saw are the single-process running means after 1, 2, 3 frames
    o3 = ob.Oracle()
    upload_scene(o3, scenes.cornell_box(with_light=True), W, H)
    for f in range(3):
        o3.render(S.default_rtx_state(W, H, maxDepth=3, frame=f), 1)
        assert np.array_equal(got["peer_frames"][f], o3.output_read()), f

