"""-m "not gpu": the N>1 host logic of bench.py (frame packing, model sharding, one broadcast per step)
on world_size 2 with the gloo backend -- no GPU, no compute kernels."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from cofusion_b200 import sharding
    rng = np.random.default_rng(7)
    rgb = rng.integers(1, 255, (H, W, 3), dtype=np.uint8)
    depth = rng.uniform(0.5, 4.0, (H, W)).astype(np.float32)
    mask = rng.integers(0, 3, (H, W), dtype=np.uint8)
    recv = torch.zeros(sharding.packed_bytes(W, H), dtype=torch.uint8)
    for step in range(3):
        if rank == 0:  # the root packs [rgb u8 3P | depth f32 4P | mask u8 P]: the layout the library broadcasts
            recv.copy_(torch.from_numpy(sharding.pack_frame(rgb + step, depth + step, mask)))
        dist.broadcast(recv, src=0)  # the single collective of the data path (ncclBroadcast in csrc/shard.cu)
        r, d, m = sharding.unpack_frame(recv.numpy(), W, H)
        assert np.array_equal(r, rgb + step) and np.array_equal(d, depth + step) and np.array_equal(m, mask)
        assert bench.frame_index(step, 4) in range(4)
    # rank r owns model r: the per-rank model lists are disjoint and cover the scene
    n_models = 5
    mine = sharding.models_of_rank(n_models, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(n_models))
    assert sharding.scene_models(1) == 1 and sharding.scene_models(8) == 8 and sharding.owner(0, world) == 0
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction used by bench.py
    assert t.item() == world
    assert abs(sharding.aggregate_value(world, 100, 50.0) - world * 2000.0) < 1e-6
    dist.destroy_process_group()


def test_two_ranks_broadcast_and_shard():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 64, 48, None), nprocs=2, join=True)


def test_algorithmic_bytes_formula():
    import bench
    # GN loop per model per frame 388.6 MB @640x480, 1.554 GB @1280x960
    assert abs(bench.gn_algorithmic_bytes(640, 480) / 1e6 - 388.6) < 0.1
    assert abs(bench.gn_algorithmic_bytes(1280, 960) / 1e9 - 1.554) < 0.001
