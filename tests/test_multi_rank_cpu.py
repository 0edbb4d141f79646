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
    P = W * H
    rng = np.random.default_rng(7)
    rgb = rng.integers(1, 255, (H, W, 3), dtype=np.uint8)
    depth = rng.uniform(0.5, 4.0, (H, W)).astype(np.float32)
    recv = torch.zeros(7 * P, dtype=torch.uint8)
    sums = []
    for step in range(3):
        if rank == 0:  # the root packs [rgb u8 3P | depth f32 4P]
            recv[:3 * P] = torch.from_numpy(rgb.reshape(-1)) + step
            recv[3 * P:] = torch.from_numpy((depth + step).reshape(-1).view(np.uint8))
        dist.broadcast(recv, src=0)  # the single collective of the data path
        r = recv[:3 * P].numpy().reshape(H, W, 3)
        d = recv[3 * P:].view(torch.float32).numpy().reshape(H, W)
        assert np.array_equal(r, rgb + step) and np.array_equal(d, depth + step)
        sums.append(float(d.sum()))
        assert bench.frame_index(step, 4) in range(4)
    # rank r owns model r: ids are disjoint and cover the model list
    models = list(range(5))
    mine = [m for m in models if m % world == rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        assert sorted(sum(gathered, [])) == models
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction used by bench.py
    assert t.item() == world
    dist.destroy_process_group()


def test_two_ranks_broadcast_and_shard():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 64, 48, None), nprocs=2, join=True)


def test_algorithmic_bytes_formula():
    import bench
    # GN loop per model per frame 388.6 MB @640x480, 1.554 GB @1280x960
    assert abs(bench.gn_algorithmic_bytes(640, 480) / 1e6 - 388.6) < 0.1
    assert abs(bench.gn_algorithmic_bytes(1280, 960) / 1e9 - 1.554) < 0.001
