"""Object sharding check (run under torchrun, N >= 2 GPUs): rank r owns model r of one N-model scene, the frame reaches
the ranks by the library's NCCL broadcast.  Every rank's model must end up with EXACTLY the pose and surfel count it has
when all N models run in one process on one GPU (the tracker is bit-identical batched or alone, the surfel stage is
per model).  Prints one line per rank, exit code 0 = all equal.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cofusion_b200 as cfb  # noqa: E402
from cofusion_b200 import sharding, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, K, frames = 640, 480, synth.K_DEFAULT, 8
n_models = sharding.scene_models(world)
seq = list(synth.room_sequence(frames, W, H, K, noise=True, seed=1234, n_boxes=n_models - 1, box_speed=0.5)) if rank == 0 else None
uid = [cfb.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
cf = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 20), device=local)
cf.shard_init(rank, world, uid[0])
ref = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 20), device=local) if rank == 0 else None
for t in range(frames):
    if rank == 0:
        rgb, d, ids = np.ascontiguousarray(seq[t][1]), np.ascontiguousarray(seq[t][2]), np.ascontiguousarray(seq[t][4].astype(np.uint8))
        cf.process_frame(rgb, d, ids)
        ref.process_frame(rgb, d, ids)
    else:
        cf.process_frame(None, None, None)
    if t == 1:
        # a new model starts at the camera pose (CoFusion.cpp:593), which only the camera model's rank tracks
        cam = torch.from_numpy(cf.model(0).pose.reshape(16).copy()).cuda() if rank == 0 else torch.empty(16, device="cuda")
        dist.broadcast(cam, src=0)
        for m in sharding.models_of_rank(n_models, rank, world):
            if m > 0:
                cf.spawn_object_model(m, cam.cpu().numpy())
        if rank == 0:
            for m in range(1, n_models):
                ref.spawn_object_model(m)
mine = cf.model(0 if rank == 0 else cf.num_models - 1)
mask_sum = int(cf.ctx_view_mask().astype(np.int64).sum())
rec = (rank, mine.info()[0], mine.pose.copy(), mine.last_count(), mask_sum)
got = [None] * world
dist.all_gather_object(got, rec)
ok = True
if rank == 0:
    for r, mid, pose, cnt, msum in got:
        j = [i for i in range(ref.num_models) if ref.model(i).info()[0] == (0 if r == 0 else mid)][0]
        same = np.array_equal(pose, ref.model(j).pose) and cnt == ref.model(j).last_count() and msum == got[0][4]
        print("rank %d model id %d: pose %s, surfels %d vs %d, mask checksum %s" % (
            r, mid, "identical" if np.array_equal(pose, ref.model(j).pose) else "DIFFERENT", cnt, ref.model(j).last_count(),
            "identical on all ranks" if msum == got[0][4] else "DIFFERENT"), flush=True)
        ok = ok and same
    print("sharded == single process: %s" % ok, flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
torch.cuda.synchronize()
del cf, ref
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
