"""Small end-to-end run for compute-sanitizer: 5 frames 320x240, background + 2 objects (batched tracker),
CRF segmentation on the last frame.  usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import cofusion_b200 as cfb
from cofusion_b200 import synth
import scenes

W, H = 320, 240
K = scenes.scaled_K(W)
seq = list(synth.room_sequence(5, W, H, K, noise=True, n_boxes=2, box_speed=0.5))
cf = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 17))
for t, (_, rgb, d, _, ids) in enumerate(seq):
    cf.process_frame(np.ascontiguousarray(rgb), np.ascontiguousarray(d), np.ascontiguousarray(ids.astype(np.uint8)))
    if t == 1:
        cf.spawn_object_model(1)
        cf.spawn_object_model(2)
ms = [cf.model(i) for i in range(cf.num_models)]
seg = cfb.Segmentation(W, H)
_, rgb, d, _, _ = seq[-1]
full, mds, hn = seg.perform_crf(torch.from_numpy(np.ascontiguousarray(rgb)).cuda(), torch.from_numpy(np.ascontiguousarray(d)).cuda(),
                                [m.info()[0] for m in ms], [m.view_ptr(3) for m in ms], [m.view_ptr(9) for m in ms], 3, True)
torch.cuda.synchronize()
print("sanitize_smoke done: models", len(ms), "labels", len(mds), "surfels", [m.last_count() for m in ms])
