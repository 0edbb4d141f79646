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
if os.environ.get("SANITIZE_PART") == "1":  # racecheck / synccheck are slow: the small part only
    print("sanitize_smoke part 1 done: models", len(ms), "labels", len(mds), "surfels", [m.last_count() for m in ms])
    sys.exit(0)
# the closed loop of tests/test_configs_gpu.py::test_config2 (640x480, spawns from frame 4, a loss within 12 frames):
# the segmentation decides spawn / loss -> pooled models (recycle), per-model streams, object-model tracker phases
W2, H2, K2 = 640, 480, synth.K_DEFAULT
p = cfb.CoFusionParams.default(1 << 19)
p.confGlobalInit = 1.5
p.enableMultipleModels = 1
p.modelSpawnOffset = 2
p.seg.unaryWeightError = 150.0
p.seg.unaryThresholdNew = 3.5
cl = cfb.CoFusion(W2, H2, K2, p)
ids_seen, lost = set(), 0
for t, (_, rgb, d, _, _) in enumerate(synth.room_sequence(14, W2, H2, K2, noise=True, n_boxes=4, box_speed=4.0, box_start=3)):
    cl.process_frame(np.ascontiguousarray(rgb), np.ascontiguousarray(d))
    ids_seen.update(cl.model(i).info()[0] for i in range(cl.num_models))
    if t > 0:
        lost += cl.last_segmentation()[3]
torch.cuda.synchronize()
print("closed loop: models seen", sorted(ids_seen), "live", cl.num_models, "lost", lost, "inactive", cl.num_inactive_models)
print("sanitize_smoke done: models", len(ms), "labels", len(mds), "surfels", [m.last_count() for m in ms])
