"""BASELINE.json configs[2]: 4 moving objects + background on ONE GPU, motion-CRF segmentation on.

Five models (background + 4 boxes) are tracked, fused, cleaned and predicted every frame, and the CRF
segmentation runs over all five models' ICP-error / confidence maps (6 labels incl. "new").  To keep
the model set stable over an arbitrarily long timed region, the objects are spawned once from the
renderer's label image (FrameData::mask path, Segmentation.cpp:61-119) and the fuse/clean stage keeps
using those labels; the segmentation result of every frame is computed (and timed) but not fed back.
Prints one JSON line.  (The CPU figure for the same segmentation -- the part that is CPU code in the
reference -- is part of bench.py's cpu_baseline: `segmentation_ms_per_frame`, 1 and 5 models.)

  python tools/bench_objects.py [--steps K] [--warmup W]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import cofusion_b200 as cfb
from cofusion_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--warmup", type=int, default=30)
ap.add_argument("--no-seg", action="store_true")
args = ap.parse_args()
W, H, K = 640, 480, synth.K_DEFAULT
NB = 4
n_render = 32
seq = list(synth.room_sequence(n_render, W, H, K, noise=True, n_boxes=NB, box_speed=0.5))
period = 2 * (n_render - 1)
fidx = lambda t: (t % period) if (t % period) < n_render else period - (t % period)

p = cfb.CoFusionParams.default(1 << 21)
cf = cfb.CoFusion(W, H, K, p)
ext = torch.cuda.ExternalStream(cf.ctx.stream)
dev = [(torch.from_numpy(np.ascontiguousarray(r)).cuda(), torch.from_numpy(np.ascontiguousarray(d)).cuda(),
        torch.from_numpy(np.ascontiguousarray(i.astype(np.uint8))).cuda()) for _, r, d, _, i in seq]
seg = cfb.Segmentation(W, H)
scratch = None
with torch.cuda.stream(ext):
    cf.process_frame(*dev[0])
    cf.process_frame(*dev[1])
    for k in range(1, NB + 1):
        cf.spawn_object_model(k)
    models = [cf.model(i) for i in range(cf.num_models)]
    cfb.lib().cfb_model_odometry.restype = cfb.C.c_void_p
    odom0 = cfb.C.c_void_p(cfb.lib().cfb_model_odometry(models[0]._h))
    cfb.check(cfb.lib().cfb_odom_enable_kernel_timing(odom0, 1))
    ids = [m.info()[0] for m in models]
    icp = [m.view_ptr(3) for m in models]
    vconf = [m.view_ptr(9) for m in models]

    def step(t):
        rgb, d, mask = dev[fidx(t)]
        cf.process_frame(rgb, d, mask)
        if not args.no_seg:
            seg.perform_crf(rgb, d, ids, icp, vconf, NB + 1, True)

    for t in range(2, 2 + args.warmup):
        step(t)
    cf.ctx.sync()
    cf.ctx.take_launch_count()
    cfb.check(cfb.lib().cfb_odom_kernel_timing(odom0, None, None, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(2 + args.warmup, 2 + args.warmup + args.steps):
        step(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
launches = cf.ctx.take_launch_count() + (0 if args.no_seg else args.steps * 21)
ksum, kn = cfb.C.c_double(0), cfb.C.c_int(0)
cfb.check(cfb.lib().cfb_odom_kernel_timing(odom0, cfb.C.byref(ksum), cfb.C.byref(kn), 0))
counts = [m.last_count() for m in models]

print(json.dumps({
    "metric": "RGB-D frames/s @640x480, 4 tracked objects + background on 1 GPU, CRF segmentation on",
    "value": args.steps / (ms / 1e3), "unit": "frames/s", "ms_per_step": ms / args.steps, "steps": args.steps,
    "warmup": args.warmup, "models": len(models),
    "tracker_kernel_ms": (ksum.value / kn.value) if kn.value else None,
    "tracker": "one persistent launch for all 5 models (gn_batched.cu)", "surfels_per_model": counts, "gpu_launches": launches,
    "segmentation": "off" if args.no_seg else "6 labels, 10 mean-field iterations, every frame",
    "data": "synthetic room + 4 moving boxes, labels from the renderer (external mask path)"}))
