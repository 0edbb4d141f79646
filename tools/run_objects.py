"""Run CoFusion with the motion segmentation on a synthetic multi-object sequence; print model counts and
per-frame time.  usage: run_objects.py [n_frames] [n_boxes] [box_speed] [wErr] [thNew]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import cofusion_b200 as cfb
from cofusion_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
speed = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
W, H, K = 640, 480, synth.K_DEFAULT
seq = list(synth.room_sequence(n, W, H, K, noise=True, n_boxes=nb, box_speed=speed, box_start=10))
p = cfb.CoFusionParams.default(1 << 21)
p.enableMultipleModels = 1
p.seg.unaryWeightError = float(sys.argv[4]) if len(sys.argv) > 4 else 150.0
p.seg.unaryThresholdNew = float(sys.argv[5]) if len(sys.argv) > 5 else 3.5
p.confGlobalInit = float(sys.argv[6]) if len(sys.argv) > 6 else 10.0
cf = cfb.CoFusion(W, H, K, p)
t0 = time.time()
for t, (ts, rgb, d, T, ids) in enumerate(seq):
    cf.process_frame(np.ascontiguousarray(rgb), np.ascontiguousarray(d))
    if t and t % 10 == 0:
        cf.ctx.sync()
        dt = (time.time() - t0) / 10
        t0 = time.time()
        mds, hn, sp, de = cf.last_segmentation()
        print(t, "models", [cf.model(i).info()[0] for i in range(cf.num_models)], "ms/frame %.2f" % (dt * 1e3),
              "spx", [m.superPixelCount for m in mds], "conf", ["%.1f" % cf.model(i).info()[1] for i in range(cf.num_models)],
              "inactive", cf.num_inactive_models, flush=True)
