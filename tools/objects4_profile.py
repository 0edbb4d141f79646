"""Per-frame wall time of the configs[2] closed loop (4 moving boxes, CRF in the loop): which frames are slow and why.
Tool only (not a bench value): every frame is followed by a device synchronise.
usage: python tools/objects4_profile.py [frames]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import cofusion_b200 as cfb
    from cofusion_b200 import synth
    import bench
    W, H = bench.W, bench.H
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    pipelined = len(sys.argv) > 2 and sys.argv[2] == "pipelined"  # no drain between frames (as the bench runs)
    n_render = 32
    seq = list(synth.room_sequence(n_render, W, H, synth.K_DEFAULT, noise=True, n_boxes=4, box_speed=1.0, seed=1234))
    dev = [(torch.from_numpy(np.ascontiguousarray(r)).cuda(), torch.from_numpy(np.ascontiguousarray(d)).cuda()) for _, r, d, _, _ in seq]
    p = cfb.CoFusionParams.default(1 << 21)
    p.enableMultipleModels = 1
    cf = cfb.CoFusion(W, H, synth.K_DEFAULT, p, device=0)
    rows = []
    if pipelined:
        t0 = time.perf_counter()
        for t in range(n):
            cf.process_frame(*dev[bench.frame_index(t, n_render)])
            if t % 20 == 19:
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                print("frames %3d-%3d: %.3f ms/frame, %d models, %.1f launches/frame" % (
                    t - 19, t, (t1 - t0) * 1e3 / 20, cf.num_models, cf.ctx.take_launch_count() / 20))
                t0 = time.perf_counter()
        return
    for t in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cf.process_frame(*dev[bench.frame_index(t, n_render)])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rows.append((t, cf.num_models, (t1 - t0) * 1e3, (t2 - t0) * 1e3, cf.ctx.take_launch_count()))
    prev = 1
    for t, nm, host, tot, l in rows:
        flag = " <- model list changed" if nm != prev else ""
        print("frame %3d models %d  call %.3f ms  call+drain %.3f ms  launches %d%s" % (t, nm, host, tot, l, flag))
        prev = nm
    a = np.array([(r[1], r[3]) for r in rows[20:]])
    for k in sorted(set(a[:, 0])):
        sel = a[a[:, 0] == k][:, 1]
        print("models=%d: %d frames, median %.3f ms, mean %.3f ms, max %.3f ms" % (k, len(sel), np.median(sel), sel.mean(), sel.max()))


if __name__ == "__main__":
    main()
