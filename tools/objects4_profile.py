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
    if len(sys.argv) > 2 and sys.argv[2] == "trace":  # phase trace of the tracker launch once two models are live
        import ctypes as C
        dbg = torch.zeros(2048, dtype=torch.int64, device="cuda")
        cfb.check(cfb.lib().cfb_cofusion_set_debug_trace(cf._h, C.c_void_p(dbg.data_ptr())))
        for t in range(n):
            cf.process_frame(*dev[bench.frame_index(t, n_render)])
            torch.cuda.synchronize()
            if cf.num_models >= 2 and t % 10 == 9:
                q = dbg.cpu().numpy().astype(np.int64)
                print("frame %d, %d models: barrier0 %.1f us, so3 %.1f us, gn %.1f us" % (
                    t, cf.num_models, (q[1] - q[0]) / 1e3, (q[2] - q[1]) / 1e3, (q[3] - q[2]) / 1e3))
                G = 144
                st = q[256:256 + 8 * G].reshape(G, 8)[:, :5]
                if st.min() > 0:
                    d = (st - st[:, :1].min()) / 1e3
                    order = np.argsort(d[:, 3])
                    print("  iteration 12, per CTA (us since the first CTA started it): start, after residual, after icp, after rgb rows, after collect")
                    for b in list(order[:3]) + list(order[-8:]):
                        print("    cta %3d (tile %2d,%2d): %s" % (b, b % 8, b // 8, " ".join("%6.2f" % v for v in d[b])))
                for it in (0, 5, 9, 10, 14, 15, 18):
                    b = 8 + it * 8
                    nxt = q[b + 8] if it < 18 else q[3]
                    print("  it %2d: residual+arriveA %.2f  icp %.2f  waitA %.2f  rgbrows %.2f  publish %.2f  collect %.2f  solve %.2f  | total %.2f" % (
                        it, (q[b+1]-q[b])/1e3, (q[b+2]-q[b+1])/1e3, (q[b+3]-q[b+2])/1e3, (q[b+4]-q[b+3])/1e3, (q[b+5]-q[b+4])/1e3,
                        (q[b+6]-q[b+5])/1e3, (q[b+7]-q[b+6])/1e3, (nxt-q[b])/1e3))
        return
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
