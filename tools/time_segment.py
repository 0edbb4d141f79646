"""Time cfb_segmentation_perform_crf on the synthetic room frame (device buffers, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import cofusion_b200 as cfb
import seg_cases

side = torch.cuda.Stream()
torch.cuda.set_stream(side)  # a capturable stream: the launch sequence is replayed as a CUDA graph
for name, c in (("1 model", seg_cases.room_case()), ("2 models", seg_cases.two_model_case())):
    H, W = c["depth"].shape
    seg = cfb.Segmentation(W, H)
    rgb = torch.from_numpy(c["rgb"]).cuda()
    depth = torch.from_numpy(c["depth"]).cuda()
    icp = [torch.from_numpy(a).cuda() for a in c["icp"]]
    vc = [torch.from_numpy(a).cuda() for a in c["vc"]]
    for _ in range(3):
        seg.perform_crf(rgb, depth, c["model_ids"], icp, vc, c["next_id"], True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        seg.perform_crf(rgb, depth, c["model_ids"], icp, vc, c["next_id"], True)
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.3f ms per performSegmentationCRF (640x480)" % (name, e0.elapsed_time(e1) / n))
