"""Runs a few tracker invocations at one resolution (for ncu launch lists / captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as gu, scenes
from test_tracker_gpu import _cuda_odometry
W = int(sys.argv[1]) if len(sys.argv) > 1 else 640
H = W * 3 // 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = scenes.room_pair(W, H)
co = _cuda_odometry(gu, case)
torch.cuda.synchronize()
for _ in range(n):
    co.track(case["T0"])
torch.cuda.synchronize()
