"""Phase trace (%globaltimer) of the persistent tracker kernel. Scratch tool."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpu_util as gu, scenes, cofusion_b200 as cfb
from test_tracker_gpu import _cuda_odometry
W = int(sys.argv[1]) if len(sys.argv) > 1 else 640
case = scenes.room_pair(W, W * 3 // 4)
co = _cuda_odometry(gu, case)
dbg = torch.zeros(2048, dtype=torch.int64, device="cuda")
cfb.check(cfb.lib().cfb_odom_set_debug_trace(co._h, C.c_void_p(dbg.data_ptr())))
for _ in range(3):
    co.track(case["T0"])
torch.cuda.synchronize()
t = dbg.cpu().numpy().astype(np.int64)
t0 = t[0]
print("barrier0 %.1f us, so3 %.1f us, gn %.1f us" % ((t[1]-t[0])/1e3, (t[2]-t[1])/1e3, (t[3]-t[2])/1e3))
for q in range(19):
    b = 8 + q * 8
    nxt = t[b + 8] if q < 18 else t[3]
    print("it %2d: residual+arriveA %.2f  icp %.2f  waitA %.2f  rgbrows %.2f  publish+wait %.2f  fold %.2f  solve %.2f  | total %.2f" % (
        q, (t[b+1]-t[b])/1e3, (t[b+2]-t[b+1])/1e3, (t[b+3]-t[b+2])/1e3, (t[b+4]-t[b+3])/1e3, (t[b+5]-t[b+4])/1e3,
        (t[b+6]-t[b+5])/1e3, (t[b+7]-t[b+6])/1e3, (nxt-t[b])/1e3))
