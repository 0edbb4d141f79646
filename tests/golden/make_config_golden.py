"""Golden fixtures for BASELINE.json's stated configurations, generated on the CPU by the oracle
(TEST INFRASTRUCTURE): the long sequences would otherwise keep a GPU box waiting on the CPU restatement.

  configs[0]  64-frame static plane, 640x480          -> plane64_oracle.npz  (oracle pose per frame, ground truth)
  configs[2]  4 moving objects, CRF closed loop, 640x480 -> objects4_oracle.npz (per frame: model ids, poses,
              label mask (compressed), ModelData, spawn / loss events)

    python tests/golden/make_config_golden.py [plane] [objects]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import orc  # noqa: E402
from cofusion_b200 import synth  # noqa: E402
from orc_pipeline import OracleCoFusion, OraclePipeline  # noqa: E402

OBJ_FRAMES = 30
OBJ_SETUP = dict(n_boxes=4, box_speed=4.0, box_start=3, conf_global=1.5, spawn_offset=2, unaryWeightError=150.0,
                 unaryThresholdNew=3.5, max_surfels=1 << 19)


def plane():
    W, H, n = 640, 480, 64
    op = OraclePipeline(W, H, synth.K_DEFAULT, 1 << 20)
    poses, gts = [], []
    T0i = None
    for ts, rgb, d, T in synth.plane_sequence(n, W, H, synth.K_DEFAULT):
        if T0i is None:
            T0i = np.linalg.inv(T)
        op.process_frame(rgb, d)
        poses.append(op.pose.copy())
        gts.append((T0i @ T).astype(np.float32))
        print("plane frame", len(poses), float(np.abs(poses[-1] - gts[-1]).max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "plane64_oracle.npz"), poses=np.array(poses), gt=np.array(gts))


def objects():
    W, H = 640, 480
    K = synth.K_DEFAULT
    s = OBJ_SETUP
    prm = orc.OrcSegParams.default()
    prm.unaryWeightError = s["unaryWeightError"]
    prm.unaryThresholdNew = s["unaryThresholdNew"]
    op = OracleCoFusion(W, H, K, max_surfels=s["max_surfels"], conf_global=s["conf_global"], spawn_offset=s["spawn_offset"],
                        seg_params=prm)
    out = {}
    seq = synth.room_sequence(OBJ_FRAMES, W, H, K, noise=True, n_boxes=s["n_boxes"], box_speed=s["box_speed"],
                              box_start=s["box_start"])
    for t, (_, rgb, d, _, _) in enumerate(seq):
        op.process_frame(np.ascontiguousarray(rgb), np.ascontiguousarray(d))
        out["ids_%d" % t] = np.array([m.id for m in op.models], np.int32)
        out["poses_%d" % t] = np.array([m.pose for m in op.models], np.float32)
        out["counts_%d" % t] = np.array([m.map.count for m in op.models], np.int64)
        out["mask_%d" % t] = op.mask.copy()
        if t > 0 and op.last_seg is not None:
            mds, has_new, spawned, lost = op.last_seg
            out["seg_%d" % t] = np.array([int(has_new), int(spawned), int(lost), len(mds)], np.int32)
            out["md_%d" % t] = np.array([[m["id"], m["superPixelCount"], m["avgConfidence"], m["depthMean"], m["depthStd"]]
                                         for m in mds], np.float64)
        print("objects frame", t, [m.id for m in op.models], op.last_seg[1:] if op.last_seg else None, flush=True)
    np.savez_compressed(os.path.join(HERE, "objects4_oracle.npz"), frames=OBJ_FRAMES, **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["plane", "objects"]
    if "plane" in what:
        plane()
    if "objects" in what:
        objects()
