"""-m gpu: CoFusion::processFrame end to end (upload -> bilateral -> track -> predict -> fuse -> clean
-> predict) through the C ABI against the same sequencing driven through the CPU oracle."""
import numpy as np
import pytest

import scenes
from cofusion_b200 import synth
from orc_pipeline import OraclePipeline

pytestmark = pytest.mark.gpu


def run(W, H, seq, frames, conf=10.0):
    import cofusion_b200 as cfb
    K = scenes.scaled_K(W)
    p = cfb.CoFusionParams.default(1 << 20)
    p.confGlobalInit = conf
    cf = cfb.CoFusion(W, H, K, p)
    op = OraclePipeline(W, H, K, 1 << 20, conf_global=conf)
    T0i = np.linalg.inv(seq[0][3])
    out = []
    for t in range(frames):
        rgb, d, T = seq[t][1], seq[t][2], seq[t][3]
        cf.process_frame(np.ascontiguousarray(rgb), np.ascontiguousarray(d))
        op.process_frame(rgb, d)
        gt = (T0i @ T).astype(np.float32)
        out.append((cf.pose(0).copy(), op.pose.copy(), gt, cf.model(0).last_count(), op.map.count))
    return out, cf, op


@pytest.mark.parametrize("conf", [10.0, 0.9], ids=["confG10_fill_in_path", "conf0.9_model_path"])
def test_room_sequence_matches_oracle(conf):
    W, H = 320, 240
    seq = list(synth.room_sequence(7, W, H, scenes.scaled_K(W), noise=True))
    res, cf, op = run(W, H, seq, 7, conf)
    for t, (pg, po, gt, ng, no) in enumerate(res):
        assert np.abs(pg - po).max() < 1e-4, (t, pg, po)          # pose parity, north_star tolerance
        assert abs(ng - no) <= max(2, 2e-3 * no), (t, ng, no)      # surfel count (pose differs at 1e-6)
    # the fused pyramid builders (RGBDOdometry::initAll) leave the same buffers as the oracle's
    # function-by-function construction
    m = cf.model(0)
    # (model-derived buffers 4,5,6 only on the fill-in path, where the "model" is the previous frame;
    #  with a splat prediction they inherit the 1e-6 pose difference between the two pipelines)
    views = ((0, 0.0), (1, 3e-6), (7, 0.0)) + (((4, 0.0), (5, 0.0), (6, 0.0)) if conf >= 10.0 else ())
    for which, tol in views:
        for lvl in range(3):
            a, b = m.odometry_view(which, lvl), op.odom.view(which, lvl)
            if a.dtype == np.float32:
                ok, msg = scenes.nan_equal(a, b, tol=tol)
                # the last tracked frame used slightly different poses/predictions (1e-6): model-side
                # buffers (2,3) are not compared, frame-side ones must agree
                assert ok, (which, lvl, msg)
            else:
                assert np.array_equal(a, b), (which, lvl)
    # tracking follows the camera (1 deg / 8.7 mm per frame): error to ground truth stays small
    pg, _, gt, _, _ = res[-1]
    assert np.abs(pg - gt).max() < 0.02, (pg, gt)


def test_static_plane_config0_drift():
    """BASELINE config[0]: static textured plane, ICP+RGB converges; drift bounded (the photometric
    term resolves in-plane motion to ~0.5 px = 2 mm at 2 m)."""
    W, H = 640, 480
    seq = [(ts, rgb, d, T) for ts, rgb, d, T in synth.plane_sequence(6, W, H, synth.K_DEFAULT)]
    res, cf, op = run(W, H, seq, 6)
    for t, (pg, po, gt, ng, no) in enumerate(res):
        assert np.abs(pg - po).max() < 1e-4, (t, pg, po)
    step_err = [np.abs(res[t][0] - res[t][2]).max() for t in range(len(res))]
    assert max(step_err) < 6e-3 * len(res), step_err
