"""-m gpu: the CUDA surfel stage (initialise / predictIndices / fuse / clean / combinedPredict / fillIn)
against the CPU restatement of the reference's GLSL (oracle/surfel.c).  Integer outputs (index maps,
ids, counts, times, colours) must be BIT-EXACT; float attribute buffers are also compared bit for
bit (both sides use the same IEEE operation sequence; tolerance 1e-4 relative is the contract)."""
import numpy as np
import pytest

import orc
import scenes
from cofusion_b200 import synth

pytestmark = pytest.mark.gpu


def rel_poses(seq):
    T0i = np.linalg.inv(seq[0][3])
    return [(T0i @ s[3]).astype(np.float32) for s in seq]


def assert_same(a, b, what, tol=0.0):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        if tol == 0.0:
            bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
            assert not bad.any(), "%s: %d of %d floats differ, max abs %g" % (
                what, bad.sum(), a.size, np.nanmax(np.abs(a[bad].astype(np.float64) - b[bad])))
        else:
            assert np.allclose(a, b, rtol=tol, atol=tol, equal_nan=True), what
    else:
        assert np.array_equal(a, b), "%s: %d of %d differ" % (what, (a != b).sum(), a.size)


@pytest.fixture(scope="module", params=[(160, 120), (640, 480)], ids=["160x120", "640x480"])
def seq(request):
    W, H = request.param
    K = scenes.scaled_K(W)
    s = list(synth.room_sequence(5, W, H, K, noise=True, n_boxes=2))
    return W, H, K, s, rel_poses(s)


def run_both(seq, conf=0.9, frames=4, check=True):
    import cofusion_b200 as cfb
    W, H, K, s, poses = seq
    cap = 1 << 19
    om = orc.OrcMap(W, H, K, cap)
    ctx = cfb.Context(W, H, K)
    gm = cfb.Model(ctx, 0, conf, cap, True)
    for t in range(frames):
        _, rgb, d, _, ids = s[t]
        mask = np.where(ids == 1, 1, 0).astype(np.uint8)  # object 1 is not background -> exercises maskID
        pose = poses[t]
        tick = t + 1
        ctx.upload_frame(np.ascontiguousarray(rgb), np.ascontiguousarray(d), np.ascontiguousarray(mask))
        ctx.preprocess(5.0)
        df = orc.bilateral(d, 5.0)
        if check:
            assert_same(ctx.view(2), df, "bilateral t=%d" % t)
        gm.override_pose(pose)
        if t == 0:
            om.initialise(rgb, d, df, tick, 20.0)
            gm.initialise(tick, 20.0)
        else:
            w_o = orc.OrcMap.fusion_weight(pose, poses[t - 1])
            gm.override_pose(poses[t - 1])
            # lastPose := previous pose, pose := current (what performTracking leaves behind)
            import ctypes as C
            cfb.check(cfb.lib().cfb_model_override_pose(gm._h, poses[t - 1].ctypes.data_as(cfb.c_float_p)))
            _set_pose_keep_last(gm, pose)
            assert abs(gm.fusion_weight() - w_o) < 1e-6
            om.predict_indices(pose, tick)
            gm.predict_indices(tick)
            if check:
                for w_, g_ in ((0, 4), (1, 5), (2, 6), (3, 7)):
                    assert_same(gm.view(g_), om.view(w_), "index map %d t=%d" % (w_, t))
            om.fuse(pose, tick, rgb, mask, d, df, 20.0, w_o, 0)
            gm.fuse(tick, 20.0, 1.0)
            if check:
                un = om.unstable()
                assert_same(gm.view(15, len(un)), un, "unstable candidates t=%d" % t)
            om.predict_indices(pose, tick)
            gm.predict_indices(tick)
            om.clean(pose, tick, conf, 200, df, mask, 0, 3.0)
            gm.clean(tick, 200, 20.0, 3.0)
        if check:
            so, sg = om.surfels(), gm.download_map()
            assert len(so) == len(sg), "surfel count t=%d: %d vs %d" % (t, len(so), len(sg))
            assert_same(sg, so, "surfels t=%d" % t)
        om.combined_predict(pose, 20.0, conf, tick, tick)
        gm.combined_predict(20.0, tick, tick)
        om.fill_in(rgb, df)
        gm.perform_fill_in()
        if check:
            for w_, g_ in ((4, 8), (5, 9), (6, 10), (7, 11), (8, 12), (9, 13), (10, 14)):
                assert_same(gm.view(g_), om.view(w_), "predict map %d t=%d" % (w_, t))
    return om, gm, ctx


def _set_pose_keep_last(gm, pose):
    """pose <- new, lastPose stays (mirrors Model::performTracking's lastPose = pose; pose = result)"""
    import ctypes as C
    import cofusion_b200 as cfb
    lib = cfb.lib()
    if not hasattr(lib, "cfb_model_set_pose_keep_last"):
        pytest.skip("cfb_model_set_pose_keep_last missing")
    cfb.check(lib.cfb_model_set_pose_keep_last(gm._h, np.ascontiguousarray(pose, np.float32).ctypes.data_as(cfb.c_float_p)))


def test_surfel_stage_matches_oracle_over_a_sequence(seq):
    om, gm, ctx = run_both(seq, conf=0.9, frames=4)
    assert om.count > 0.9 * seq[0] * seq[1]
    # after 4 frames the splat prediction is dense enough that fill-in is not required any more
    assert (om.view(4)[..., 3] > 0).mean() > 0.5


def test_high_confidence_threshold_uses_fill_in(seq):
    if seq[0] > 160:
        pytest.skip("small case only")
    om, gm, ctx = run_both(seq, conf=10.0, frames=2)
    assert om.requires_fill_in()
    assert (om.view(4)[..., 3] > 0).mean() == 0.0  # nothing confident enough to splat yet


def test_empty_and_tiny_maps():
    import cofusion_b200 as cfb
    W, H = 160, 120
    K = scenes.scaled_K(W)
    ctx = cfb.Context(W, H, K)
    gm = cfb.Model(ctx, 0, 0.5, 4096, True)
    om = orc.OrcMap(W, H, K, 4096)
    rgb = np.full((H, W, 3), 90, np.uint8)
    d = np.zeros((H, W), np.float32)  # no depth at all -> empty map
    ctx.upload_frame(rgb, d, None)
    ctx.preprocess(5.0)
    gm.initialise(1)
    om.initialise(rgb, d, orc.bilateral(d, 5.0), 1)
    assert gm.last_count() == 0 == om.count
    pose = np.eye(4, dtype=np.float32)
    gm.predict_indices(1)
    om.predict_indices(pose, 1)
    assert_same(gm.view(4), om.view(0), "empty index map")
    gm.combined_predict(20.0, 1, 1)
    gm.perform_fill_in()
    assert (gm.view(8) == 0).all()
    # capacity clamp: more valid pixels than max_surfels
    d[:] = 1.5
    ctx.upload_frame(rgb, d, None)
    ctx.preprocess(5.0)
    gm.initialise(1)
    om.initialise(rgb, d, orc.bilateral(d, 5.0), 1)
    assert gm.last_count() == 4096 == om.count
    assert_same(gm.download_map(), om.surfels(), "clamped init")
