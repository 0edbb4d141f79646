"""GPU parity of the motion segmentation: cfb_segmentation_* through the C
ABI against oracle/segment.c on the same seeded inputs.  Everything is integer / index work or
order-pinned float sums, so the bar is BIT-EXACT: SLIC labels, super-pixel maps, unaries, the low-res
label map, the full-resolution mask and ModelData."""
import numpy as np
import pytest

import orc
import seg_cases

pytestmark = pytest.mark.gpu


def _run_gpu(c, allow_new=True, params=None):
    import torch
    import cofusion_b200 as cfb
    H, W = c["depth"].shape
    seg = cfb.Segmentation(W, H)
    dev = "cuda:0"
    rgb = torch.from_numpy(c["rgb"]).to(dev)
    depth = torch.from_numpy(c["depth"]).to(dev)
    icp = [torch.from_numpy(a).to(dev) for a in c["icp"]]
    vc = [torch.from_numpy(a).to(dev) for a in c["vc"]]
    torch.cuda.synchronize()
    full, mds, has_new = seg.perform_crf(rgb, depth, c["model_ids"], icp, vc, c["next_id"], allow_new, params)
    torch.cuda.synchronize()
    return seg, full.cpu().numpy(), mds, has_new


def _check(c, allow_new=True, oprm=None, gprm=None):
    seg_o, mds_o, new_o, lab_o, unary_o, low_o = orc.segment_crf(c["rgb"], c["depth"], c["model_ids"], c["icp"],
                                                                 c["vc"], c["next_id"], allow_new, oprm)
    seg, full, mds, has_new = _run_gpu(c, allow_new, gprm)
    assert np.array_equal(seg.view(0), lab_o), "SLIC labels"
    assert np.array_equal(seg.view(1), np.bincount(lab_o.ravel(), minlength=seg.N)), "super-pixel counts"
    u = seg.view(2)
    assert u.shape == unary_o.shape and np.array_equal(u.view(np.uint32), unary_o.view(np.uint32)), "unaries"
    assert np.array_equal(seg.view(3), low_o.ravel()), "low-res label map"
    assert np.array_equal(full, seg_o), "full-resolution mask"
    assert has_new == new_o and len(mds) == len(mds_o)
    for a, b in zip(mds, mds_o):
        got = a.astuple()
        want = (b["id"], b["superPixelCount"], b["avgConfidence"], b["depthMean"], b["depthStd"], b["top"],
                b["right"], b["bottom"], b["left"])
        assert np.array_equal(np.array(got, np.float64), np.array(want, np.float64)), (got, want)
    return seg, full, mds, has_new


def test_room_new_label_640x480():
    seg, full, mds, has_new = _check(seg_cases.room_case())
    assert has_new and mds[-1].id == 1


def test_room_320x240_and_static_scene():
    _check(seg_cases.room_case(320, 240))
    _, full, mds, has_new = _check(seg_cases.room_case(320, 240, err=0.001))
    assert not has_new and not full.any()


def test_two_models_third_label():
    seg, full, mds, has_new = _check(seg_cases.two_model_case())
    assert [m.id for m in mds[:2]] == [0, 1]
    _check(seg_cases.two_model_case(320, 240))


def test_allow_new_false_and_size_gate():
    import cofusion_b200 as cfb
    c = seg_cases.room_case(320, 240)
    _check(c, allow_new=False)
    op, gp = orc.OrcSegParams.default(), cfb.SegParams.default()
    op.minRelSizeNew = gp.minRelSizeNew = 0.3
    _, full, mds, has_new = _check(c, True, op, gp)
    assert not has_new and (full == 255).any()


def test_noise_and_depth_holes():
    # empty thresholded super-pixels (Slic.h:117-122 quirk), random unaries, many small components
    _check(seg_cases.noise_case())
    _check(seg_cases.noise_case(640, 480, seed=9))


def test_changed_crf_parameters():
    import cofusion_b200 as cfb
    c = seg_cases.two_model_case(320, 240, seed=4)
    op, gp = orc.OrcSegParams.default(), cfb.SegParams.default()
    for p in (op, gp):
        p.crfIterations = 3
        p.weightAppearance = 4.0
        p.weightSmoothness = 5.0
        p.unaryThresholdNew = 4.0
        p.scaleFeaturesRGB = 0.05
    _check(c, True, op, gp)


def test_default_params_match_oracle():
    import cofusion_b200 as cfb
    op, gp = orc.OrcSegParams.default(), cfb.SegParams.default()
    for name, _ in cfb.SegParams._fields_:
        assert getattr(op, name) == getattr(gp, name), name


def test_graph_replay_on_a_side_stream_matches():
    """On a capturable stream the launch sequence is replayed as a CUDA graph; results must not change,
    also when the arguments change between calls (graph re-capture)."""
    import torch
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        c = seg_cases.room_case(320, 240)
        for _ in range(2):
            _check(c)
        _check(c, allow_new=False)
        _check(seg_cases.two_model_case(320, 240))
    torch.cuda.synchronize()
