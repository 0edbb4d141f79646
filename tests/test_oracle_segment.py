"""CPU invariants of the segmentation oracle (oracle/segment.c).  The oracle is "parity unpinned"
(gSLICr / densecrf are not in the reference tree), so these tests pin what CAN be pinned: the parts
restated from Core/Segmentation line by line behave as the reference code reads, and the restated
library parts satisfy their published definitions."""
import numpy as np
import pytest

import orc
import seg_cases


@pytest.fixture(scope="module")
def room():
    c = seg_cases.room_case()
    out = orc.segment_crf(c["rgb"], c["depth"], c["model_ids"], c["icp"], c["vc"], c["next_id"], True)
    return c, out


def test_slic_labels_are_local_and_cover_the_grid():
    c = seg_cases.room_case()
    lab = orc.slic(c["rgb"])
    H, W = lab.shape
    mx = W // 16
    assert lab.min() >= 0 and lab.max() < mx * (H // 16)
    # a pixel can only be assigned to a centre of its own or a neighbouring grid cell (gSLICr 3x3 search)
    ys, xs = np.mgrid[0:H, 0:W]
    assert np.all(np.abs(lab % mx - xs // 16) <= 1) and np.all(np.abs(lab // mx - ys // 16) <= 1)
    counts = np.bincount(lab.ravel(), minlength=mx * (H // 16))
    assert counts.sum() == W * H and counts.min() > 0
    assert np.array_equal(lab, orc.slic(c["rgb"]))  # deterministic


def test_slic_is_channel_order_invariant():
    # Slic::setInputImage swaps R and B (Slic.cpp:52-60); the distance is symmetric in the channels
    c = seg_cases.room_case(320, 240)
    assert np.array_equal(orc.slic(c["rgb"]), orc.slic(np.ascontiguousarray(c["rgb"][..., ::-1])))


def test_moving_box_becomes_a_new_label(room):
    c, (seg, mds, has_new, lab, unary, low) = room
    assert has_new and len(mds) == 2 and mds[1]["id"] == 1 and mds[0]["id"] == 0
    box = c["ids"] == 1
    assert ((seg == 1) & box).sum() > 0.4 * box.sum()
    assert ((seg == 1) & ~box).sum() < 0.1 * (seg == 1).sum() + 0.02 * seg.size
    # ModelData bookkeeping (Segmentation.cpp:570-649)
    assert mds[0]["superPixelCount"] == (low == 0).sum() and mds[1]["superPixelCount"] == (low == 1).sum()
    N = low.size
    assert N * 0.015 <= mds[1]["superPixelCount"] <= N * 0.4
    ys, xs = np.nonzero(low == 1)
    assert (mds[1]["left"], mds[1]["right"]) == (xs.min() * 16 + 8, xs.max() * 16 + 8)
    assert (mds[1]["top"], mds[1]["bottom"]) == (ys.min() * 16 + 8, ys.max() * 16 + 8)
    assert abs(mds[0]["avgConfidence"] - 10.0) < 1e-3
    assert np.array_equal(seg, low.ravel()[lab])  # Slic::upsample


def test_unaries_follow_segmentation_cpp(room):
    c, (seg, mds, has_new, lab, unary, low) = room
    # unary(new) = max(thresholdNew - wErr * lowest, 0.01), unary(m) = wErr * err/range, floor 1e-5
    assert unary.shape[1] == 2 and unary.min() >= 1e-5
    u_new = np.maximum(np.float32(5.5) - unary[:, 0], np.float32(0.01))
    assert np.allclose(unary[:, 1], u_new, rtol=1e-6, atol=1e-6)


def test_static_scene_has_no_new_label():
    c = seg_cases.room_case(320, 240, err=0.001)
    seg, mds, has_new, *_ = orc.segment_crf(c["rgb"], c["depth"], [0], c["icp"], c["vc"], 1, True)
    assert not has_new and len(mds) == 1 and np.all(seg == 0)
    assert mds[0]["superPixelCount"] == (320 // 16) * (240 // 16)


def test_allow_new_false_and_size_gate():
    c = seg_cases.room_case(320, 240)
    seg, mds, has_new, *_ = orc.segment_crf(c["rgb"], c["depth"], [0], c["icp"], c["vc"], 1, False)
    assert not has_new and len(mds) == 1 and set(np.unique(seg)) <= {0}
    prm = orc.OrcSegParams.default()
    prm.minRelSizeNew = 0.3  # the box is far smaller than 30 % of the image -> rejected, relabelled 255
    seg, mds, has_new, *_ = orc.segment_crf(c["rgb"], c["depth"], [0], c["icp"], c["vc"], 1, True, prm)
    assert not has_new and len(mds) == 1 and set(np.unique(seg)) <= {0, 255}


def test_two_models_and_a_third_label():
    c = seg_cases.two_model_case(320, 240)
    seg, mds, has_new, lab, unary, low = orc.segment_crf(c["rgb"], c["depth"], c["model_ids"], c["icp"], c["vc"],
                                                        c["next_id"], True)
    assert unary.shape[1] == 3 and [m["id"] for m in mds[:2]] == [0, 1]
    b1, b2 = c["ids"] == 1, c["ids"] == 2
    assert ((seg == 1) & b1).sum() > 0.4 * b1.sum()
    if has_new:
        assert mds[2]["id"] == 2 and ((seg == 2) & b2).sum() > 0
    # onlyKeepLargest: every object label owns exactly one connected component of the low-res map
    for lid in (1, 2):
        m = low == lid
        if m.any():
            import scipy.ndimage as ndi
            assert ndi.label(m)[1] == 1


def test_depth_hole_quirk_path_is_finite():
    c = seg_cases.noise_case()
    seg, mds, has_new, lab, unary, low = orc.segment_crf(c["rgb"], c["depth"], [0], c["icp"], c["vc"], 1, True)
    assert np.isfinite(unary).all() and set(np.unique(seg)) <= {0, 1, 255}
