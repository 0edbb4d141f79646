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


def test_lattice_witness_agrees_with_exact_kernels():
    """Second witness for the dense CRF (row a19): the library the reference links (densecrf) evaluates the
    kernel products on a permutohedral lattice, the oracle of record and the CUDA kernels evaluate them
    exactly.  oracle/lattice.c restates the lattice; the two realisations must take the same decisions and
    agree on (nearly) every label -- the measured figures are in profiles/crf_witness_r02.txt."""
    import ctypes as C
    o = orc.orc()
    # the lattice reproduces the Gaussian it approximates up to a constant scale (removed by the symmetric normalisation)
    o.orc_lattice_create.restype = C.c_void_p
    o.orc_lattice_create.argtypes = [C.c_void_p, C.c_int, C.c_int]
    o.orc_lattice_destroy.argtypes = [C.c_void_p]
    o.orc_lattice_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    W, H = 20, 15
    f = np.array([[i / 2.0, j / 2.0] for j in range(H) for i in range(W)], np.float32)
    L = o.orc_lattice_create(f.ctypes.data, 2, W * H)
    q = np.random.default_rng(0).random((W * H, 2)).astype(np.float32)
    out = np.zeros_like(q)
    o.orc_lattice_compute(L, q.ctypes.data, 2, out.ctypes.data)
    o.orc_lattice_destroy(L)
    K = np.exp(-0.5 * ((f[:, None, :] - f[None, :, :]) ** 2).sum(-1))
    ex = K @ q
    scale = (out * ex).sum() / (ex * ex).sum()
    assert 0.7 < scale < 1.1 and np.linalg.norm(out - scale * ex) / np.linalg.norm(scale * ex) < 0.08
    for case, allow_new in ((seg_cases.room_case(320, 240), True), (seg_cases.noise_case(320, 240), True)):
        res = {}
        for mode in (0, 1):
            o.orc_segment_set_kernel_mode(mode)
            res[mode] = orc.segment_crf(case["rgb"], case["depth"], case["model_ids"], case["icp"], case["vc"],
                                        case["next_id"], allow_new)
        o.orc_segment_set_kernel_mode(0)
        (seg_e, md_e, new_e, _, _, low_e), (seg_l, md_l, new_l, _, _, low_l) = res[0], res[1]
        assert new_e == new_l and len(md_e) == len(md_l)
        assert (low_e == low_l).mean() >= 0.98 and (seg_e == seg_l).mean() >= 0.98
