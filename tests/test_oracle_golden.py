"""-m "not gpu": pins the CPU restatement (oracle/tracker.c) against outputs of the REFERENCE's own
CUDA kernels (Core/Cuda/reduce.cu, cudafuncs.cu compiled unmodified for sm_100a and run on a B200 by
tests/golden/make_golden.py -> tests/golden/tracker_ref_sm100a.npz).

Tolerances: the reference build uses --prec-div=false --prec-sqrt=false --ftz=true and FMA
contraction, the oracle plain IEEE: float images agree to a few ulp, integer images exactly up to
1-LSB flips at exact .0 boundaries, reduction sums to 1e-4 relative (f32 tree vs f64 sums)."""
import os

import numpy as np
import pytest

import orc
import scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracker_ref_sm100a.npz")
ANGLE = float(np.sin(np.deg2rad(20.0)))


@pytest.fixture(scope="module")
def g():
    if not os.path.exists(G):
        pytest.skip("golden fixture missing")
    return dict(np.load(G))


def close_nan(a, b, tol):
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    return np.abs(a[m] - b[m]).max() <= tol * max(1.0, np.abs(b[m]).max())


def test_image_preparation_restatement_matches_reference_kernels(g):
    K = tuple(g["K"])
    df, grey = g["in_depth"], g["in_grey"]
    assert close_nan(orc.pyr_down_f(df), g["pyr_f"], 2e-6)
    assert np.array_equal(orc.pyr_down_u8(grey), g["pyr_u8"])
    dx, dy = orc.derivative_images(grey)
    assert np.abs(dx.astype(int) - g["dx"]).max() <= 1 and (dx != g["dx"]).mean() < 2e-3
    assert np.abs(dy.astype(int) - g["dy"]).max() <= 1 and (dy != g["dy"]).mean() < 2e-3
    H = df.shape[0]
    v = orc.create_vmap(df, K, 20.0)
    assert np.array_equal(np.isnan(v[:H]), np.isnan(g["vmap"][:H]))
    m = ~np.isnan(v[:H])
    for k in range(3):
        assert np.abs(v[k * H:(k + 1) * H][m] - g["vmap"][k * H:(k + 1) * H][m]).max() < 2e-6 * 5
    n = orc.create_nmap(g["vmap"])
    mn = ~np.isnan(g["nmap"][:H])
    assert np.array_equal(np.isnan(n[:H]), np.isnan(g["nmap"][:H]))
    for k in range(3):
        assert np.abs(n[k * H:(k + 1) * H][mn] - g["nmap"][k * H:(k + 1) * H][mn]).max() < 5e-6
    cv, cn = orc.copy_maps(g["in_v4"], g["in_n4"])
    assert close_nan(cv, g["copy_v"], 0) and close_nan(cn, g["copy_n"], 0)
    H2 = H // 2
    rv = orc.resize_map(g["copy_v"], False)
    rn = orc.resize_map(g["copy_n"], True)
    for a, b, tol in ((rv, g["resize_v"], 2e-6), (rn, g["resize_n"], 5e-6)):
        mm = ~np.isnan(b[:H2])
        assert np.array_equal(np.isnan(a[:H2]), np.isnan(b[:H2]))
        for k in range(3):
            assert np.abs(a[k * H2:(k + 1) * H2][mm] - b[k * H2:(k + 1) * H2][mm]).max() <= tol * 5
    T0 = g["in_T0"]
    tv, tn = orc.transform_maps(g["copy_v"], g["copy_n"], T0[:3, :3], T0[:3, 3])
    mt = ~np.isnan(g["tr_v"][:H])
    for k in range(3):
        assert np.abs(tv[k * H:(k + 1) * H][mt] - g["tr_v"][k * H:(k + 1) * H][mt]).max() < 5e-6
        assert np.abs(tn[k * H:(k + 1) * H][mt] - g["tr_n"][k * H:(k + 1) * H][mt]).max() < 5e-6
    assert np.array_equal(orc.vertices_to_depth(g["in_v4"], 6.0), g["v2d"], equal_nan=True)
    assert close_nan(orc.project_cloud(g["v2d"], K), g["cloud"], 2e-6)


def test_reduction_steps_restatement_matches_reference_kernels(g):
    K = tuple(g["K"])
    T0, T = g["in_T0"].astype(np.float64), g["in_T"].astype(np.float64)
    A, b, r, _ = orc.icp_step(T[:3, :3], T[:3, 3], g["vmap"], g["nmap"], g["in_Rpi"], T0[:3, 3], K, g["tr_v"],
                              g["tr_n"], 0.10, ANGLE)
    assert abs(r[1] - g["icp_res"][1]) <= max(2, 2e-3 * g["icp_res"][1])
    assert scenes.relerr(A, g["icp_A"]) < 1e-3 and scenes.relerr(b, g["icp_b"]) < 1e-3
    c, s, n = orc.rgb_residual(64.0, g["dx"], g["dy"], g["v2d"], g["v2d"], g["in_grey0"], g["in_grey"], 0.07,
                               g["in_kt"], g["in_krk"])
    assert abs(n - int(g["res_count"])) <= max(2, 2e-3 * n) and abs(s - int(g["res_sigma"])) <= max(500, 5e-3 * abs(s))
    A, b = orc.rgb_step(g["res_corres"], float(g["res_count"]), g["cloud"], K, g["dx"], g["dy"], 0.125)
    assert scenes.relerr(A, g["rgb_A"]) < 1e-4 and scenes.relerr(b, g["rgb_b"]) < 1e-4
    A, b, r = orc.so3_step(g["in_grey0"], g["in_grey"], g["in_so3_H"], g["in_so3_Kinv"], g["in_so3_KR"])
    assert r[1] == g["so3_res"][1]
    assert scenes.relerr(A, g["so3_A"]) < 1e-4 and scenes.relerr(b, g["so3_b"]) < 1e-4


def test_full_tracker_restatement_matches_reference_kernel_loop(g):
    """getIncrementalTransformation through the restated host loop: CPU steps vs the reference's
    CUDA steps (same loop, oracle/ref_driver.cu) -- pose within 1e-4."""
    case = scenes.room_pair(160, 120)
    if abs(float(case["d1"].astype(np.float64).sum()) - float(g["track_in_d1_sum"])) > 1e-3:
        pytest.skip("synthetic generator changed since the fixture was made")
    oo, _ = scenes.oracle_odometry(case)
    pose, st, _, _ = oo.track(case["T0"])
    assert np.abs(pose - g["track_pose"]).max() < 1e-4
    assert abs(st.lastICPCount - float(g["track_icp_count"])) <= 3e-3 * st.lastICPCount


def test_oracle_tracker_converges_on_a_well_conditioned_view():
    case = scenes.room_pair(160, 120, noise=False, holes=False)
    oo, _ = scenes.oracle_odometry(case)
    pose, st, _, _ = oo.track(case["T0"])
    assert np.abs(pose - case["T1"]).max() < 3e-3
    assert st.lastICPCount > 0.5 * 160 * 120


def test_oracle_outputs_match_the_committed_hashes():
    """tests/golden/oracle_hashes.json (made by tests/golden/make_hashes.py): SHA-256 of oracle outputs of
    the GL-restated stages and the segmentation -- everything whose arithmetic is IEEE + - x / sqrt, fma,
    rint only.  Pins the frozen semantics (F1-F6, the CRF summation order, the deterministic exp) against
    accidental change; the GPU tests pin the kernels to the oracle."""
    import importlib.util
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_hashes", os.path.join(here, "make_hashes.py"))
    mh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mh)
    want = json.load(open(os.path.join(here, "oracle_hashes.json")))
    got = mh.compute()
    if got["inputs"] != want["inputs"]:
        pytest.skip("the synthetic renderer produced different inputs on this host (numpy SIMD dispatch)")
    assert got == want, {k: (got[k], want[k]) for k in want if got.get(k) != want[k]}
