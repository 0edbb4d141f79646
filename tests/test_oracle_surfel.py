"""-m "not gpu": invariants of the surfel-stage restatement (oracle/surfel.c). The reference has no
golden vectors for its GL stage (parity unpinned), so these pin the restatement's own semantics."""
import numpy as np

import orc
import scenes
from cofusion_b200 import synth


def _frame(W=160, H=120, t=0):
    K = scenes.scaled_K(W)
    seq = list(synth.room_sequence(t + 2, W, H, K, noise=False))
    return K, seq


def test_initialise_emits_one_surfel_per_valid_pixel_in_column_major_order():
    W, H = 160, 120
    K, seq = _frame(W, H)
    _, rgb, d, _, _ = seq[0]
    d = d.copy()
    d[10:20, 30:40] = 0
    df = orc.bilateral(d, 5.0)
    m = orc.OrcMap(W, H, K, 1 << 16)
    m.initialise(rgb, d, df, 1)
    assert m.count == int((d > 0).sum())
    s = m.surfels()
    # first surfel is pixel (0,0), second (0,1) [column-major]: x equal, y increasing
    assert abs(s[0, 0] - s[1, 0]) < 0.02 and s[1, 1] > s[0, 1]
    assert (s[:, 6] == 1).all() and (s[:, 7] == 1).all() and (s[:, 5] == 0).all()   # init_unstable.vert
    rgbv = s[:, 4].astype(np.int64)
    assert ((rgbv >> 16) & 255).max() <= 255 and (s[:, 3] > 0).all() and (s[:, 3] <= 1.0).all()
    assert np.allclose(np.linalg.norm(s[:, 8:11], axis=1), 1.0, atol=1e-5)


def test_index_map_depth_test_and_id_zero_quirk():
    W, H = 160, 120
    K, seq = _frame(W, H)
    _, rgb, d, _, _ = seq[0]
    m = orc.OrcMap(W, H, K, 1 << 16)
    m.initialise(rgb, d, orc.bilateral(d, 5.0), 1)
    m.predict_indices(np.eye(4, dtype=np.float32), 1)
    idx = m.view(0)
    vc = m.view(1)
    assert idx.max() < m.count
    # every surfel was generated from its own pixel: it must win exactly that pixel
    hit = idx > 0
    assert hit.mean() > 0.95
    ys, xs = np.nonzero(hit)
    ids = idx[hit]
    assert np.array_equal(ids // H, xs) and np.array_equal(ids % H, ys)
    # surfel 0 lands on pixel (0,0) but is indistinguishable from "empty" (index_map.vert:49)
    assert idx[0, 0] == 0 and vc[0, 0, 2] > 0
    # time window: nothing is visible once time - lastTime > timeDelta
    m.predict_indices(np.eye(4, dtype=np.float32), 500, 20.0, 200)
    assert (m.view(0) == 0).all()


def test_fuse_merges_a_repeated_frame_and_clean_drops_the_merge_records():
    W, H = 160, 120
    K, seq = _frame(W, H)
    _, rgb, d, _, _ = seq[0]
    df = orc.bilateral(d, 5.0)
    mask = np.zeros((H, W), np.uint8)
    m = orc.OrcMap(W, H, K, 1 << 16)
    m.initialise(rgb, d, df, 1)
    n0 = m.count
    pose = np.eye(4, dtype=np.float32)
    m.predict_indices(pose, 2)
    m.fuse(pose, 2, rgb, mask, d, df, 20.0, 1.0, 0)
    un = m.unstable()
    assert len(un) <= (W // 2) * (H // 2) and len(un) > 0.8 * (W // 2) * (H // 2)   # eligible = parity-0 pixels
    assert ((un[:, 7] == -1) | (un[:, 7] == -2)).all()
    merged = int((un[:, 7] == -1).sum())
    assert merged > 0.9 * len(un)
    s = m.surfels()
    assert int((s[:, 7] == 2).sum()) <= merged and int((s[:, 7] == 2).sum()) > 0.8 * merged
    assert s[:, 3].max() > 1.5                      # confidences add up
    m.predict_indices(pose, 2)
    m.clean(pose, 2, 10.0, 200, df, mask, 0, 3.0)
    assert m.count == n0 + int((un[:, 7] == -2).sum())   # -1 records dropped, -2 appended
    # a foreign mask id keeps every pixel out of this model
    m2 = orc.OrcMap(W, H, K, 1 << 16)
    m2.initialise(rgb, d, df, 1)
    m2.predict_indices(pose, 2)
    m2.fuse(pose, 2, rgb, mask + 3, d, df, 20.0, 1.0, 0)
    assert len(m2.unstable()) == 0


def test_splat_prediction_and_fill_in():
    W, H = 160, 120
    K, seq = _frame(W, H)
    _, rgb, d, _, _ = seq[0]
    df = orc.bilateral(d, 5.0)
    m = orc.OrcMap(W, H, K, 1 << 16)
    m.initialise(rgb, d, df, 1)
    pose = np.eye(4, dtype=np.float32)
    m.combined_predict(pose, 20.0, 10.0, 1, 1)
    assert (m.view(4) == 0).all() and m.requires_fill_in()        # nothing above confidence 10 yet
    m.combined_predict(pose, 20.0, 0.5, 1, 1)
    img, vc = m.view(4), m.view(5)
    cover = img[..., 3] > 0
    assert cover.mean() > 0.9 and not m.requires_fill_in()
    # predicted depth reproduces the input depth where covered (half-pixel GL convention included)
    assert np.abs(vc[..., 2][cover] - d[cover]).mean() < 0.02
    assert np.abs(img[..., 0][cover].astype(int) - rgb[..., 0][cover]).mean() < 40  # sprites overlap across checker edges
    m.fill_in(rgb, df)
    fv, fi = m.view(9), m.view(8)
    assert (fv[..., 2][~cover] == df[~cover]).all() and (fi[..., :3][~cover] == rgb[~cover]).all()
    assert (fi[cover] == img[cover]).all()


def test_fusion_weight():
    I = np.eye(4, dtype=np.float32)
    assert orc.OrcMap.fusion_weight(I, I) == 1.0
    T = I.copy()
    T[0, 3] = 0.004
    assert abs(orc.OrcMap.fusion_weight(T, I) - 0.6) < 1e-5
    T[0, 3] = 0.5
    assert orc.OrcMap.fusion_weight(T, I, 2.0) == 1.0   # floor 0.5 x multiplier
