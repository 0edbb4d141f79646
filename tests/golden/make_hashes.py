"""Regenerates tests/golden/oracle_hashes.json: SHA-256 of oracle outputs that depend on nothing but
IEEE + - x / sqrt, fma and rint (no libm transcendentals), so that they are identical on every x86-64
host.  They pin the restated GL / segmentation semantics against accidental change; the GPU tests pin
the product to the oracle, so the product is pinned transitively.

    python tests/golden/make_hashes.py
"""
import hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
import orc
import seg_cases
from cofusion_b200 import synth


def h(*arrays):
    m = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        m.update(str(a.dtype).encode() + str(a.shape).encode())
        m.update(a.tobytes())
    return m.hexdigest()


def compute():
    out = {}
    # bilateral filter + depth pyramid, 640x480 room frame (a1, a2)
    seq = list(synth.room_sequence(3, 640, 480, synth.K_DEFAULT, noise=True))
    _, rgb, d, _, _ = seq[2]
    out["inputs"] = h(rgb, d)  # if the renderer differs on this host (numpy SIMD dispatch) the test skips
    df = orc.bilateral(d, 5.0)
    out["bilateral_640x480"] = h(df)
    out["pyr_down_gauss_f"] = h(orc.pyr_down_f(df))
    # SLIC + CRF segmentation, 640x480, one model + new label (a19)
    c = seg_cases.room_case()
    out["inputs"] += h(c["rgb"], c["depth"], c["icp"][0])
    seg, mds, has_new, lab, unary, low = orc.segment_crf(c["rgb"], c["depth"], c["model_ids"], c["icp"], c["vc"], c["next_id"], True)
    out["slic_labels_640x480"] = h(lab)
    out["seg_unary_640x480"] = h(unary)
    out["seg_mask_640x480"] = h(seg, low)
    out["seg_modeldata_640x480"] = h(np.array([[m[k] for k in ("id", "superPixelCount", "avgConfidence", "depthMean", "depthStd",
                                                               "top", "right", "bottom", "left")] for m in mds], np.float64))
    c2 = seg_cases.two_model_case(320, 240)
    out["inputs"] += h(c2["rgb"], c2["depth"], c2["icp"][0], c2["icp"][1], c2["vc"][1])
    seg2, mds2, _, lab2, unary2, low2 = orc.segment_crf(c2["rgb"], c2["depth"], c2["model_ids"], c2["icp"], c2["vc"], c2["next_id"], True)
    out["seg_two_models_320x240"] = h(lab2, unary2, seg2, low2)
    # surfel stage with GIVEN poses (initialise, index map, fuse, clean, splat prediction), 320x240 (a13-a18, a20)
    fx, fy, cx, cy = synth.K_DEFAULT
    K = (fx / 2, fy / 2, cx / 2, cy / 2)
    W, H = 320, 240
    seq = list(synth.room_sequence(3, W, H, K, noise=True, n_boxes=1))
    T0i = np.linalg.inv(seq[0][3])
    poses = [(T0i @ T).astype(np.float32) for _, _, _, T, _ in seq]
    out["inputs"] += h(*[s_[1] for s_ in seq], *[s_[2] for s_ in seq], *poses)
    m = orc.OrcMap(W, H, K, 1 << 18)
    mask = np.zeros((H, W), np.uint8)
    for t, (_, rgb, d, T, _) in enumerate(seq):
        df = orc.bilateral(d, 5.0)
        pose = poses[t]
        if t == 0:
            m.initialise(rgb, d, df, 1, 20.0)
        else:
            m.predict_indices(pose, t + 1, 20.0, 200)
            m.fuse(pose, t + 1, rgb, mask, d, df, 20.0, 0.5, 0)
            m.predict_indices(pose, t + 1, 20.0, 200)
            m.clean(pose, t + 1, 0.9, 200, df, mask, 0, 3.0)
        m.combined_predict(pose, 20.0, 0.9, t + 1, t + 1, 200)
    out["surfel_map_3_frames_320x240"] = h(m.surfels())
    out["surfel_count"] = int(m.count)
    out["index_map_320x240"] = h(m.view(0))
    out["splat_prediction_320x240"] = h(m.view(4), m.view(5), m.view(6))
    return out


if __name__ == "__main__":
    res = compute()
    json.dump(res, open(os.path.join(HERE, "oracle_hashes.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))
