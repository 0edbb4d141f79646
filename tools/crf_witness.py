"""Exact-kernel CRF (oracle of record, = the CUDA kernels bit for bit) vs the permutohedral-lattice CRF of
densecrf (oracle/lattice.c, second witness): label agreement and ModelData deltas on the segmentation
cases.  CPU only.  `python tools/crf_witness.py` prints the table recorded in DESIGN.md section 2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import orc  # noqa: E402
import seg_cases  # noqa: E402


def run(case, allow_new=True, params=None):
    out = {}
    for mode, name in ((0, "exact"), (1, "lattice")):
        orc.orc().orc_segment_set_kernel_mode(mode)
        out[name] = orc.segment_crf(case["rgb"], case["depth"], case["model_ids"], case["icp"], case["vc"],
                                    case["next_id"], allow_new, params)
    orc.orc().orc_segment_set_kernel_mode(0)
    return out


def compare(name, case, allow_new=True, params=None):
    r = run(case, allow_new, params)
    (seg_e, md_e, new_e, _, _, low_e), (seg_l, md_l, new_l, _, _, low_l) = r["exact"], r["lattice"]
    row = {"case": name, "labels": len(md_e), "low_agree": float((low_e == low_l).mean()),
           "full_agree": float((seg_e == seg_l).mean()), "new_exact": new_e, "new_lattice": new_l}
    d = []
    for a, b in zip(md_e, md_l):
        d.append({"id": a["id"], "spx": (a["superPixelCount"], b["superPixelCount"]),
                  "depthMean": (round(a["depthMean"], 4), round(b["depthMean"], 4)),
                  "avgConf": (round(a["avgConfidence"], 4), round(b["avgConfidence"], 4))})
    row["model_data"] = d
    return row


def cases():
    yield "room 640x480, 1 model + new", seg_cases.room_case(640, 480), True, None
    yield "room 320x240, 1 model + new", seg_cases.room_case(320, 240), True, None
    yield "two models 640x480 + new", seg_cases.two_model_case(640, 480), True, None
    yield "two models 640x480, no new label", seg_cases.two_model_case(640, 480), False, None
    yield "noise 320x240 (depth holes)", seg_cases.noise_case(320, 240), True, None
    p = orc.OrcSegParams.default()
    p.weightAppearance, p.weightSmoothness = 3.0, 4.0
    yield "room 640x480, weights 3 / 4", seg_cases.room_case(640, 480, seed=2), True, p


def main():
    rows = [compare(*c) for c in cases()]
    print("%-36s %6s %10s %10s %8s" % ("case", "labels", "low-res =", "full-res =", "new e/l"))
    for r in rows:
        print("%-36s %6d %9.2f%% %9.2f%% %5s/%-5s" % (r["case"], r["labels"], 100 * r["low_agree"], 100 * r["full_agree"],
                                                      r["new_exact"], r["new_lattice"]))
        for d in r["model_data"]:
            print("      id %3d: super-pixels %5d / %-5d  depthMean %.4f / %.4f  avgConf %.4f / %.4f" % (
                d["id"], d["spx"][0], d["spx"][1], d["depthMean"][0], d["depthMean"][1], d["avgConf"][0], d["avgConf"][1]))
    return rows


if __name__ == "__main__":
    main()
