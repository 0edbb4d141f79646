"""Generates tests/golden/tracker_ref_sm100a.npz on a B200 box:

    gpurun -- 'python tests/golden/make_golden.py'

Outputs of the REFERENCE's own CUDA kernels (Core/Cuda/reduce.cu + cudafuncs.cu, compiled unmodified
for sm_100a with the reference flags into oracle/_ref/libcfref.so by `make -C oracle ref`) on small
seeded inputs.  The inputs are stored next to the outputs so the CPU-only tests can pin the oracle
restatement against the reference itself without a GPU (tests/test_oracle_golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
import scenes  # noqa: E402

ANGLE = float(np.sin(np.deg2rad(20.0)))


def main():
    ref = orc.ref()
    assert ref is not None and ref.ref_device_ok(), "needs oracle/_ref/libcfref.so and a GPU"
    out = {}
    case = scenes.room_pair(64, 48)
    K = case["K"]
    out["K"] = np.array(K, np.float32)
    df = orc.bilateral(case["d1"], 5.0)
    g = orc.rgb_to_intensity(case["rgb1"])
    g0 = orc.rgb_to_intensity(case["img"])
    out.update(in_depth=df, in_grey=g, in_v4=case["v4"], in_n4=case["n4"], in_T0=case["T0"])
    out["pyr_f"] = orc.pyr_down_f(df, ref)
    out["pyr_u8"] = orc.pyr_down_u8(g, ref)
    out["dx"], out["dy"] = orc.derivative_images(g, ref)
    out["vmap"] = orc.create_vmap(df, K, 20.0, ref)
    out["nmap"] = orc.create_nmap(out["vmap"], ref)
    out["copy_v"], out["copy_n"] = orc.copy_maps(case["v4"], case["n4"], ref)
    out["resize_v"] = orc.resize_map(out["copy_v"], False, ref)
    out["resize_n"] = orc.resize_map(out["copy_n"], True, ref)
    R, t = case["T0"][:3, :3], case["T0"][:3, 3]
    out["tr_v"], out["tr_n"] = orc.transform_maps(out["copy_v"], out["copy_n"], R, t, ref)
    out["v2d"] = orc.vertices_to_depth(case["v4"], 6.0, ref)
    out["cloud"] = orc.project_cloud(out["v2d"], K, ref)
    # steps (level-0 sized data, perturbed pose)
    T0 = case["T0"].astype(np.float64)
    T = T0 @ np.array(scenes.synth.make_pose(scenes.synth.rot_y(0.004) @ scenes.synth.rot_x(-0.002), [0.003, -0.002, 0.004]))
    Rpi = np.linalg.inv(T0[:3, :3]).astype(np.float32)
    out.update(in_T=T.astype(np.float32), in_Rpi=Rpi)
    A, b, r, _ = orc.icp_step(T[:3, :3], T[:3, 3], out["vmap"], out["nmap"], Rpi, T0[:3, 3], K, out["tr_v"],
                              out["tr_n"], 0.10, ANGLE, lib=ref)
    out.update(icp_A=A, icp_b=b, icp_res=r)
    krk, kt = scenes.warp_for(K, np.linalg.inv(T) @ T0)
    out.update(in_krk=krk, in_kt=kt, in_grey0=g0)
    depth_m = out["v2d"]
    c, s, n = orc.rgb_residual(64.0, out["dx"], out["dy"], depth_m, depth_m, g0, g, 0.07, kt, krk, lib=ref)
    out.update(res_corres=c, res_sigma=np.int64(s), res_count=np.int64(n))
    A, b = orc.rgb_step(c, float(n), out["cloud"], K, out["dx"], out["dy"], 0.125, lib=ref)
    out.update(rgb_A=A, rgb_b=b)
    fx, fy, cx, cy = [float(k) for k in K]
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
    Rr = scenes.synth.rot_y(0.01) @ scenes.synth.rot_x(0.004)
    Hm, Kinv, KR = Km @ Rr @ np.linalg.inv(Km), np.linalg.inv(Km), Km @ Rr
    A, b, r = orc.so3_step(g0, g, Hm, Kinv, KR, lib=ref)
    out.update(in_so3_H=Hm.astype(np.float32), in_so3_Kinv=Kinv.astype(np.float32), in_so3_KR=KR.astype(np.float32),
               so3_A=A, so3_b=b, so3_res=r)
    # full tracker through the reference kernels (inputs regenerated from the seed by the tests)
    case2 = scenes.room_pair(160, 120)
    oo, _ = scenes.oracle_odometry(case2)
    pose, st, _, extra = oo.track(case2["T0"], use_ref=True)
    out.update(track_pose=pose, track_T0=case2["T0"], track_icp_count=np.float32(st.lastICPCount),
               track_rgb_count=np.float32(st.lastRGBCount), track_in_d1_sum=np.float64(case2["d1"].astype(np.float64).sum()))
    path = os.path.join(HERE, "tracker_ref_sm100a.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
