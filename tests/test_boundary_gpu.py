"""-m gpu: the drop-in boundary beyond plain processFrame -- device-side frame ingest (row f1), the inPose /
bootstrap / timestamp arguments of CoFusion::processFrame, pose logging and the PLY / pose exports (row f2), and the
compiled reference-side binding of INTEGRATION.md section 1."""
import os
import struct
import subprocess

import numpy as np
import pytest

import scenes
from cofusion_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _seq(n, W=320, H=240):
    return list(synth.room_sequence(n, W, H, scenes.scaled_K(W), noise=True))


def test_device_ingest_u16_bgr_is_bit_identical_to_host_conversion():
    """GUI/Tools/KlgLogReader.cpp:53-84 converts raw u16 depth with cv::Mat::convertTo(CV_32FC1, 0.001) and
    FrameData::flipColors swaps BGR -> RGB on the CPU; here both happen in one device kernel.  Feeding the raw
    frame must give the same poses and the same map, bit for bit, as feeding the host-converted one."""
    import cofusion_b200 as cfb
    W, H = 320, 240
    K = scenes.scaled_K(W)
    seq = _seq(5, W, H)
    a = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 19))
    b = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 19))
    for t, (_, rgb, d, _, _) in enumerate(seq):
        d16 = np.clip(np.rint(d * 1000.0), 0, 65535).astype(np.uint16)
        d32 = d16.astype(np.float32) * np.float32(0.001)          # what convertTo computes
        bgr = np.ascontiguousarray(rgb[..., ::-1])
        a.process_frame(np.ascontiguousarray(rgb), d32)
        b.process_frame_ex(bgr, depth_u16=d16, depth_scale=0.001, flip_colors=True, timestamp=33 * t)
        assert np.array_equal(a.pose(0), b.pose(0)), t
    assert np.array_equal(a.ctx.view(0), b.ctx.view(0)) and np.array_equal(a.ctx.view(1), b.ctx.view(1))
    ma, mb = a.model(0).download_map(), b.model(0).download_map()
    assert len(ma) == len(mb) and np.array_equal(ma, mb)


def test_in_pose_override_and_bootstrap():
    """processFrame(frame, inPose, weight, bootstrap) (CoFusion.cpp:210-222, :343-345)"""
    import cofusion_b200 as cfb
    W, H = 320, 240
    K = scenes.scaled_K(W)
    seq = _seq(4, W, H)
    T0i = np.linalg.inv(seq[0][3])
    gt = [(T0i @ s[3]).astype(np.float32) for s in seq]
    # (a) pose supplied by the caller: nothing is tracked, the map is fused at that pose
    a = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 19))
    a.process_frame(np.ascontiguousarray(seq[0][1]), np.ascontiguousarray(seq[0][2]))
    n0 = a.model(0).last_count()
    for t in (1, 2):
        a.process_frame_ex(np.ascontiguousarray(seq[t][1]), depth=np.ascontiguousarray(seq[t][2]), in_pose=gt[t])
        assert np.array_equal(a.pose(0), gt[t])
    assert a.model(0).last_count() > n0  # the fuse / clean block still ran
    # (b) bootstrap: track, then pose <- pose * inPose
    b = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 19))
    c = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 19))
    for t in range(2):
        for f in (b, c):
            f.process_frame(np.ascontiguousarray(seq[t][1]), np.ascontiguousarray(seq[t][2]))
    D = np.eye(4, dtype=np.float32)
    D[0, 3] = 0.001
    b.process_frame_ex(np.ascontiguousarray(seq[2][1]), depth=np.ascontiguousarray(seq[2][2]), in_pose=D, bootstrap=True)
    c.process_frame(np.ascontiguousarray(seq[2][1]), np.ascontiguousarray(seq[2][2]))
    # c's pose is the tracked pose of frame 2; b's is that pose times D (float32 product in the reference's order)
    assert np.abs(b.pose(0) - (c.pose(0) @ D)).max() < 1e-6


def test_pose_log_and_exports(tmp_path):
    """Model pose log (CoFusion.cpp:503-518), exportPoses (:758-783) and savePly (:646-756)"""
    import cofusion_b200 as cfb
    W, H = 320, 240
    K = scenes.scaled_K(W)
    seq = _seq(5, W, H)
    f = cfb.CoFusion(W, H, K, cfb.CoFusionParams.default(1 << 19))
    f.enable_pose_logging()
    poses = []
    for t, (ts, rgb, d, _, _) in enumerate(seq):
        f.process_frame_ex(np.ascontiguousarray(rgb), depth=np.ascontiguousarray(d), timestamp=ts)
        poses.append(f.pose(0).copy())
    ts, p7 = f.pose_log(0)
    assert ts.tolist() == [s[0] for s in seq] and p7.shape == (5, 7)
    for k in range(5):  # t.xyz, then a unit quaternion x y z w that reproduces the rotation
        assert np.array_equal(p7[k, :3], poses[k][:3, 3])
        x, y, z, w = p7[k, 3:].astype(np.float64)
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert abs(x * x + y * y + z * z + w * w - 1) < 1e-5 and np.abs(R - poses[k][:3, :3]).max() < 1e-5
    f.export_poses(tmp_path)
    f.save_ply(tmp_path)
    lines = open(tmp_path / "poses-0.txt").read().strip().split("\n")
    assert len(lines) == 5 and lines[3].split()[0] == str(seq[3][0]) and len(lines[3].split()) == 8
    assert np.allclose([float(v) for v in lines[4].split()[1:]], p7[4], atol=1e-6)
    raw = open(tmp_path / "cloud-0.ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    nv = int([l for l in head.decode().split("\n") if l.startswith("element vertex")][0].split()[2])
    m = f.model(0).download_map()
    conf_thr = f.model(0).info()[1]
    keep = m[:, 3] > conf_thr
    assert nv == int(keep.sum()) and len(body) == nv * 31  # 3 f32 + 3 u8 + 3 f32 + 1 f32
    if nv:
        x, y, z, r, g, b, nx, ny, nz, rad = struct.unpack("<fffBBBffff", body[:31])
        s = m[keep][0]
        assert np.allclose([x, y, z], s[:3], atol=1e-6) and rad == s[11]  # camera model: Tp = identity
        assert (r << 16 | g << 8 | b) == int(s[4]) and np.allclose([nx, ny, nz], -s[8:11], atol=1e-6)


def test_compiled_reference_side_binding(tmp_path):
    """tests/binding/cofusion_binding.cpp against the real library on cuda:0"""
    import cofusion_b200  # noqa: F401  (builds / locates the library)
    libdir = os.path.join(ROOT, "cofusion_b200")
    exe = tmp_path / "binding_test"
    subprocess.run(["g++", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "binding", "cofusion_binding.cpp"), "-L", libdir, "-lcofusion_b200",
                    "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    out = tmp_path / "export"
    out.mkdir()
    r = subprocess.run([str(exe), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "binding: ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
    assert (out / "cloud-0.ply").exists() and (out / "poses-0.txt").exists()
