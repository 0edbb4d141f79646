"""Synthetic RGB-D sequences for the BASELINE.json configs.

There is no dataset in the image and no network, so every test / bench input is ray-cast here
with numpy (analytic plane / box room / moving boxes), seeded and deterministic.

Conventions: K = (fx, fy, cx, cy); poses are 4x4 camera->world (float64 here, cast by callers);
depth is f32 metres (0 = invalid); rgb is u8 HxWx3.
"""
import numpy as np

K_DEFAULT = (528.0, 528.0, 320.0, 240.0)  # GUI/MainController.cpp:109-110


def albedo(X, Y):
    """Grey procedural texture, non-zero everywhere (required by reduce.cu:808, :842).

    SURVEY.md 8(d) proposed only the two smooth sinusoids; at 2 m they give < 3 grey levels per
    pixel, below the tracker's gradient gate at pyramid level 0 ((5/sobelScale)^2 = 1600 on the 3x3
    response, RGBDOdometry.cpp:103-105,:365), so no photometric row survives and a fronto-parallel
    plane is rank deficient for ICP alone (hard part 9).  A 12.5 cm checker
    (step edges of 80 grey levels) is added so the RGB term is active on every level, like a real
    textured scene."""
    g = 128.0 + 40.0 * np.sin(8.0 * X) * np.cos(6.0 * Y) + 20.0 * np.sin(23.0 * X + 5.0 * Y)
    g = g + 40.0 * np.sign(np.sin(25.0 * X)) * np.sign(np.sin(25.0 * Y))
    return np.clip(g, 1.0, 255.0)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def make_pose(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def _rays(W, H, K):
    fx, fy, cx, cy = K
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    return np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)  # HxWx3, z = 1


def _grey_to_rgb(g):
    g8 = g.astype(np.uint8)
    return np.stack([g8, g8, g8], axis=-1)


def render_plane(T_wc, W=640, H=480, K=K_DEFAULT, n=(0.2, -0.1, -1.0), p0=(0.0, 0.0, 2.0)):
    """Config 1: one textured plane with normal n through p0 (world frame)."""
    n = np.asarray(n, dtype=np.float64)
    n = n / np.linalg.norm(n)
    p0 = np.asarray(p0, dtype=np.float64)
    d_c = _rays(W, H, K)
    R, t = T_wc[:3, :3], T_wc[:3, 3]
    d_w = d_c @ R.T
    lam = (n @ (p0 - t)) / (d_w @ n)
    P = t + lam[..., None] * d_w
    depth = np.where(lam > 0, lam, 0.0).astype(np.float32)
    rgb = _grey_to_rgb(albedo(P[..., 0], P[..., 1]))
    return rgb, depth


def render_room(T_wc, W=640, H=480, K=K_DEFAULT, half=(2.0, 1.5, 1.5), boxes=(), max_depth=5.0,
                noise_rng=None):
    """Configs 2-4: inside of a 4x3x3 m box; optional moving boxes (centre, half-size, R_wo).

    Returns rgb, depth and the per-pixel object id (0 = room, i+1 = boxes[i]).
    """
    half = np.asarray(half, dtype=np.float64)
    d_c = _rays(W, H, K)
    R, t = T_wc[:3, :3], T_wc[:3, 3]
    d_w = d_c @ R.T
    # exit point of the ray from the room (camera is inside)
    with np.errstate(divide="ignore", invalid="ignore"):
        l1 = (half - t) / d_w
        l2 = (-half - t) / d_w
    lam_axis = np.where(d_w > 0, l1, l2)
    lam_axis = np.where(np.isfinite(lam_axis) & (lam_axis > 0), lam_axis, np.inf)
    axis = np.argmin(lam_axis, axis=-1)
    lam = np.min(lam_axis, axis=-1)
    P = t + lam[..., None] * d_w
    # wall-local 2-D coordinates: the two non-hit axes (+ per-wall offset so walls differ)
    a = np.where(axis == 0, P[..., 1], P[..., 0])
    b = np.where(axis == 2, P[..., 1], P[..., 2])
    grey = albedo(a + 0.37 * axis, b - 0.21 * axis)
    ids = np.zeros((H, W), dtype=np.uint8)
    for bi, (c, hs, R_wo) in enumerate(boxes):
        c = np.asarray(c, dtype=np.float64)
        hs = np.asarray(hs, dtype=np.float64)
        R_wo = np.asarray(R_wo, dtype=np.float64)
        o_l = (t - c) @ R_wo  # R_wo^T (t - c)
        d_l = d_w @ R_wo
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-hs - o_l) / d_l
            t2 = (hs - o_l) / d_l
        tn = np.nanmax(np.minimum(t1, t2), axis=-1)
        tf = np.nanmin(np.maximum(t1, t2), axis=-1)
        hit = (tn <= tf) & (tn > 0) & (tn < lam)
        Pl = o_l + tn[..., None] * d_l
        face = np.argmax(np.abs(Pl) / hs, axis=-1)
        a2 = np.where(face == 0, Pl[..., 1], Pl[..., 0])
        b2 = np.where(face == 2, Pl[..., 1], Pl[..., 2])
        g2 = albedo(3.0 * a2 + 1.3 * (bi + 1), 3.0 * b2 - 0.7 * (bi + 1))
        grey = np.where(hit, g2, grey)
        lam = np.where(hit, tn, lam)
        ids = np.where(hit, np.uint8(bi + 1), ids)
    depth = lam.astype(np.float64)
    if noise_rng is not None:
        depth = depth + noise_rng.standard_normal(depth.shape) * (0.0012 * depth * depth)
    depth = np.where((depth > 0) & (depth <= max_depth), depth, 0.0).astype(np.float32)
    return _grey_to_rgb(grey), depth, ids


def plane_sequence(n_frames=64, W=640, H=480, K=K_DEFAULT):
    """Config 1: +2 mm/frame along x, +0.1 deg/frame yaw; timestamps i*33."""
    for i in range(n_frames):
        T = make_pose(rot_y(np.deg2rad(0.1 * i)), np.array([0.002 * i, 0.0, 0.0]))
        rgb, depth = render_plane(T, W, H, K)
        yield i * 33, rgb, depth, T


def room_sequence(n_frames=200, W=640, H=480, K=K_DEFAULT, noise=True, seed=1234, n_boxes=0,
                  deg_per_frame=1.0, radius=0.5, box_speed=1.0, box_start=0):
    """Config 2 (n_boxes=0) / 3 (n_boxes=4) / 4 (n_boxes=8): camera on a 0.5 m orbit, 1 deg/frame."""
    rng = np.random.default_rng(seed) if noise else None
    brng = np.random.default_rng(seed + 1)
    specs = []
    # boxes start inside the first camera's field of view (1.2-1.9 m in front of it) and drift on
    # constant-velocity trajectories of 5-20 mm/frame
    R0 = rot_y(np.pi / 2 + 0.6) @ rot_x(-0.4)
    t0 = np.array([radius, 0.0, 0.0])
    for b in range(n_boxes):
        off = np.array([0.55 * np.cos(2.4 * b + 0.4), 0.32 * np.sin(1.7 * b + 0.9), 1.25 + 0.16 * (b % 4)])
        c0 = t0 + R0 @ off
        hs = brng.uniform(0.10, 0.18, size=3)
        vel = brng.uniform(-1, 1, size=3)
        vel = vel / np.linalg.norm(vel) * brng.uniform(0.005, 0.02)
        w = brng.uniform(-0.01, 0.01)
        specs.append((c0, hs, vel, w))
    for i in range(n_frames):
        a = np.deg2rad(deg_per_frame * i)
        t = np.array([radius * np.cos(a), 0.0, radius * np.sin(a)])
        # look across the room towards a corner, pitched so that floor/ceiling + walls are in view:
        # a single visible wall would make point-to-plane ICP rank deficient
        T = make_pose(rot_y(-a + np.pi / 2 + 0.6) @ rot_x(-0.4 + 0.1 * np.sin(2 * a)), t)
        bi = max(i - box_start, 0)  # the boxes stand still until frame `box_start`
        boxes = [(c0 + vel * box_speed * bi, hs, rot_y(w * bi)) for (c0, hs, vel, w) in specs]
        rgb, depth, ids = render_room(T, W, H, K, boxes=boxes, noise_rng=rng)
        yield i * 33, rgb, depth, T, ids


def prediction_from_depth(depth, rgb, K, conf=10.0):
    """Perfect 'model prediction' in the camera frame from an exact depth image: AoS float4 vertex
    map (x,y,z,conf), normal map (nx,ny,nz,radius) by central differences, and the image.  Used to
    exercise the tracker before / independently of the surfel stage."""
    H, W = depth.shape
    fx, fy, cx, cy = K
    u, v = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    z = depth.astype(np.float32)
    X = (u - np.float32(cx)) * z / np.float32(fx)
    Y = (v - np.float32(cy)) * z / np.float32(fy)
    P = np.stack([X, Y, z], axis=-1)
    dx = np.zeros_like(P)
    dy = np.zeros_like(P)
    dx[:, 1:-1] = P[:, 2:] - P[:, :-2]
    dy[1:-1, :] = P[2:, :] - P[:-2, :]
    n = np.cross(dx, dy)
    nn = np.linalg.norm(n, axis=-1, keepdims=True)
    valid = (z > 0) & (nn[..., 0] > 0)
    valid[:, 0] = valid[:, -1] = False
    valid[0, :] = valid[-1, :] = False
    n = np.where(nn > 0, n / np.maximum(nn, 1e-20), 0)
    rad = z * np.float32(np.sqrt(2.0)) / np.float32((fx + fy) / 2)
    v4 = np.concatenate([P, np.full((H, W, 1), conf, np.float32)], axis=-1).astype(np.float32)
    n4 = np.concatenate([n, rad[..., None]], axis=-1).astype(np.float32)
    v4[~valid] = 0
    n4[~valid] = 0
    img = rgb.copy()
    img[~valid] = 0
    return np.ascontiguousarray(v4), np.ascontiguousarray(n4), np.ascontiguousarray(img)
