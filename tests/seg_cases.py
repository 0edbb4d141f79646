"""Seeded inputs of the segmentation tests (shared by the CPU invariants and the GPU parity test)."""
import numpy as np

from cofusion_b200 import synth


def room_case(W=640, H=480, n_boxes=2, err=0.16, seed=0):
    """One frame of the synthetic room with `n_boxes` boxes; the ICP error of the background model is
    fabricated: `err` metres on box 1, a millimetre elsewhere; background confidence 10 everywhere."""
    fx, fy, cx, cy = synth.K_DEFAULT
    K = (fx * W / 640.0, fy * H / 480.0, cx * W / 640.0, cy * H / 480.0)
    seq = list(synth.room_sequence(2, W, H, K, noise=True, n_boxes=n_boxes, seed=seed))
    _, rgb, depth, _, ids = seq[1]
    rng = np.random.RandomState(seed + 17)
    icp = np.where(ids == 1, err, 0.001).astype(np.float32)
    icp += rng.uniform(0, 2e-4, icp.shape).astype(np.float32)
    vc = np.zeros((H, W, 4), np.float32)
    vc[..., 3] = 10.0
    return dict(rgb=np.ascontiguousarray(rgb), depth=np.ascontiguousarray(depth, np.float32), ids=ids, icp=[icp],
                vc=[vc], model_ids=[0], next_id=1)


def two_model_case(W=640, H=480, seed=1):
    """Background + object model 1 (covering box 1, confident there) + box 2 moving -> new label 2."""
    c = room_case(W, H, n_boxes=2, seed=seed)
    ids = c["ids"]
    rng = np.random.RandomState(seed + 5)
    icp0 = np.where(ids == 1, 0.12, np.where(ids == 2, 0.2, 0.001)).astype(np.float32)
    icp1 = np.where(ids == 1, 0.002, 0.09).astype(np.float32)
    icp1 += rng.uniform(0, 1e-4, icp1.shape).astype(np.float32)
    vc0 = np.zeros((H, W, 4), np.float32)
    vc0[..., 3] = 10.0
    vc1 = np.zeros((H, W, 4), np.float32)
    vc1[..., 3] = np.where(ids == 1, 3.0, 0.0)
    vc1[: H // 12, :, 3] = np.nan  # unset texels of the splat (isfinite fix-up, Segmentation.cpp:200-203)
    c.update(icp=[icp0, icp1], vc=[vc0, vc1], model_ids=[0, 1], next_id=2)
    return c


def noise_case(W=320, H=240, seed=3):
    """Random image, depth with a large hole (empty thresholded super-pixels -> Slic.h:117-122 quirk path)"""
    rng = np.random.RandomState(seed)
    rgb = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    rgb[:, : W // 2] = (rgb[:, : W // 2] // 8) + 100  # a flat half so that SLIC has structure
    depth = rng.uniform(0.5, 4.0, (H, W)).astype(np.float32)
    depth[H // 4: H // 2, W // 8: W // 2] = 0.0
    depth[::3, ::3] = 0.01
    icp = rng.uniform(0, 0.2, (H, W)).astype(np.float32)
    icp[:, W // 2:] *= 0.01
    vc = np.zeros((H, W, 4), np.float32)
    vc[..., 3] = rng.uniform(0, 12, (H, W))
    return dict(rgb=rgb, depth=depth, ids=None, icp=[icp], vc=[vc], model_ids=[0], next_id=1)
