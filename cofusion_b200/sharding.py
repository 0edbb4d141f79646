"""Host-side logic of the object-sharded deployment shared by bench.py, the tools and
the world-size-2 gloo test: which rank owns which model, how a frame is packed for the ONE broadcast per time
step, and how the per-rank timings combine into the whole-job figure.  The device side lives in
csrc/shard.cu (NCCL) and CoFusion::processFrameEx."""
import numpy as np


def owner(model_index, world):
    """rank that owns the model at list position `model_index` (0 = camera / background model); shard.cuh"""
    return model_index % world if world > 0 else 0


def models_of_rank(n_models, rank, world):
    return [m for m in range(n_models) if owner(m, world) == rank]


def scene_models(world):
    """weak scaling: a world of N ranks tracks ONE scene with N models (background + N - 1 objects), so that
    N = 1 is BASELINE.json configs[1] and N = 8 is configs[3] (8 models, one per GPU)"""
    return max(1, world)


def packed_bytes(W, H):
    return 8 * W * H


def pack_frame(rgb, depth, mask=None):
    """[rgb u8 3P | depth f32 4P | mask u8 P] -- the layout CoFusion::processFrameEx broadcasts"""
    H, W = depth.shape
    P = W * H
    buf = np.zeros(8 * P, np.uint8)
    buf[:3 * P] = np.ascontiguousarray(rgb, np.uint8).reshape(-1)
    buf[3 * P:7 * P] = np.ascontiguousarray(depth, np.float32).reshape(-1).view(np.uint8)
    if mask is not None:
        buf[7 * P:] = np.ascontiguousarray(mask, np.uint8).reshape(-1)
    return buf


def unpack_frame(buf, W, H):
    P = W * H
    buf = np.asarray(buf, np.uint8)
    return (buf[:3 * P].reshape(H, W, 3), buf[3 * P:7 * P].view(np.float32).reshape(H, W), buf[7 * P:8 * P].reshape(H, W))


def aggregate_value(n_models, steps, max_rank_ms):
    """whole-job throughput in per-model frames/s: every model advanced `steps` frames in the time of the slowest rank"""
    return n_models * steps / (max_rank_ms / 1e3)
