"""cofusion_b200 -- Python harness around libcofusion_b200.so (the C ABI of include/cofusion_b200.h).

The product is the CUDA library; this package only loads it, declares the prototypes and gives the
tests / bench a thin object layer.  PyTorch is used by callers for device memory and streams
(`tensor.data_ptr()` is what crosses the ABI).  There is NO CPU fallback: if the extension is not
built, or no CUDA device is visible, the constructors raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcofusion_b200.so")
_lib = None

c_float_p = C.POINTER(C.c_float)
c_void_pp = C.POINTER(C.c_void_p)


class TrackStats(C.Structure):
    _fields_ = [("lastICPError", C.c_float), ("lastICPCount", C.c_float), ("lastRGBError", C.c_float),
                ("lastRGBCount", C.c_float), ("lastSO3Error", C.c_float), ("lastSO3Count", C.c_float),
                ("lastA", C.c_double * 36), ("lastb", C.c_double * 6), ("so3_iterations", C.c_int),
                ("pad_", C.c_int)]


class CfbError(RuntimeError):
    pass


def lib():
    """Load libcofusion_b200.so (raises if it has not been built -- no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CfbError("libcofusion_b200.so is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.cfb_last_error.restype = C.c_char_p
        _lib.cfb_step_scratch_bytes.restype = C.c_size_t
    return _lib


def check(rc):
    if rc != 0:
        raise CfbError("cfb error %d: %s" % (rc, lib().cfb_last_error().decode()))


def _p(t):
    """device pointer of a torch tensor (or None)"""
    return C.c_void_p(0 if t is None else t.data_ptr())


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.ctypes.data_as(c_float_p), a


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Odometry:
    """cfb_odom_* = RGBDOdometry (Core/Utils/RGBDOdometry.h). Inputs are torch CUDA tensors."""

    def __init__(self, W, H, K, dist_thresh=0.10, angle_thresh=float(np.sin(np.deg2rad(20.0)))):
        fx, fy, cx, cy = K
        self.W, self.H, self.K = W, H, K
        self._h = C.c_void_p()
        check(lib().cfb_odom_create(W, H, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy),
                                    C.c_float(dist_thresh), C.c_float(angle_thresh), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.cfb_odom_destroy(self._h)
            self._h = C.c_void_p()

    def set_mode(self, mode):
        check(lib().cfb_odom_set_mode(self._h, int(mode)))

    def init_icp(self, depth_pyr, cutoff):
        ptrs = (C.c_void_p * 3)(*[d.data_ptr() for d in depth_pyr])
        pitch = (C.c_size_t * 3)(*[d.stride(0) * 4 for d in depth_pyr])
        check(lib().cfb_odom_init_icp(self._h, ptrs, pitch, C.c_float(cutoff), _stream()))

    def init_icp_model(self, v4, n4, cutoff, pose):
        pp, keep = _f(pose)
        check(lib().cfb_odom_init_icp_model(self._h, _p(v4), _p(n4), C.c_float(cutoff), pp, _stream()))

    def init_rgb_model(self, img):
        check(lib().cfb_odom_init_rgb_model(self._h, _p(img), C.c_size_t(img.stride(0)), img.shape[2], _stream()))

    def init_rgb(self, img):
        check(lib().cfb_odom_init_rgb(self._h, _p(img), C.c_size_t(img.stride(0)), img.shape[2], _stream()))

    def init_first_rgb(self, img):
        check(lib().cfb_odom_init_first_rgb(self._h, _p(img), C.c_size_t(img.stride(0)), img.shape[2], _stream()))

    def track(self, pose, rgb_only=False, icp_weight=10.0, pyramid=True, fast_odom=False, so3=True,
              error_map=None, force_host_loop=False):
        """getIncrementalTransformation; pose: 4x4 camera->world (numpy) -> new 4x4 + stats."""
        pose = np.asarray(pose, dtype=np.float32)
        trans = np.ascontiguousarray(pose[:3, 3]).copy()
        rot = np.ascontiguousarray(pose[:3, :3]).copy()
        st = TrackStats()
        check(lib().cfb_odom_get_incremental_transformation(
            self._h, trans.ctypes.data_as(c_float_p), rot.ctypes.data_as(c_float_p), int(rgb_only),
            C.c_float(icp_weight), int(pyramid), int(fast_odom), int(so3), _p(error_map),
            C.c_size_t(0 if error_map is None else error_map.stride(0) * 4), int(force_host_loop), C.byref(st),
            _stream()))
        out = np.eye(4, dtype=np.float32)
        out[:3, :3] = rot
        out[:3, 3] = trans
        return out, st

    def view(self, which, level):
        """copy of an internal pyramid buffer as a numpy array"""
        import torch
        ptr, pitch = C.c_void_p(), C.c_size_t()
        check(lib().cfb_odom_view(self._h, which, level, C.byref(ptr), C.byref(pitch)))
        w, h = self.W >> level, self.H >> level
        shapes = {0: (3 * h, w, np.float32), 1: (3 * h, w, np.float32), 2: (3 * h, w, np.float32),
                  3: (3 * h, w, np.float32), 4: (h, w, np.float32), 5: (h, w, np.float32), 6: (h, w, np.uint8),
                  7: (h, w, np.uint8), 8: (h, w, np.int16), 9: (h, w, np.int16), 10: (h, w, np.uint8),
                  11: (h, w * 3, np.float32), 12: (h, w * 4, np.int32)}
        r, c, dt = shapes[which]
        out = np.empty((r, c), dtype=dt)
        torch.cuda.synchronize()
        _cudart_memcpy_d2h(out, ptr.value)
        return out


def _cudart_memcpy_d2h(dst_np, src_ptr):
    check(lib().cfb_download(dst_np.ctypes.data_as(C.c_void_p), C.c_void_p(src_ptr), C.c_size_t(dst_np.nbytes),
                             _stream()))
