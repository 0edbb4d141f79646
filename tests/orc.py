"""ctypes bindings of the oracle (oracle/_build/liboracle.so) and, when present, of the compiled
reference CUDA sources (oracle/_ref/libcfref.so).  TEST INFRASTRUCTURE -- see oracle/cf_oracle.h."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
ORC_LIB = os.path.join(ORC_DIR, "_build", "liboracle.so")
REF_LIB = os.path.join(ORC_DIR, "_ref", "libcfref.so")

fp = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)


class OrcTrackStats(C.Structure):
    _fields_ = [("lastICPError", C.c_float), ("lastICPCount", C.c_float), ("lastRGBError", C.c_float),
                ("lastRGBCount", C.c_float), ("lastSO3Error", C.c_float), ("lastSO3Count", C.c_float),
                ("lastA", C.c_double * 36), ("lastb", C.c_double * 6), ("so3_iterations", C.c_int)]


_orc = None
_ref = None


def build_oracle():
    srcs = [os.path.join(ORC_DIR, f) for f in os.listdir(ORC_DIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(ORC_LIB) or any(os.path.getmtime(s) > os.path.getmtime(ORC_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORC_DIR, "-s"])
    return ORC_LIB


def orc():
    global _orc
    if _orc is None:
        _orc = C.CDLL(build_oracle())
        _orc.orc_odom_create.restype = C.c_void_p
        _orc.orc_odom_view.restype = C.c_void_p
    return _orc


def ref():
    """compiled reference kernels, or None when oracle/_ref was not built"""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        orc()
        _ref = C.CDLL(REF_LIB)
    return _ref


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def P(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def cf(x):
    return C.c_float(float(x))


# ---------------------------------------------------------------- image preparation
def bilateral(depth, maxD):
    H, W = depth.shape
    out = np.empty((H, W), np.float32)
    orc().orc_bilateral_filter(P(f32(depth)), W, H, cf(maxD), P(out))
    return out


def pyr_down_f(src, lib=None):
    H, W = src.shape
    out = np.empty((H // 2, W // 2), np.float32)
    (lib or orc()).__getattr__("ref_pyr_down_gauss_f" if lib else "orc_pyr_down_gauss_f")(P(f32(src)), W, H, P(out))
    return out


def pyr_down_u8(src, lib=None):
    H, W = src.shape
    out = np.empty((H // 2, W // 2), np.uint8)
    src = np.ascontiguousarray(src, np.uint8)
    (lib or orc()).__getattr__("ref_pyr_down_uchar_gauss" if lib else "orc_pyr_down_uchar_gauss")(P(src), W, H, P(out))
    return out


def create_vmap(depth, K, cutoff, lib=None):
    H, W = depth.shape
    out = np.empty((3 * H, W), np.float32)
    fx, fy, cx, cy = K
    (lib or orc()).__getattr__("ref_create_vmap" if lib else "orc_create_vmap")(
        P(f32(depth)), W, H, cf(fx), cf(fy), cf(cx), cf(cy), cf(cutoff), P(out))
    return out


def create_nmap(vmap, lib=None):
    H3, W = vmap.shape
    out = np.empty((H3, W), np.float32)
    (lib or orc()).__getattr__("ref_create_nmap" if lib else "orc_create_nmap")(P(f32(vmap)), W, H3 // 3, P(out))
    return out


def copy_maps(v4, n4, lib=None):
    H, W = v4.shape[:2]
    v = np.empty((3 * H, W), np.float32)
    n = np.empty((3 * H, W), np.float32)
    (lib or orc()).__getattr__("ref_copy_maps" if lib else "orc_copy_maps")(P(f32(v4)), P(f32(n4)), W, H, P(v), P(n))
    return v, n


def resize_map(m, normalize, lib=None):
    H3, W = m.shape
    H = H3 // 3
    out = np.empty((3 * (H // 2), W // 2), np.float32)
    (lib or orc()).__getattr__("ref_resize_map" if lib else "orc_resize_map")(P(f32(m)), W, H, int(normalize), P(out))
    return out


def transform_maps(v, n, R, t, lib=None):
    H3, W = v.shape
    vd = np.empty_like(v)
    nd = np.empty_like(n)
    R = f32(R).reshape(9)
    t = f32(t).reshape(3)
    (lib or orc()).__getattr__("ref_transform_maps" if lib else "orc_transform_maps")(
        P(f32(v)), P(f32(n)), W, H3 // 3, P(R), P(t), P(vd), P(nd))
    return vd, nd


def vertices_to_depth(v4, cutoff, lib=None):
    H, W = v4.shape[:2]
    out = np.empty((H, W), np.float32)
    (lib or orc()).__getattr__("ref_vertices_to_depth" if lib else "orc_vertices_to_depth")(P(f32(v4)), W, H, cf(cutoff), P(out))
    return out


def rgb_to_intensity(img):
    H, W, ch = img.shape
    out = np.empty((H, W), np.uint8)
    orc().orc_rgb_to_intensity(P(np.ascontiguousarray(img, np.uint8)), ch, W, H, P(out))
    return out


def derivative_images(img, lib=None):
    H, W = img.shape
    dx = np.empty((H, W), np.int16)
    dy = np.empty((H, W), np.int16)
    (lib or orc()).__getattr__("ref_derivative_images" if lib else "orc_derivative_images")(
        P(np.ascontiguousarray(img, np.uint8)), W, H, P(dx), P(dy))
    return dx, dy


def project_cloud(depth, K, lib=None):
    H, W = depth.shape
    out = np.empty((H, W * 3), np.float32)
    fx, fy, cx, cy = K
    (lib or orc()).__getattr__("ref_project_to_point_cloud" if lib else "orc_project_to_point_cloud")(
        P(f32(depth)), W, H, cf(fx), cf(fy), cf(cx), cf(cy), P(out))
    return out


# ---------------------------------------------------------------- reduction steps
def icp_step(Rcurr, tcurr, vc, nc, Rprev_inv, tprev, K, vp, np_, dist, angle, want_error=False, lib=None):
    H3, W = vc.shape
    H = H3 // 3
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    res = np.zeros(2, np.float32)
    fx, fy, cx, cy = K
    args = [P(f32(Rcurr).reshape(9)), P(f32(tcurr)), P(f32(vc)), P(f32(nc)), P(f32(Rprev_inv).reshape(9)),
            P(f32(tprev)), cf(fx), cf(fy), cf(cx), cf(cy), P(f32(vp)), P(f32(np_)), cf(dist), cf(angle), W, H,
            P(A), P(b), P(res)]
    if lib:
        lib.ref_icp_step(*args)
        return A.reshape(6, 6), b, res, None
    err = np.zeros((H, W), np.float32) if want_error else None
    orc().orc_icp_step(*args, P(err))
    return A.reshape(6, 6), b, res, err


def rgb_residual(minScale, dIdx, dIdy, lastDepth, nextDepth, lastImage, nextImage, maxDelta, kt, krkinv, lib=None):
    H, W = lastDepth.shape
    corres = np.zeros((H, W, 4), np.int32)
    sigma = C.c_int(0)
    count = C.c_int(0)
    (lib or orc()).__getattr__("ref_rgb_residual" if lib else "orc_rgb_residual")(
        cf(minScale), P(np.ascontiguousarray(dIdx, np.int16)), P(np.ascontiguousarray(dIdy, np.int16)),
        P(f32(lastDepth)), P(f32(nextDepth)), P(np.ascontiguousarray(lastImage, np.uint8)),
        P(np.ascontiguousarray(nextImage, np.uint8)), P(corres), cf(maxDelta), P(f32(kt)), P(f32(krkinv).reshape(9)),
        W, H, C.byref(sigma), C.byref(count))
    return corres, sigma.value, count.value


def rgb_step(corres, sigma, cloud, K, dIdx, dIdy, sobelScale, lib=None):
    H, W = dIdx.shape
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    fx, fy = K[0], K[1]
    (lib or orc()).__getattr__("ref_rgb_step" if lib else "orc_rgb_step")(
        P(np.ascontiguousarray(corres, np.int32)), cf(sigma), P(f32(cloud)), cf(fx), cf(fy),
        P(np.ascontiguousarray(dIdx, np.int16)), P(np.ascontiguousarray(dIdy, np.int16)), cf(sobelScale), W, H,
        P(A), P(b))
    return A.reshape(6, 6), b


def so3_step(lastImage, nextImage, imageBasis, kinv, krlr, lib=None):
    H, W = lastImage.shape
    A = np.zeros(9, np.float32)
    b = np.zeros(3, np.float32)
    res = np.zeros(2, np.float32)
    (lib or orc()).__getattr__("ref_so3_step" if lib else "orc_so3_step")(
        P(np.ascontiguousarray(lastImage, np.uint8)), P(np.ascontiguousarray(nextImage, np.uint8)),
        P(f32(imageBasis).reshape(9)), P(f32(kinv).reshape(9)), P(f32(krlr).reshape(9)), W, H, P(A), P(b), P(res))
    return A.reshape(3, 3), b, res


def corres_valid(corres):
    """valid flag / fields of a raw DataTerm image (H,W,4) int32"""
    c = np.ascontiguousarray(corres, np.int32)
    valid = (c[..., 3] & 0xff) != 0
    zero_x = (c[..., 0] & 0xffff).astype(np.int16)
    zero_y = ((c[..., 0] >> 16) & 0xffff).astype(np.int16)
    diff = c[..., 2].view(np.float32)
    return valid, zero_x, zero_y, diff


# ---------------------------------------------------------------- RGBDOdometry restatement
class OrcOdometry:
    VIEW_SHAPE = {0: (3, 4, np.float32), 1: (3, 4, np.float32), 2: (3, 4, np.float32), 3: (3, 4, np.float32),
                  4: (1, 4, np.float32), 5: (1, 4, np.float32), 6: (1, 1, np.uint8), 7: (1, 1, np.uint8),
                  8: (1, 2, np.int16), 9: (1, 2, np.int16), 10: (1, 1, np.uint8)}

    def __init__(self, W, H, K, dist=0.10, angle=float(np.sin(np.deg2rad(20.0)))):
        fx, fy, cx, cy = K
        self.W, self.H, self.K = W, H, K
        self.dist, self.angle = dist, angle
        self.h = C.c_void_p(orc().orc_odom_create(W, H, cf(cx), cf(cy), cf(fx), cf(fy), cf(dist), cf(angle)))

    def __del__(self):
        if getattr(self, "h", None) and _orc is not None:
            _orc.orc_odom_destroy(self.h)
            self.h = None

    def init_model(self, v4, n4, img, pose):
        orc().orc_odom_init_model(self.h, P(f32(v4)), P(f32(n4)), P(np.ascontiguousarray(img, np.uint8)),
                                  img.shape[2], P(f32(pose).reshape(16)))

    def init_frame(self, depth_filtered, rgb, cutoff):
        orc().orc_odom_init_frame(self.h, P(f32(depth_filtered)), P(np.ascontiguousarray(rgb, np.uint8)),
                                  rgb.shape[2], cf(cutoff))

    def init_first_rgb(self, rgb):
        orc().orc_odom_init_first_rgb(self.h, P(np.ascontiguousarray(rgb, np.uint8)), rgb.shape[2])

    def track(self, pose, rgb_only=False, icp_weight=10.0, pyramid=True, fast_odom=False, so3=True,
              want_error=False, use_ref=False):
        pose = np.asarray(pose, np.float32)
        trans = np.ascontiguousarray(pose[:3, 3]).copy()
        rot = np.ascontiguousarray(pose[:3, :3]).copy()
        st = OrcTrackStats()
        err = np.zeros((self.H, self.W), np.float32) if want_error else None
        extra = {}
        if use_ref:
            ms = C.c_double(0)
            steps = C.c_int(0)
            ref().ref_odom_track(self.h, P(trans), P(rot), int(rgb_only), cf(icp_weight), int(pyramid),
                                 int(fast_odom), int(so3), cf(self.dist), cf(self.angle), C.byref(st),
                                 C.byref(ms), C.byref(steps))
            extra = {"step_ms": ms.value, "steps": steps.value}
        else:
            orc().orc_odom_track(self.h, P(trans), P(rot), int(rgb_only), cf(icp_weight), int(pyramid),
                                 int(fast_odom), int(so3), P(err), C.byref(st))
        out = np.eye(4, dtype=np.float32)
        out[:3, :3] = rot
        out[:3, 3] = trans
        return out, st, err, extra

    def view(self, which, level):
        planes, esz, dt = self.VIEW_SHAPE[which]
        w, h = self.W >> level, self.H >> level
        ptr = orc().orc_odom_view(self.h, which, level)
        n = planes * h * w
        buf = (C.c_char * (n * esz)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(planes * h, w).copy()


# ---------------------------------------------------------------- surfel map restatement
class OrcMap:
    VIEWS = {0: (1, np.uint32), 1: (4, np.float32), 2: (4, np.float32), 3: (4, np.float32), 4: (4, np.uint8),
             5: (4, np.float32), 6: (4, np.float32), 7: (1, np.uint16), 8: (4, np.uint8), 9: (4, np.float32),
             10: (4, np.float32)}

    def __init__(self, W, H, K, max_surfels=1 << 20):
        fx, fy, cx, cy = K
        self.W, self.H, self.K = W, H, K
        o = orc()
        o.orc_map_create.restype = C.c_void_p
        o.orc_map_surfels.restype = C.c_void_p
        o.orc_map_unstable.restype = C.c_void_p
        o.orc_map_view.restype = C.c_void_p
        o.orc_fusion_weight.restype = C.c_float
        self.h = C.c_void_p(o.orc_map_create(W, H, cf(fx), cf(fy), cf(cx), cf(cy), max_surfels))

    def __del__(self):
        if getattr(self, "h", None) and _orc is not None:
            _orc.orc_map_destroy(self.h)
            self.h = None

    @staticmethod
    def _u8(a):
        return P(np.ascontiguousarray(a, np.uint8))

    def initialise(self, rgb, depth_raw, depth_filtered, time, max_depth=20.0):
        orc().orc_map_initialise(self.h, self._u8(rgb), P(f32(depth_raw)), P(f32(depth_filtered)), int(time), cf(max_depth))

    def predict_indices(self, pose, time, max_depth=20.0, time_delta=200):
        orc().orc_map_predict_indices(self.h, P(f32(pose).reshape(16)), int(time), cf(max_depth), int(time_delta))

    def fuse(self, pose, time, rgb, mask, depth_raw, depth_filtered, max_depth, weighting, mask_id=0):
        orc().orc_map_fuse(self.h, P(f32(pose).reshape(16)), int(time), self._u8(rgb), self._u8(mask), P(f32(depth_raw)),
                           P(f32(depth_filtered)), cf(max_depth), cf(weighting), int(mask_id))

    def clean(self, pose, time, conf_threshold, time_delta, depth_filtered, mask, mask_id=0, outlier_coeff=3.0):
        orc().orc_map_clean(self.h, P(f32(pose).reshape(16)), int(time), cf(conf_threshold), int(time_delta),
                            P(f32(depth_filtered)), self._u8(mask), int(mask_id), cf(outlier_coeff))

    def combined_predict(self, pose, max_depth, conf_threshold, time, max_time, time_delta=200):
        orc().orc_map_combined_predict(self.h, P(f32(pose).reshape(16)), cf(max_depth), cf(conf_threshold), int(time),
                                       int(max_time), int(time_delta))

    def fill_in(self, rgb, depth_filtered, passthrough_geom=False, passthrough_rgb=False):
        orc().orc_map_fill_in(self.h, self._u8(rgb), P(f32(depth_filtered)), int(passthrough_geom), int(passthrough_rgb))

    def requires_fill_in(self, ratio=0.75):
        return bool(orc().orc_map_requires_fill_in(self.h, cf(ratio)))

    @staticmethod
    def fusion_weight(pose, last_pose, mult=1.0):
        return float(orc().orc_fusion_weight(P(f32(pose).reshape(16)), P(f32(last_pose).reshape(16)), cf(mult)))

    @property
    def count(self):
        return orc().orc_map_count(self.h)

    def surfels(self):
        n = self.count
        ptr = orc().orc_map_surfels(self.h)
        buf = (C.c_char * (n * 48)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(n, 12).copy()

    def unstable(self):
        n = orc().orc_map_unstable_count(self.h)
        ptr = orc().orc_map_unstable(self.h)
        buf = (C.c_char * (n * 48)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(n, 12).copy()

    def set_surfels(self, s):
        s = np.ascontiguousarray(s, np.float32)
        orc().orc_map_set_surfels(self.h, P(s), s.shape[0])

    def view(self, which):
        ch, dt = self.VIEWS[which]
        n = self.W * self.H * ch
        ptr = orc().orc_map_view(self.h, which)
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        a = np.frombuffer(buf, dtype=dt).copy()
        return a.reshape(self.H, self.W, ch) if ch > 1 else a.reshape(self.H, self.W)


# ---------------------------------------------------------------- segmentation restatement
class OrcSegParams(C.Structure):
    _fields_ = [("crfIterations", C.c_int), ("scaleFeaturesRGB", C.c_float), ("scaleFeaturesDepth", C.c_float),
                ("scaleFeaturesPos", C.c_float), ("weightAppearance", C.c_float), ("weightSmoothness", C.c_float),
                ("unaryThresholdNew", C.c_float), ("unaryKError", C.c_float), ("unaryWeightError", C.c_float),
                ("maxRelSizeNew", C.c_float), ("minRelSizeNew", C.c_float)]

    @staticmethod
    def default():
        p = OrcSegParams()
        orc().orc_seg_default_params(C.byref(p))
        return p


class OrcModelData(C.Structure):
    _fields_ = [("id", C.c_uint), ("superPixelCount", C.c_uint), ("avgConfidence", C.c_float),
                ("depthMean", C.c_float), ("depthStd", C.c_float), ("top", C.c_ushort), ("right", C.c_ushort),
                ("bottom", C.c_ushort), ("left", C.c_ushort)]


def slic(rgb, spixel=16, iters=5, coh=0.6):
    H, W = rgb.shape[:2]
    lab = np.zeros((H, W), np.int32)
    orc().orc_slic(P(np.ascontiguousarray(rgb, np.uint8)), W, H, spixel, iters, cf(coh), P(lab))
    return lab


def segment_crf(rgb, depth, model_ids, icp_errors, vert_confs, next_id, allow_new, params=None):
    """returns fullSeg (HxW u8), list of ModelData dicts, hasNew, slic labels, unary (N x L), low map"""
    H, W = depth.shape
    n = len(model_ids)
    prm = params or OrcSegParams.default()
    N = (W // 16) * (H // 16)
    L = n + (1 if allow_new else 0)
    icp = [f32(a) for a in icp_errors]
    vc = [f32(a) for a in vert_confs]
    icp_p = (C.c_void_p * n)(*[a.ctypes.data for a in icp])
    vc_p = (C.c_void_p * n)(*[a.ctypes.data for a in vc])
    ids = np.ascontiguousarray(model_ids, np.uint8)
    seg = np.zeros((H, W), np.uint8)
    md = (OrcModelData * (n + 1))()
    has_new = C.c_int(0)
    lab = np.zeros((H, W), np.int32)
    unary = np.zeros((N, L), np.float32)
    low = np.zeros(N, np.uint8)
    cnt = orc().orc_segment_crf(P(np.ascontiguousarray(rgb, np.uint8)), P(f32(depth)), W, H, n, P(ids), icp_p, vc_p,
                                int(next_id), int(allow_new), C.byref(prm), P(seg), md, C.byref(has_new), P(lab),
                                P(unary), P(low))
    mds = [dict(id=m.id, superPixelCount=m.superPixelCount, avgConfidence=m.avgConfidence, depthMean=m.depthMean,
                depthStd=m.depthStd, top=m.top, right=m.right, bottom=m.bottom, left=m.left) for m in md[:cnt]]
    return seg, mds, bool(has_new.value), lab, unary, low.reshape(H // 16, W // 16)
