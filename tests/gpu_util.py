"""Helpers for the -m gpu tests: torch owns device memory, the C ABI does the work."""
import ctypes as C

import numpy as np
import torch

import cofusion_b200 as cfb
from cofusion_b200 import check, lib

DEV = "cuda"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def cf(x):
    return C.c_float(float(x))


def P(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pitch(t):
    return C.c_size_t(t.stride(0) * t.element_size())


_scratch = None


def scratch():
    global _scratch
    if _scratch is None:
        n = lib().cfb_step_scratch_bytes()
        _scratch = torch.zeros(n, dtype=torch.uint8, device=DEV)
    return _scratch


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ---- seam-1 wrappers (device in, device/host out) -------------------------------------------------
def bilateral(depth, maxD):
    d = dev(depth.astype(np.float32))
    out = torch.empty_like(d)
    check(lib().cfb_bilateral_filter(P(d), pitch(d), d.shape[1], d.shape[0], cf(maxD), P(out), pitch(out), S()))
    return host(out)


def pyr_down_f(src):
    s = dev(src.astype(np.float32))
    out = torch.empty((s.shape[0] // 2, s.shape[1] // 2), dtype=torch.float32, device=DEV)
    check(lib().cfb_pyr_down_gauss_f(P(s), pitch(s), s.shape[1], s.shape[0], P(out), pitch(out), S()))
    return host(out)


def pyr_down_u8(src):
    s = dev(src.astype(np.uint8))
    out = torch.empty((s.shape[0] // 2, s.shape[1] // 2), dtype=torch.uint8, device=DEV)
    check(lib().cfb_pyr_down_uchar_gauss(P(s), pitch(s), s.shape[1], s.shape[0], P(out), pitch(out), S()))
    return host(out)


def create_vmap(depth, K, cutoff):
    d = dev(depth.astype(np.float32))
    H, W = d.shape
    out = torch.empty((3 * H, W), dtype=torch.float32, device=DEV)
    fx, fy, cx, cy = K
    check(lib().cfb_create_vmap(cf(fx), cf(fy), cf(cx), cf(cy), P(d), pitch(d), W, H, P(out), pitch(out), cf(cutoff), S()))
    return host(out)


def create_nmap(vmap):
    v = dev(vmap.astype(np.float32))
    H3, W = v.shape
    out = torch.empty_like(v)
    check(lib().cfb_create_nmap(P(v), pitch(v), W, H3 // 3, P(out), pitch(out), S()))
    return host(out)


def copy_maps(v4, n4):
    v = dev(v4.astype(np.float32))
    n = dev(n4.astype(np.float32))
    H, W = v4.shape[:2]
    vo = torch.empty((3 * H, W), dtype=torch.float32, device=DEV)
    no = torch.empty((3 * H, W), dtype=torch.float32, device=DEV)
    check(lib().cfb_copy_maps(P(v), P(n), W, H, P(vo), pitch(vo), P(no), pitch(no), S()))
    return host(vo), host(no)


def resize_map(m, normalize):
    i = dev(m.astype(np.float32))
    H3, W = i.shape
    H = H3 // 3
    out = torch.empty((3 * (H // 2), W // 2), dtype=torch.float32, device=DEV)
    fn = lib().cfb_resize_nmap if normalize else lib().cfb_resize_vmap
    check(fn(P(i), pitch(i), W, H, P(out), pitch(out), S()))
    return host(out)


def transform_maps(v, n, R, t):
    vs, ns = dev(v.astype(np.float32)), dev(n.astype(np.float32))
    H3, W = vs.shape
    vd, nd = torch.empty_like(vs), torch.empty_like(ns)
    R = np.ascontiguousarray(R, np.float32).reshape(9)
    t = np.ascontiguousarray(t, np.float32).reshape(3)
    check(lib().cfb_tranform_maps(P(vs), pitch(vs), P(ns), pitch(ns), W, H3 // 3, fptr(R), fptr(t), P(vd), pitch(vd),
                                  P(nd), pitch(nd), S()))
    return host(vd), host(nd)


def vertices_to_depth(v4, cutoff):
    v = dev(v4.astype(np.float32))
    H, W = v4.shape[:2]
    out = torch.empty((H, W), dtype=torch.float32, device=DEV)
    check(lib().cfb_vertices_to_depth(P(v), W, H, P(out), pitch(out), cf(cutoff), S()))
    return host(out)


def rgb_to_intensity(img):
    i = dev(img.astype(np.uint8))
    H, W, ch = img.shape
    out = torch.empty((H, W), dtype=torch.uint8, device=DEV)
    check(lib().cfb_image_bgr_to_intensity(P(i), C.c_size_t(W * ch), ch, W, H, P(out), pitch(out), S()))
    return host(out)


def derivative_images(img):
    i = dev(img.astype(np.uint8))
    H, W = img.shape
    dx = torch.empty((H, W), dtype=torch.int16, device=DEV)
    dy = torch.empty((H, W), dtype=torch.int16, device=DEV)
    check(lib().cfb_compute_derivative_images(P(i), pitch(i), W, H, P(dx), P(dy), pitch(dx), S()))
    return host(dx), host(dy)


def project_cloud(depth, K):
    d = dev(depth.astype(np.float32))
    H, W = d.shape
    out = torch.empty((H, W * 3), dtype=torch.float32, device=DEV)
    fx, fy, cx, cy = K
    check(lib().cfb_project_to_point_cloud(P(d), pitch(d), W, H, cf(fx), cf(fy), cf(cx), cf(cy), P(out), pitch(out), S()))
    return host(out)


def icp_step(Rcurr, tcurr, vc, nc, Rprev_inv, tprev, K, vp, np_, dist, angle, want_error=False):
    vc_, nc_, vp_, np__ = [dev(a.astype(np.float32)) for a in (vc, nc, vp, np_)]
    H3, W = vc.shape
    H = H3 // 3
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    res = np.zeros(2, np.float32)
    err = torch.full((H, W), -1.0, dtype=torch.float32, device=DEV) if want_error else None
    fx, fy, cx, cy = K
    f = lambda a, n: fptr(np.ascontiguousarray(a, np.float32).reshape(n))
    Rc, tc, Rp, tp = [np.ascontiguousarray(a, np.float32).reshape(-1) for a in (Rcurr, tcurr, Rprev_inv, tprev)]
    check(lib().cfb_icp_step(fptr(Rc), fptr(tc), P(vc_), pitch(vc_), P(nc_), pitch(nc_), fptr(Rp), fptr(tp), cf(fx),
                             cf(fy), cf(cx), cf(cy), P(vp_), pitch(vp_), P(np__), pitch(np__), cf(dist), cf(angle),
                             W, H, P(scratch()), fptr(A), fptr(b), fptr(res), P(err),
                             C.c_size_t(W * 4), S()))
    return A.reshape(6, 6), b, res, (host(err) if want_error else None)


def rgb_residual(minScale, dIdx, dIdy, lastDepth, nextDepth, lastImage, nextImage, maxDelta, kt, krkinv):
    H, W = lastDepth.shape
    gx, gy = dev(dIdx.astype(np.int16)), dev(dIdy.astype(np.int16))
    ld, nd = dev(lastDepth.astype(np.float32)), dev(nextDepth.astype(np.float32))
    li, ni = dev(lastImage.astype(np.uint8)), dev(nextImage.astype(np.uint8))
    corres = torch.zeros((H, W, 4), dtype=torch.int32, device=DEV)
    sigma, count = C.c_int(0), C.c_int(0)
    kt = np.ascontiguousarray(kt, np.float32)
    kr = np.ascontiguousarray(krkinv, np.float32).reshape(9)
    check(lib().cfb_compute_rgb_residual(cf(minScale), P(gx), P(gy), pitch(gx), P(ld), P(nd), pitch(ld), P(li), P(ni),
                                         pitch(li), P(corres), P(scratch()), cf(maxDelta), fptr(kt), fptr(kr), W, H,
                                         C.byref(sigma), C.byref(count), S()))
    return host(corres), sigma.value, count.value


def rgb_step(corres, sigma, cloud, K, dIdx, dIdy, sobelScale):
    H, W = dIdx.shape
    c = dev(np.ascontiguousarray(corres, np.int32))
    cl = dev(cloud.astype(np.float32))
    gx, gy = dev(dIdx.astype(np.int16)), dev(dIdy.astype(np.int16))
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    check(lib().cfb_rgb_step(P(c), cf(sigma), P(cl), pitch(cl), cf(K[0]), cf(K[1]), P(gx), P(gy), pitch(gx),
                             cf(sobelScale), W, H, P(scratch()), fptr(A), fptr(b), S()))
    return A.reshape(6, 6), b


def so3_step(lastImage, nextImage, imageBasis, kinv, krlr):
    H, W = lastImage.shape
    li, ni = dev(lastImage.astype(np.uint8)), dev(nextImage.astype(np.uint8))
    A = np.zeros(9, np.float32)
    b = np.zeros(3, np.float32)
    res = np.zeros(2, np.float32)
    m = [np.ascontiguousarray(x, np.float32).reshape(9) for x in (imageBasis, kinv, krlr)]
    check(lib().cfb_so3_step(P(li), P(ni), pitch(li), fptr(m[0]), fptr(m[1]), fptr(m[2]), W, H, P(scratch()), fptr(A),
                             fptr(b), fptr(res), S()))
    return A.reshape(3, 3), b, res
