"""ctypes view of oracle/_ref/libacmmp_oracle.so (oracle/acmmp_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
LIB = ORACLE_DIR / "_ref" / "libacmmp_oracle.so"


class OrcCamera(C.Structure):
    _fields_ = [("model", C.c_int32), ("params", C.c_float * 4), ("R", C.c_float * 9), ("t", C.c_float * 3),
                ("K", C.c_float * 9), ("width", C.c_int32), ("height", C.c_int32), ("depth_min", C.c_float),
                ("depth_max", C.c_float)]


class OrcImage(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("width", C.c_int), ("height", C.c_int)]


class OrcPassFlags(C.Structure):
    _fields_ = [("geom", C.c_int), ("prior", C.c_int), ("hierarchy", C.c_int), ("as_compiled", C.c_int),
                ("depth_min", C.c_float), ("depth_max", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "_ref/libacmmp_oracle.so"], check=True)
        l = C.CDLL(str(LIB))
        l.orc_tex2d.restype = C.c_float
        l.orc_tex2d.argtypes = [C.POINTER(OrcImage), C.c_float, C.c_float]
        l.orc_curand_uniform.restype = C.c_float
        l.orc_curand.restype = C.c_uint32
        l.orc_curand_init.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        l.orc_random_init.argtypes = [C.c_int, C.POINTER(OrcImage), C.POINTER(OrcCamera), C.c_float, C.c_float, C.c_uint64,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
        _lib = l
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _cams(cams):
    arr = (OrcCamera * len(cams))()
    for i, c in enumerate(cams):
        C.memmove(C.byref(arr[i]), C.byref(c), 120)
    return arr


def _imgs(images):
    keep = [_f32(im) for im in images]
    arr = (OrcImage * len(keep))()
    for i, im in enumerate(keep):
        arr[i].data = _fp(im)
        arr[i].width, arr[i].height = im.shape[1], im.shape[0]
    return arr, keep


def num_threads():
    return int(lib().orc_num_threads())


def ncc_map(images, cams, planes4, view):
    l = lib()
    ims, keep = _imgs(images)
    cs = _cams(cams)
    pl = _f32(planes4)
    H, W = keep[0].shape
    out = np.empty((H, W), np.float32)
    l.orc_ncc_map(C.byref(ims[0]), C.byref(cs[0]), C.byref(ims[view]), C.byref(cs[view]), _fp(pl), _fp(out))
    return out


def geom_map(depth_maps, cams, planes4, view):
    l = lib()
    dms, keep = _imgs(depth_maps)
    cs = _cams(cams)
    pl = _f32(planes4)
    H, W = cams[0].height, cams[0].width
    out = np.empty((H, W), np.float32)
    l.orc_geom_map(C.byref(dms[view]), C.byref(cs[0]), C.byref(cs[view]), _fp(pl), _fp(out))
    return out


def warp_map(cams, planes4, view):
    l = lib()
    cs = _cams(cams)
    pl = _f32(planes4)
    H, W = cams[0].height, cams[0].width
    out = np.empty((H, W, 4), np.float32)
    l.orc_warp_map(C.byref(cs[0]), C.byref(cs[view]), _fp(pl), _fp(out))
    return out


def initcost_map(images, cams, planes4):
    l = lib()
    ims, keep = _imgs(images)
    cs = _cams(cams)
    pl = _f32(planes4)
    H, W = keep[0].shape
    out = np.empty((H, W), np.float32)
    views = np.empty((H, W), np.uint32)
    l.orc_initcost_map(C.c_int(len(images)), ims, cs, _fp(pl), _fp(out), _u32p(views))
    return out, views


def jbu(image, coarse_depth):
    l = lib()
    img, dep = _f32(image), _f32(coarse_depth)
    out = np.empty_like(img)
    l.orc_jbu(_fp(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _fp(dep), C.c_int(dep.shape[1]), C.c_int(dep.shape[0]), _fp(out))
    return out


def depth_normal(cam, planes4):
    l = lib()
    cs = _cams([cam])
    pl = _f32(planes4).copy()
    l.orc_depth_normal(C.byref(cs[0]), _fp(pl))
    return pl


def median_filter(planes4, costs, colour):
    l = lib()
    pl = _f32(planes4).copy()
    co = _f32(costs)
    l.orc_median_filter(C.c_int(pl.shape[1]), C.c_int(pl.shape[0]), _fp(pl), _fp(co), C.c_int(colour))
    return pl


def curand_states(seed, width, height):
    """{d, v0..v4} of curand_init(seed, y, x) for every pixel."""
    l = lib()
    out = np.empty((height, width, 6), np.uint32)
    st = (C.c_uint32 * 6)()
    for y in range(height):
        l.orc_curand_init(C.c_uint64(seed), C.c_uint64(y), C.c_uint64(0), st)
        for x in range(width):
            out[y, x] = np.frombuffer(st, dtype=np.uint32)
            l.orc_curand(st)
    return out


def random_init(images, cams, seed, with_costs=True):
    l = lib()
    ims, keep = _imgs(images)
    cs = _cams(cams)
    H, W = keep[0].shape
    planes = np.zeros((H, W, 4), np.float32)
    costs = np.zeros((H, W), np.float32)
    views = np.zeros((H, W), np.uint32)
    rand6 = np.zeros((H, W, 6), np.uint32)
    dmin = np.float32(cams[0].depth_min) * np.float32(0.6)      # ACMMP.cpp:645-646
    dmax = np.float32(cams[0].depth_max) * np.float32(1.2)
    l.orc_random_init(len(images), ims, cs, C.c_float(dmin), C.c_float(dmax), C.c_uint64(seed), _fp(planes), _fp(costs),
                      _u32p(views), _u32p(rand6), 1 if with_costs else 0)
    return dict(planes=planes, costs=costs, views=views, rand=rand6)


def checkerboard_pass(images, cams, state, colour, it, geom=False, prior=False, hierarchy=False, as_compiled=True,
                      depth_maps=None, prior_planes=None, plane_masks=None, late_planes=None, want_center=False):
    """One pass with read-old/write-new neighbour semantics; returns the new state dict.
    Planar-prior mode only: `want_center` adds out["center"] = what each updated pixel holds after the reference's
    IN-PASS write (ACMMP.cu:1283 / :1295); `late_planes` makes the re-reads of same-colour neighbours in the prior block
    (:1262, :1279, :1291) read that array -- the other extreme of the reference's data race (see prior_pass_race)."""
    l = lib()
    ims, keep = _imgs(images)
    cs = _cams(cams)
    H, W = keep[0].shape
    dms, keepd = (_imgs(depth_maps) if depth_maps is not None else (None, None))
    fl = OrcPassFlags(int(geom), int(prior), int(hierarchy), int(as_compiled),
                      float(np.float32(cams[0].depth_min) * np.float32(0.6)), float(np.float32(cams[0].depth_max) * np.float32(1.2)))
    planes_in, costs_in = _f32(state["planes"]), _f32(state["costs"])
    planes_out, costs_out = np.empty_like(planes_in), np.empty_like(costs_in)
    views = np.ascontiguousarray(state["views"], np.uint32).copy()
    rand6 = np.ascontiguousarray(state["rand"], np.uint32).copy()
    pre = _f32(state["pre_costs"]) if state.get("pre_costs") is not None else None
    pp = _f32(prior_planes) if prior_planes is not None else None
    pm = np.ascontiguousarray(plane_masks, np.uint32) if plane_masks is not None else None
    late = _f32(late_planes) if late_planes is not None else None
    center = planes_in.copy() if want_center else None
    l.orc_set_race_emulation(_fp(late) if late is not None else None, _fp(center) if center is not None else None)
    l.orc_checkerboard_pass(C.c_int(len(images)), ims, dms, cs, C.byref(fl), C.c_int(colour), C.c_int(it), _fp(planes_in),
                            _fp(costs_in), _fp(planes_out), _fp(costs_out), _fp(pre) if pre is not None else None,
                            _u32p(views), _u32p(rand6), _fp(pp) if pp is not None else None,
                            _u32p(pm) if pm is not None else None)
    l.orc_set_race_emulation(None, None)
    return dict(planes=planes_out, costs=costs_out, views=views, rand=rand6, pre_costs=state.get("pre_costs"), center=center)


def prior_pass_race(images, cams, state, colour, it, prior_planes, plane_masks, hierarchy=True):
    """The two extremes of the data race the reference has in planar-prior mode.  There a thread writes
    plane_hypotheses[center] in the MIDDLE of its work (the accepted neighbour, ACMMP.cu:1283 / :1295) while other threads
    re-read plane_hypotheses[positions[..]] -- six of the seven `near` positions of a direction are SAME-colour pixels,
    i.e. pixels being updated by this very launch -- at the same point of theirs (:1262, :1279, :1291).  Without the
    prior the only mid-pass writes are costs / view masks and the plane re-read comes long before anybody's final
    write, so the race never fires; with it the outcome depends on warp timing.
      early : every re-read sees the pre-pass plane (what the double-buffered B200 pass computes)
      late  : every re-read of a same-colour pixel sees that pixel's in-pass write
    Returns (early, late, sensitive): the two result states and the mask of pixels whose plane differs between them."""
    early = checkerboard_pass(images, cams, state, colour, it, prior=True, hierarchy=hierarchy, prior_planes=prior_planes,
                              plane_masks=plane_masks, want_center=True)
    late = checkerboard_pass(images, cams, state, colour, it, prior=True, hierarchy=hierarchy, prior_planes=prior_planes,
                             plane_masks=plane_masks, late_planes=early["center"])
    a, b = early["planes"], late["planes"]
    sensitive = ~np.all((np.abs(a - b) <= 1e-6 + 1e-6 * np.abs(b)) | (np.isnan(a) & np.isnan(b)), axis=-1)
    return early, late, sensitive


def fuse_view(cams, depths, normals, grays, ref, src_indices, colours=None):
    """SimpleFusionKernel (ACMMP.cu:1664-1814) for one reference view -> (points [h, w, 9], flags [h, w] bool).
    cams: cameras with width / height = the maps' size; colours: optional uint8 [h, w, 3] images in B, G, R order."""
    n = len(cams)
    carr = _cams(cams)
    for i in range(n):
        carr[i].width, carr[i].height = depths[i].shape[1], depths[i].shape[0]
    d = [_f32(x) for x in depths]
    nm = [_f32(x) for x in normals]
    g = [_f32(x) for x in grays]
    FP, BP = C.POINTER(C.c_float), C.POINTER(C.c_ubyte)
    bgr = None
    if colours is not None:
        keep = [np.ascontiguousarray(x, np.uint8) for x in colours]
        bgr = (BP * n)(*[x.ctypes.data_as(BP) for x in keep])
    h, w = d[ref].shape
    pts = np.zeros((h, w, 9), np.float32)
    flags = np.zeros((h, w), np.int32)
    src = (C.c_int * len(src_indices))(*[int(s) for s in src_indices])
    lib().orc_fuse_view(C.c_int(n), carr, (FP * n)(*[_fp(x) for x in d]), (FP * n)(*[_fp(x) for x in nm]), (FP * n)(*[_fp(x) for x in g]),
                        bgr, C.c_int(ref), C.c_int(len(src_indices)), src, _fp(pts), flags.ctypes.data_as(C.POINTER(C.c_int)))
    return pts, flags.astype(bool)
