"""ctypes binding of libacmmp_b200.so (the C ABI declared in include/acmmp_b200.h).

Used by tests/, bench.py and __graft_entry__.py.  The product's host side is the C++ `ACMMP`
class in ../host (same surface as the reference's ACMMP.h:57-111); this module is the thin
Python view of the same C ABI.  There is no CPU fallback: if the shared object is missing,
importing `lib()` raises, and without an sm_100 device `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("ACMMP_B200_LIB", PKG_DIR / "lib" / "libacmmp_b200.so"))     # override: development aid

MODEL_PINHOLE = 0
MODEL_SPHERE = 11


class Camera(C.Structure):
    """Mirror of the reference `struct Camera` (main.h:40-54), 120 bytes."""
    _fields_ = [
        ("model", C.c_int32),
        ("params", C.c_float * 4),
        ("R", C.c_float * 9),
        ("t", C.c_float * 3),
        ("K", C.c_float * 9),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("depth_min", C.c_float),
        ("depth_max", C.c_float),
    ]


class Params(C.Structure):
    """Mirror of the reference `struct PatchMatchParams` (ACMMP.h:32-55), 68 bytes."""
    _fields_ = [
        ("max_iterations", C.c_int32), ("patch_size", C.c_int32), ("num_images", C.c_int32),
        ("max_image_size", C.c_int32), ("radius_increment", C.c_int32),
        ("sigma_spatial", C.c_float), ("sigma_color", C.c_float), ("top_k", C.c_int32),
        ("baseline", C.c_float), ("depth_min", C.c_float), ("depth_max", C.c_float),
        ("disparity_min", C.c_float), ("disparity_max", C.c_float),
        ("scaled_cols", C.c_float), ("scaled_rows", C.c_float),
        ("geom_consistency", C.c_uint8), ("planar_prior", C.c_uint8), ("multi_geometry", C.c_uint8),
        ("hierarchy", C.c_uint8), ("upsample", C.c_uint8), ("pad_", C.c_uint8 * 3),
    ]


assert C.sizeof(Camera) == 120 and C.sizeof(Params) == 68


def make_camera(model, R, t, K=None, sphere=None, width=0, height=0, depth_min=0.0, depth_max=0.0) -> Camera:
    cam = Camera()
    cam.model = model
    R = np.asarray(R, dtype=np.float32).reshape(9)
    t = np.asarray(t, dtype=np.float32).reshape(3)
    for i in range(9):
        cam.R[i] = float(R[i])
    for i in range(3):
        cam.t[i] = float(t[i])
    if K is not None:
        K = np.asarray(K, dtype=np.float32).reshape(9)
        for i in range(9):
            cam.K[i] = float(K[i])
    if sphere is not None:
        for i, v in enumerate(sphere):
            cam.params[i] = float(v)
    cam.width, cam.height = int(width), int(height)
    cam.depth_min, cam.depth_max = float(depth_min), float(depth_max)
    return cam


_EXPORTS = [
    "acmmp_default_params", "acmmp_version", "acmmp_abi_sizeof_camera", "acmmp_abi_sizeof_params",
    "acmmp_create", "acmmp_destroy", "acmmp_last_error", "acmmp_set_views", "acmmp_set_views_device",
    "acmmp_set_geom_consistency", "acmmp_set_hierarchy", "acmmp_set_planar_prior", "acmmp_set_max_iterations",
    "acmmp_get_params", "acmmp_reset_modes", "acmmp_park", "acmmp_reserve_device_memory", "acmmp_reserve_pinned", "acmmp_pool_alloc", "acmmp_pool_free", "acmmp_last_jbu_ms", "acmmp_set_depth_maps", "acmmp_set_depth_maps_device", "acmmp_set_planes",
    "acmmp_set_hierarchy_inputs", "acmmp_next_level", "acmmp_next_level_device", "acmmp_result_host", "acmmp_set_planar_prior_inputs", "acmmp_support_points",
    "acmmp_planar_prior_from_triangles", "acmmp_download_prior", "acmmp_set_seed",
    "acmmp_set_plane_now_semantics", "acmmp_set_sphere_tap_pruning", "acmmp_run_patch_match", "acmmp_run_patch_match_resident", "acmmp_download_result", "acmmp_random_init", "acmmp_checkerboard_pass",
    "acmmp_finalize", "acmmp_synchronize", "acmmp_get_result", "acmmp_width", "acmmp_height",
    "acmmp_device_buffers", "acmmp_export_depth_device", "acmmp_export_depth_device_sync", "acmmp_download_state", "acmmp_upload_state",
    "acmmp_jbu", "acmmp_jbu_device", "acmmp_probe_ncc", "acmmp_probe_coords", "acmmp_probe_geom", "acmmp_probe_warp",
    "acmmp_probe_initcost", "acmmp_last_timings", "acmmp_launch_count",
    "acmmp_fusion_create", "acmmp_fusion_destroy", "acmmp_fusion_last_error", "acmmp_fusion_set_view", "acmmp_fusion_set_view_colour", "acmmp_fusion_set_view_device",
    "acmmp_fusion_run", "acmmp_fusion_run_ply", "acmmp_fusion_last_flags",
]

_lib = None


def lib() -> C.CDLL:
    """Load libacmmp_b200.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    l = C.CDLL(str(LIB_PATH))
    for name in _EXPORTS:
        getattr(l, name)          # AttributeError when the library does not export a declared symbol
    l.acmmp_version.restype = C.c_char_p
    l.acmmp_last_error.restype = C.c_char_p
    l.acmmp_last_error.argtypes = [C.c_void_p]
    l.acmmp_launch_count.restype = C.c_int64
    l.acmmp_launch_count.argtypes = [C.c_void_p]
    l.acmmp_set_seed.argtypes = [C.c_void_p, C.c_uint64]
    l.acmmp_last_jbu_ms.restype = C.c_float
    _lib = l
    return l


def exported_symbols():
    return list(_EXPORTS)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class AcmmpError(RuntimeError):
    pass


class Context:
    """One reference view being processed (== one reference `ACMMP` object)."""

    def __init__(self, device: int = 0):
        self._l = lib()
        self._h = C.c_void_p()
        rc = self._l.acmmp_create(C.byref(self._h), C.c_int(device))
        if rc != 0:
            raise AcmmpError(f"acmmp_create failed with {rc}: no usable sm_100 CUDA device (no CPU fallback exists)")
        self._keep = []
        self.W = self.H = 0
        self.n = 0

    def close(self):
        if self._h:
            self._l.acmmp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            msg = self._l.acmmp_last_error(self._h)
            raise AcmmpError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # ---- inputs -------------------------------------------------------------------------
    def set_views(self, images, cams):
        n = len(images)
        imgs = [_f32(im) for im in images]
        ptrs = (C.POINTER(C.c_float) * n)(*[_fp(im) for im in imgs])
        ws = (C.c_int32 * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_int32 * n)(*[im.shape[0] for im in imgs])
        carr = (Camera * n)(*cams)
        self._ck(self._l.acmmp_set_views(self._h, C.c_int(n), ptrs, ws, hs, carr), "acmmp_set_views")
        self.H, self.W = imgs[0].shape
        self.n = n

    def set_views_device(self, dev_ptrs, widths, heights, cams):
        n = len(dev_ptrs)
        ptrs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in dev_ptrs])
        ws = (C.c_int32 * n)(*widths)
        hs = (C.c_int32 * n)(*heights)
        carr = (Camera * n)(*cams)
        self._ck(self._l.acmmp_set_views_device(self._h, C.c_int(n), ptrs, ws, hs, carr), "acmmp_set_views_device")
        self.H, self.W = int(heights[0]), int(widths[0])
        self.n = n

    def set_geom_consistency(self, multi_geometry=False):
        self._ck(self._l.acmmp_set_geom_consistency(self._h, C.c_int(1 if multi_geometry else 0)), "set_geom")

    def set_hierarchy(self):
        self._ck(self._l.acmmp_set_hierarchy(self._h), "set_hierarchy")

    def set_planar_prior(self):
        self._ck(self._l.acmmp_set_planar_prior(self._h), "set_planar_prior")

    def set_max_iterations(self, n):
        self._ck(self._l.acmmp_set_max_iterations(self._h, C.c_int(n)), "set_max_iterations")

    def reset_modes(self):
        self._ck(self._l.acmmp_reset_modes(self._h), "acmmp_reset_modes")

    def park(self, keep_prior=False, keep_host_result=False):
        """Give everything but the stage state (planes, costs, coarse planes) back to the device's pool; go on with
        set_views / set_views_device of the same shapes (include/acmmp_b200.h: acmmp_park)."""
        self._ck(self._l.acmmp_park(self._h, C.c_int(int(keep_prior)), C.c_int(int(keep_host_result))), "acmmp_park")

    def params(self) -> Params:
        p = Params()
        self._ck(self._l.acmmp_get_params(self._h, C.byref(p)), "get_params")
        return p

    def set_depth_maps(self, maps):
        """maps[0] may be None: the reference view's depth map is then taken from the device-resident state."""
        n = len(maps)
        ms = [None if m is None else _f32(m) for m in maps]
        ptrs = (C.POINTER(C.c_float) * n)(*[C.POINTER(C.c_float)() if m is None else _fp(m) for m in ms])
        ws = (C.c_int32 * n)(*[self.W if m is None else m.shape[1] for m in ms])
        hs = (C.c_int32 * n)(*[self.H if m is None else m.shape[0] for m in ms])
        self._ck(self._l.acmmp_set_depth_maps(self._h, C.c_int(n), ptrs, ws, hs), "acmmp_set_depth_maps")

    def next_level(self, images, cams):
        """Move to the next pyramid level on the device (JBU + hierarchy inputs), see acmmp_next_level."""
        n = len(images)
        imgs = [_f32(im) for im in images]
        ptrs = (C.POINTER(C.c_float) * n)(*[_fp(im) for im in imgs])
        ws = (C.c_int32 * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_int32 * n)(*[im.shape[0] for im in imgs])
        carr = (Camera * n)(*cams)
        self._ck(self._l.acmmp_next_level(self._h, C.c_int(n), ptrs, ws, hs, carr), "acmmp_next_level")
        self.H, self.W = imgs[0].shape
        self.n = n

    def result_host(self):
        """Zero-copy numpy views of the pinned host result buffers (valid until the next run)."""
        pp, pc = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        self._ck(self._l.acmmp_result_host(self._h, C.byref(pp), C.byref(pc)), "acmmp_result_host")
        planes = np.ctypeslib.as_array(pp, shape=(self.H, self.W, 4))
        costs = np.ctypeslib.as_array(pc, shape=(self.H, self.W))
        return planes, costs

    def set_depth_maps_device(self, dev_ptrs, widths, heights):
        n = len(dev_ptrs)
        ptrs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in dev_ptrs])
        ws = (C.c_int32 * n)(*widths)
        hs = (C.c_int32 * n)(*heights)
        self._ck(self._l.acmmp_set_depth_maps_device(self._h, C.c_int(n), ptrs, ws, hs), "acmmp_set_depth_maps_device")

    def set_planes(self, planes4, costs):
        p, c = _f32(planes4), _f32(costs)
        self._ck(self._l.acmmp_set_planes(self._h, _fp(p), _fp(c)), "acmmp_set_planes")

    def set_hierarchy_inputs(self, coarse_planes4, fine_depth):
        cp, fd = _f32(coarse_planes4), _f32(fine_depth)
        sh, sw = cp.shape[0], cp.shape[1]
        self._ck(self._l.acmmp_set_hierarchy_inputs(self._h, _fp(cp), C.c_int(sw), C.c_int(sh), _fp(fd)),
                 "acmmp_set_hierarchy_inputs")

    def set_planar_prior_inputs(self, plane_params4, masks):
        pp = _f32(plane_params4).reshape(-1, 4)
        mk = _f32(masks)
        self._ck(self._l.acmmp_set_planar_prior_inputs(self._h, _fp(pp), C.c_int(pp.shape[0]), _fp(mk)),
                 "acmmp_set_planar_prior_inputs")

    def support_points(self):
        """GetSupportPoints on the device (costs of the current state) -> int32 array [n, 2] of (x, y), reference order."""
        cap = ((self.W + 4) // 5) * ((self.H + 4) // 5)
        xy = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        self._ck(self._l.acmmp_support_points(self._h, xy.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(cap), C.byref(n)),
                 "acmmp_support_points")
        return xy[: n.value].copy()

    def planar_prior_from_triangles(self, tri_xy):
        """tri_xy: int32 [n, 3, 2] triangle vertices (x, y) inside the image, ids 1..n in this order: plane fit,
        rasteriser, depth-range test and prior upload on the device."""
        t = np.ascontiguousarray(tri_xy, np.int32).reshape(-1, 6)
        self._ck(self._l.acmmp_planar_prior_from_triangles(self._h, t.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(t.shape[0])),
                 "acmmp_planar_prior_from_triangles")

    def download_prior(self):
        """(prior planes [H, W, 4], triangle ids [H, W] uint32) as the device holds them."""
        pp = np.empty((self.H, self.W, 4), np.float32)
        mk = np.empty((self.H, self.W), np.uint32)
        self._ck(self._l.acmmp_download_prior(self._h, _fp(pp), mk.ctypes.data_as(C.POINTER(C.c_uint32))), "acmmp_download_prior")
        return pp, mk

    def set_seed(self, seed):
        self._ck(self._l.acmmp_set_seed(self._h, C.c_uint64(seed)), "acmmp_set_seed")

    def set_sphere_tap_pruning(self, relative_weight: float):
        """SPHERE: skip window taps whose bilateral weight is below relative_weight x (sum of the 36); 0 = sample all."""
        self._ck(self._l.acmmp_set_sphere_tap_pruning(self._h, C.c_float(relative_weight)), "acmmp_set_sphere_tap_pruning")

    def set_plane_now_semantics(self, as_compiled: bool):
        self._ck(self._l.acmmp_set_plane_now_semantics(self._h, C.c_int(1 if as_compiled else 0)), "set_plane_now_semantics")

    # ---- compute ------------------------------------------------------------------------
    def run_patch_match(self, download=True):
        """One stage (RunPatchMatch).  download=False keeps the result on the device (GPU-resident chaining)."""
        if download:
            self._ck(self._l.acmmp_run_patch_match(self._h), "acmmp_run_patch_match")
        else:
            self._ck(self._l.acmmp_run_patch_match_resident(self._h), "acmmp_run_patch_match_resident")

    def download_result(self):
        self._ck(self._l.acmmp_download_result(self._h), "acmmp_download_result")

    def random_init(self):
        self._ck(self._l.acmmp_random_init(self._h), "acmmp_random_init")

    def checkerboard_pass(self, colour, it):
        self._ck(self._l.acmmp_checkerboard_pass(self._h, C.c_int(colour), C.c_int(it)), "acmmp_checkerboard_pass")

    def finalize(self):
        self._ck(self._l.acmmp_finalize(self._h), "acmmp_finalize")

    def synchronize(self):
        self._ck(self._l.acmmp_synchronize(self._h), "acmmp_synchronize")

    def get_result(self):
        planes = np.empty((self.H, self.W, 4), np.float32)
        costs = np.empty((self.H, self.W), np.float32)
        self._ck(self._l.acmmp_get_result(self._h, _fp(planes), _fp(costs)), "acmmp_get_result")
        return planes, costs

    def device_buffers(self):
        p, c = C.c_void_p(), C.c_void_p()
        self._ck(self._l.acmmp_device_buffers(self._h, C.byref(p), C.byref(c)), "acmmp_device_buffers")
        return p.value, c.value

    def export_depth_device(self, dev_ptr):
        self._ck(self._l.acmmp_export_depth_device(self._h, C.c_void_p(int(dev_ptr))), "acmmp_export_depth_device")

    def download_state(self, rand=True, pre_costs=True):
        planes = np.empty((self.H, self.W, 4), np.float32)
        costs = np.empty((self.H, self.W), np.float32)
        views = np.empty((self.H, self.W), np.uint32)
        rand6 = np.empty((self.H, self.W, 6), np.uint32) if rand else None
        pre = np.empty((self.H, self.W), np.float32) if pre_costs else None
        self._ck(self._l.acmmp_download_state(
            self._h, _fp(planes), _fp(costs), views.ctypes.data_as(C.POINTER(C.c_uint32)),
            rand6.ctypes.data_as(C.POINTER(C.c_uint32)) if rand else None,
            _fp(pre) if pre_costs else None), "acmmp_download_state")
        return dict(planes=planes, costs=costs, views=views, rand=rand6, pre_costs=pre)

    def upload_state(self, planes=None, costs=None, views=None, rand=None, pre_costs=None):
        keep = []

        def fp(a, dt, ct):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ct))
        self._ck(self._l.acmmp_upload_state(self._h, fp(planes, np.float32, C.c_float), fp(costs, np.float32, C.c_float),
                                            fp(views, np.uint32, C.c_uint32), fp(rand, np.uint32, C.c_uint32),
                                            fp(pre_costs, np.float32, C.c_float)), "acmmp_upload_state")

    # ---- probes -------------------------------------------------------------------------
    def probe_ncc(self, planes4, view):
        p = _f32(planes4)
        out = np.empty((self.H, self.W), np.float32)
        self._ck(self._l.acmmp_probe_ncc(self._h, _fp(p), C.c_int(view), _fp(out)), "acmmp_probe_ncc")
        return out

    def probe_coords(self, planes4, view, variant=0):
        """(H, W, 36, 2): the fetch coordinates of quad_ncc's 36 samples, reference tap order.
        variant 1 / 2 (SPHERE): through the packed two-hypothesis form, as its first / second hypothesis."""
        p = _f32(planes4)
        out = np.empty((self.H, self.W, 36, 2), np.float32)
        self._ck(self._l.acmmp_probe_coords(self._h, _fp(p), C.c_int(view), C.c_int(variant), _fp(out)), "acmmp_probe_coords")
        return out

    def probe_geom(self, planes4, view):
        p = _f32(planes4)
        out = np.empty((self.H, self.W), np.float32)
        self._ck(self._l.acmmp_probe_geom(self._h, _fp(p), C.c_int(view), _fp(out)), "acmmp_probe_geom")
        return out

    def probe_warp(self, planes4, view):
        p = _f32(planes4)
        out = np.empty((self.H, self.W, 4), np.float32)
        self._ck(self._l.acmmp_probe_warp(self._h, _fp(p), C.c_int(view), _fp(out)), "acmmp_probe_warp")
        return out

    def probe_initcost(self, planes4):
        p = _f32(planes4)
        out = np.empty((self.H, self.W), np.float32)
        views = np.empty((self.H, self.W), np.uint32)
        self._ck(self._l.acmmp_probe_initcost(self._h, _fp(p), _fp(out), views.ctypes.data_as(C.POINTER(C.c_uint32))),
                 "acmmp_probe_initcost")
        return out, views

    def timings(self):
        t = (C.c_float * 8)()
        self._ck(self._l.acmmp_last_timings(self._h, t), "acmmp_last_timings")
        return dict(init_ms=t[0], pass_sum_ms=t[1], finalize_ms=t[2], n_pass=int(t[3]), last_pass_ms=t[4], jbu_ms=t[5])

    def launch_count(self):
        return int(self._l.acmmp_launch_count(self._h))


def reserve_device_memory(device, nbytes):
    """One allocation that the device's pool carves its blocks out of (acmmp_reserve_device_memory).  Returns the status
    code: 0, or ACMMP_E_ARG when the device already has a reservation."""
    return int(lib().acmmp_reserve_device_memory(C.c_int(device), C.c_size_t(int(nbytes))))


def pool_alloc(device, nbytes):
    """A device buffer out of the device's pool (acmmp_pool_alloc) -> device pointer (int)."""
    p = C.c_void_p()
    rc = lib().acmmp_pool_alloc(C.c_int(device), C.c_size_t(int(nbytes)), C.byref(p))
    if rc != 0:
        raise RuntimeError(f"acmmp_pool_alloc failed ({rc})")
    return int(p.value)


def pool_free(device, ptr):
    rc = lib().acmmp_pool_free(C.c_int(device), C.c_void_p(int(ptr)))
    if rc != 0:
        raise RuntimeError(f"acmmp_pool_free failed ({rc})")


def jbu(image, coarse_depth, device=0):
    """Joint-bilateral upsampling of a coarse depth map (replaces RunJBU, ACMMP.cpp:1071-1122)."""
    l = lib()
    img, dep = _f32(image), _f32(coarse_depth)
    out = np.empty_like(img)
    rc = l.acmmp_jbu(C.c_int(device), _fp(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _fp(dep),
                     C.c_int(dep.shape[1]), C.c_int(dep.shape[0]), _fp(out))
    if rc != 0:
        raise AcmmpError(f"acmmp_jbu failed ({rc})")
    return out


def last_jbu_ms() -> float:
    return float(lib().acmmp_last_jbu_ms())


class Fusion:
    """Depth-map fusion of a scene on the device (acmmp_fusion_*; reference RunFusionCuda / SimpleFusionKernel,
    ACMMP.cu:1664-2105).  Views are set once (camera scaled to the depth map's size), then every reference view is fused
    against its source views; points come back compacted in pixel order as an [n, 9] float32 array (PointList)."""

    def __init__(self, n_views: int, device: int = 0):
        self._l = lib()
        self._l.acmmp_fusion_last_error.restype = C.c_char_p
        h = C.c_void_p()
        rc = self._l.acmmp_fusion_create(C.c_int(device), C.c_int(n_views), C.byref(h))
        if rc != 0:
            raise AcmmpError(f"acmmp_fusion_create failed ({rc}): no sm_100 device? there is no CPU fallback")
        self._h, self.n = h, n_views
        self.sizes = {}
        self.kernel_ms = 0.0

    def _ck(self, rc, what, allow=()):
        if rc != 0 and rc not in allow:
            raise AcmmpError(f"{what} failed ({rc}): {self._l.acmmp_fusion_last_error(self._h).decode()}")
        return rc

    def set_view(self, index, cam, depth, normals, gray):
        d, n3, g = _f32(depth), _f32(normals), _f32(gray)
        h, w = d.shape
        assert n3.shape == (h, w, 3) and g.shape == (h, w)
        self._ck(self._l.acmmp_fusion_set_view(self._h, C.c_int(index), C.byref(cam), C.c_int(w), C.c_int(h), _fp(d), _fp(n3), _fp(g)),
                 "acmmp_fusion_set_view")
        self.sizes[index] = (h, w)

    def set_view_colour(self, index, bgr):
        """Optional colour image of a view already set: uint8 [h, w, 3] in OpenCV's B, G, R order."""
        c = np.ascontiguousarray(bgr, np.uint8)
        h, w = self.sizes[index]
        assert c.shape == (h, w, 3)
        self._ck(self._l.acmmp_fusion_set_view_colour(self._h, C.c_int(index), c.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(w), C.c_int(h)),
                 "acmmp_fusion_set_view_colour")

    def run(self, ref, src_indices):
        src = np.ascontiguousarray(src_indices, np.int32)
        h, w = self.sizes[ref]
        cap = max(h * w // 4, 1024)
        while True:
            pts = np.empty((cap, 9), np.float32)
            n, ms = C.c_int(0), C.c_float(0)
            rc = self._ck(self._l.acmmp_fusion_run(self._h, C.c_int(ref), C.c_int(len(src)), src.ctypes.data_as(C.POINTER(C.c_int32)), _fp(pts),
                                                   C.c_int(cap), C.byref(n), C.byref(ms)), "acmmp_fusion_run", allow=(-1,))
            if rc == 0:
                self.kernel_ms = float(ms.value)
                return pts[: n.value].copy()
            if n.value <= cap:
                self._ck(rc, "acmmp_fusion_run")
            cap = n.value

    def run_ply(self, ref, src_indices):
        """The points of one reference view as the PLY file's 27-byte vertex records -> uint8 [n, 27]."""
        src = np.ascontiguousarray(src_indices, np.int32)
        h, w = self.sizes[ref]
        rec = np.empty((h * w, 27), np.uint8)
        n, ms = C.c_int(0), C.c_float(0)
        self._ck(self._l.acmmp_fusion_run_ply(self._h, C.c_int(ref), C.c_int(len(src)), src.ctypes.data_as(C.POINTER(C.c_int32)),
                                              rec.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(h * w), C.byref(n), C.byref(ms)), "acmmp_fusion_run_ply")
        self.kernel_ms = float(ms.value)
        return rec[: n.value].copy()

    def last_flags(self, ref):
        h, w = self.sizes[ref]
        flags = np.zeros((h, w), np.uint8)
        self._ck(self._l.acmmp_fusion_last_flags(self._h, C.c_int(ref), flags.ctypes.data_as(C.POINTER(C.c_ubyte))), "acmmp_fusion_last_flags")
        return flags

    def close(self):
        if self._h:
            self._l.acmmp_fusion_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
