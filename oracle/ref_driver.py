"""ctypes driver of oracle/_ref/libacmmp_ref.so -- TEST INFRASTRUCTURE ONLY.

libacmmp_ref.so is the UNMODIFIED reference (ACMMP.cu + ACMMP.cpp) compiled for sm_100 behind
oracle/ref_harness.cu (see there).  Only tests/, __graft_entry__.smoke() and bench.py's reference /
baseline legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import tempfile
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_LIB = ORACLE_DIR / "_ref" / "libacmmp_ref.so"

_lib = None


def available() -> bool:
    return REF_LIB.exists()


def lib():
    global _lib
    if _lib is None:
        if not REF_LIB.exists():
            raise RuntimeError(f"{REF_LIB} missing: run `make -C oracle` where /root/reference exists")
        l = C.CDLL(str(REF_LIB))
        l.ref_create.restype = C.c_void_p
        l.ref_run_patch_match.restype = C.c_float
        l.ref_launch_init.restype = C.c_float
        l.ref_launch_pass.restype = C.c_float
        l.ref_launch_finalize.restype = C.c_float
        l.ref_set_seed.argtypes = [C.c_uint64]
        l.ref_fusion_create.restype = C.c_void_p
        l.ref_fusion_run.restype = C.c_float
        _lib = l
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def write_dmb(path, arr):
    """dmb = int32 type=1, h, w, nb + float32 data (reference ACMMP.cpp:395-479)."""
    a = _f32(arr)
    h, w = a.shape[0], a.shape[1]
    nb = 1 if a.ndim == 2 else a.shape[2]
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", 1, h, w, nb))
        f.write(a.tobytes())


def read_dmb(path):
    with open(path, "rb") as f:
        t, h, w, nb = struct.unpack("<iiii", f.read(16))
        assert t == 1
        a = np.frombuffer(f.read(4 * h * w * nb), dtype=np.float32)
    return a.reshape(h, w) if nb == 1 else a.reshape(h, w, nb)


class RefACMMP:
    """One reference `ACMMP` object, fed in-memory inputs (raw float images + Camera structs)."""

    def __init__(self, images, cams, seed=1234, geom=False, multi_geom=False, hierarchy=False,
                 depth_maps=None, prev_planes=None, prev_costs=None, coarse_normals=None, coarse_costs=None,
                 coarse_depth=None, fine_depth=None, workdir=None, ref_image_id=0):
        from acmmp_b200 import Camera      # the struct mirror only (layout == reference Camera)
        self._l = lib()
        self._l.ref_set_seed(C.c_uint64(seed))
        self._h = C.c_void_p(self._l.ref_create())
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="acmmp_ref_")
            workdir = self._tmp.name
        self.workdir = workdir
        n = len(images)
        imgs = [_f32(im) for im in images]
        self.H, self.W = imgs[0].shape
        if geom:
            self._l.ref_set_geom(self._h, C.c_int(1 if multi_geom else 0))
        if hierarchy:
            self._l.ref_set_hierarchy(self._h)
        ptrs = (C.POINTER(C.c_float) * n)(*[_fp(im) for im in imgs])
        ws = (C.c_int * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_int * n)(*[im.shape[0] for im in imgs])
        carr = (Camera * n)(*cams)
        self._l.ref_set_images(self._h, C.c_int(n), ptrs, ws, hs, carr)
        folder = Path(workdir) / "ACMMP" / ("2333_%08d" % ref_image_id)
        folder.mkdir(parents=True, exist_ok=True)
        if geom:
            dm = [_f32(m) for m in depth_maps]
            dptrs = (C.POINTER(C.c_float) * n)(*[_fp(m) for m in dm])
            dws = (C.c_int * n)(*[m.shape[1] for m in dm])
            dhs = (C.c_int * n)(*[m.shape[0] for m in dm])
            self._l.ref_set_depths(self._h, C.c_int(n), dptrs, dws, dhs)
            # own previous state, reloaded by CudaSpaceInitialization (ACMMP.cpp:753-785)
            suffix = "depths_geom.dmb" if multi_geom else "depths.dmb"
            write_dmb(folder / suffix, prev_planes[..., 3])
            write_dmb(folder / "normals.dmb", prev_planes[..., :3])
            write_dmb(folder / "costs.dmb", prev_costs)
        if hierarchy:
            # ACMMP.cpp:788-844: depths.dmb = fine (JBU) depth, normals/costs.dmb = coarse level
            write_dmb(folder / "depths.dmb", fine_depth if fine_depth is not None else coarse_depth)
            write_dmb(folder / "normals.dmb", coarse_normals)
            write_dmb(folder / "costs.dmb", coarse_costs)
        self._l.ref_cuda_space_init(self._h, str(workdir).encode(), C.c_int(ref_image_id), C.c_int(1))
        self._check()

    def _check(self):
        e = self._l.ref_last_error()
        if e != 0:
            raise RuntimeError(f"reference harness reported CUDA error {e}")

    def close(self):
        if self._h:
            self._l.ref_destroy(self._h)
            self._h = None
        if self._tmp is not None:
            self._tmp.cleanup()
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_max_iterations(self, n):
        self._l.ref_set_max_iterations(self._h, C.c_int(n))

    def set_prior(self, plane_params4, masks):
        pp = _f32(plane_params4).reshape(-1, 4)
        mk = _f32(masks)
        self._l.ref_prior_init(self._h, _fp(pp), C.c_int(pp.shape[0]), _fp(mk))
        self._check()

    def set_planar_prior_flag(self):
        self._l.ref_set_planar_prior(self._h)

    def run_patch_match(self) -> float:
        ms = float(self._l.ref_run_patch_match(self._h))
        self._check()
        return ms

    def get_result(self):
        planes = np.empty((self.H, self.W, 4), np.float32)
        costs = np.empty((self.H, self.W), np.float32)
        self._l.ref_get_result(self._h, _fp(planes), _fp(costs))
        return planes, costs

    def launch_init(self) -> float:
        ms = float(self._l.ref_launch_init(self._h)); self._check(); return ms

    def launch_pass(self, colour, it) -> float:
        ms = float(self._l.ref_launch_pass(self._h, C.c_int(colour), C.c_int(it))); self._check(); return ms

    def launch_finalize(self) -> float:
        ms = float(self._l.ref_launch_finalize(self._h)); self._check(); return ms

    def download_state(self, rand=True, pre_costs=False):
        planes = np.empty((self.H, self.W, 4), np.float32)
        costs = np.empty((self.H, self.W), np.float32)
        views = np.empty((self.H, self.W), np.uint32)
        rand6 = np.empty((self.H, self.W, 6), np.uint32) if rand else None
        pre = np.empty((self.H, self.W), np.float32) if pre_costs else None
        self._l.ref_download_state(self._h, _fp(planes), _fp(costs), _u32p(views), _u32p(rand6) if rand else None,
                                   _fp(pre) if pre_costs else None)
        self._check()
        return dict(planes=planes, costs=costs, views=views, rand=rand6, pre_costs=pre)

    def upload_state(self, planes=None, costs=None, views=None, rand=None, pre_costs=None):
        keep = []

        def p(a, dt, conv):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return conv(a)
        self._l.ref_upload_state(self._h, p(planes, np.float32, _fp), p(costs, np.float32, _fp), p(views, np.uint32, _u32p),
                                 p(rand, np.uint32, _u32p), p(pre_costs, np.float32, _fp))
        self._check()

    def probe_ncc(self, planes4, view):
        pl = _f32(planes4)
        out = np.empty((self.H, self.W), np.float32)
        self._l.ref_probe_ncc(self._h, _fp(pl), C.c_int(view), _fp(out)); self._check()
        return out

    def probe_coords(self, planes4, view):
        """(H, W, 36, 2): the fetch coordinates of ComputeBilateralNCC's 36 samples (ACMMP.cu:450-476)."""
        pl = _f32(planes4)
        out = np.empty((self.H, self.W, 36, 2), np.float32)
        self._l.ref_probe_coords(self._h, _fp(pl), C.c_int(view), _fp(out)); self._check()
        return out

    def probe_geom(self, planes4, view):
        pl = _f32(planes4)
        out = np.empty((self.H, self.W), np.float32)
        self._l.ref_probe_geom(self._h, _fp(pl), C.c_int(view), _fp(out)); self._check()
        return out

    def probe_warp(self, planes4, view):
        pl = _f32(planes4)
        out = np.empty((self.H, self.W, 4), np.float32)
        self._l.ref_probe_warp(self._h, _fp(pl), C.c_int(view), _fp(out)); self._check()
        return out

    def probe_initcost(self, planes4):
        pl = _f32(planes4)
        out = np.empty((self.H, self.W), np.float32)
        views = np.empty((self.H, self.W), np.uint32)
        self._l.ref_probe_initcost(self._h, _fp(pl), _fp(out), _u32p(views)); self._check()
        return out, views


class quiet_stdout:
    """Route fd 1 away for the duration of a reference call.  The unmodified reference prints to stdout from inside
    the timed calls -- `iteration: %d` per iteration (ACMMP.cu:1542), `depthe range` (ACMMP.cpp:647), and in RunJBU one
    `wrong!` + flush PER NaN PIXEL (ACMMP.cpp:1102-1104) -- tens of MB per benchmark run.  capture=True keeps the text
    (in an anonymous memory file) so that a caller can read the reference's own printed timing."""

    def __init__(self, capture=False):
        self.capture = capture
        self.text = ""

    def __enter__(self):
        C.CDLL(None).fflush(None)
        self._saved = os.dup(1)
        self._fd = os.memfd_create("acmmp_ref_stdout") if self.capture else os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._fd, 1)
        return self

    def __exit__(self, *exc):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        if self.capture:
            # head + tail: RunJBU prints its timing line first and the per-pixel `wrong!` flood after it
            size = os.lseek(self._fd, 0, os.SEEK_END)
            os.lseek(self._fd, 0, os.SEEK_SET)
            self.text = os.read(self._fd, min(size, 1 << 16)).decode("ascii", "replace")
            self.bytes = size
            if size > (1 << 16):
                tail = min(size - (1 << 16), 1 << 16)
                os.lseek(self._fd, size - tail, os.SEEK_SET)
                self.text += os.read(self._fd, tail).decode("ascii", "replace")
        os.close(self._fd)
        return False


def run_jbu(image, coarse_depth, ref_image_id=0, with_ms=False):
    """The reference's RunJBU (ACMMP.cpp:1071-1122); returns the depths.dmb it writes.
    with_ms=True: also the reference's own CUDA-event time of JBU::CudaRun -- kernel + device->host copy, the figure it
    prints as `Total time needed for computation` (ACMMP.cu:1631-1648) -- in milliseconds."""
    l = lib()
    img, dep = _f32(image), _f32(coarse_depth)
    with tempfile.TemporaryDirectory(prefix="acmmp_ref_jbu_") as d:
        os.makedirs(os.path.join(d, "ACMMP"), exist_ok=True)
        with quiet_stdout(capture=True) as q:
            l.ref_run_jbu(_fp(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), _fp(dep), C.c_int(dep.shape[1]),
                          C.c_int(dep.shape[0]), d.encode(), C.c_int(ref_image_id))
        out = read_dmb(os.path.join(d, "ACMMP", "2333_%08d" % ref_image_id, "depths.dmb"))
    if not with_ms:
        return out
    ms = float("nan")
    key = "Total time needed for computation:"
    k = q.text.find(key)
    if k >= 0:
        ms = float(q.text[k + len(key):].split()[0]) * 1e3
    return out, ms


class RefFusion:
    """The reference's SimpleFusionKernel (ACMMP.cu:1664-1814) behind the texture set-up of RunFusionCuda (see ref_harness.cu);
    run(ref, src) returns the points of one reference view filtered on the host in pixel order like ACMMP.cu:2069-2076."""

    def __init__(self, cams, depths, normals, grays, colours=None):
        """colours: optional list of uint8 [h, w, 3] images in OpenCV's B, G, R order (the reference's cv::imread(IMREAD_COLOR))."""
        from acmmp_b200 import Camera
        self._l = lib()
        n = len(cams)
        self.d = [_f32(x) for x in depths]
        self.nm = [_f32(x) for x in normals]
        self.g = [_f32(x) for x in grays]
        self.sizes = [x.shape for x in self.d]
        ws = (C.c_int * n)(*[s[1] for s in self.sizes])
        hs = (C.c_int * n)(*[s[0] for s in self.sizes])
        FP = C.POINTER(C.c_float)
        if colours is None:
            self._h = C.c_void_p(self._l.ref_fusion_create(C.c_int(n), (Camera * n)(*cams), ws, hs, (FP * n)(*[_fp(x) for x in self.d]),
                                                           (FP * n)(*[_fp(x) for x in self.nm]), (FP * n)(*[_fp(x) for x in self.g])))
        else:
            self.c = [np.ascontiguousarray(x, np.uint8) for x in colours]
            assert all(c.shape == s + (3,) for c, s in zip(self.c, self.sizes))
            BP = C.POINTER(C.c_ubyte)
            self._l.ref_fusion_create_bgr.restype = C.c_void_p
            self._h = C.c_void_p(self._l.ref_fusion_create_bgr(C.c_int(n), (Camera * n)(*cams), ws, hs, (FP * n)(*[_fp(x) for x in self.d]),
                                                               (FP * n)(*[_fp(x) for x in self.nm]), (FP * n)(*[_fp(x) for x in self.g]),
                                                               (BP * n)(*[x.ctypes.data_as(BP) for x in self.c])))
        self.kernel_ms = 0.0

    def run(self, ref, src_indices):
        h, w = self.sizes[ref]
        pts = np.empty((h * w, 9), np.float32)
        flags = np.empty(h * w, np.int32)
        src = (C.c_int * len(src_indices))(*[int(s) for s in src_indices])
        self.kernel_ms = float(self._l.ref_fusion_run(self._h, C.c_int(ref), C.c_int(len(src_indices)), src, _fp(pts),
                                                      flags.ctypes.data_as(C.POINTER(C.c_int))))
        if self._l.ref_last_error() != 0:
            raise RuntimeError("reference fusion harness reported a CUDA error")
        return pts[flags != 0], flags.reshape(h, w)

    def close(self):
        if self._h:
            self._l.ref_fusion_destroy(self._h)
            self._h = None
