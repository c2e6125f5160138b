"""The reference's per-view schedule driven through the UNMODIFIED reference -- TEST / BASELINE INFRASTRUCTURE ONLY.

`ReferenceBackend` plugs into acmmp_b200.pipeline.run_view like the product's B200Backend, but every stage runs in
oracle/_ref/libacmmp_ref.so (reference ACMMP.cu + ACMMP.cpp compiled for sm_100): one new `ACMMP` object per
ProcessProblem call (main.cpp:73-210), state handed over through .dmb files in a temporary folder, RunJBU between
levels (main.cpp:212-238).  Only tests/, __graft_entry__.smoke() and bench.py's reference arm import this module;
nothing under acmmp-spherical_b200/ does (tests/test_cpu_host.py greps for it).
"""
from __future__ import annotations

import time

import numpy as np

from acmmp_b200.pipeline import StageTimes
from oracle.ref_driver import RefACMMP, run_jbu


class ReferenceBackend:
    """The same stages through the UNMODIFIED reference (oracle/_ref/libacmmp_ref.so): one new
    `ACMMP` object per ProcessProblem call, state through .dmb files, exactly like main.cpp.
    t.gpu_ms: CUDA-event time of RunPatchMatch (harness events around the unmodified method) + the reference's own
    CUDA-event figure of JBU::CudaRun (ACMMP.cu:1631-1648); t.wall_s: wall clock of everything incl. host set-up."""
    name = "reference"

    def __init__(self, device=0, seed=1234):
        self.seed = seed
        self.t = StageTimes()
        self.obj = None
        self.prev = None
        self.last = None

    def begin_level(self, level, prev=None, first=True):
        self.prev = prev

    def _finish(self, stage, finest, t_setup):
        obj = self.obj
        t0 = time.perf_counter()
        ms = obj.run_patch_match()
        planes, costs = obj.get_result()
        self.t.wall_s += t_setup + time.perf_counter() - t0
        self.t.gpu_ms += ms
        n_pass = 2 * (2 if stage == "geom" else 3)
        self.t.passes += n_pass
        if finest:
            self.t.pass_ms.setdefault(stage, []).append(ms / n_pass)      # includes init + finalize (~5 %)
        self.last = (planes, costs)
        return planes, costs

    def photometric(self, level, finest=False):
        if self.obj is not None:
            self.obj.close()
        t0 = time.perf_counter()
        if self.prev is None:
            self.obj = RefACMMP(level.images, level.cams, seed=self.seed)
        else:
            planes_prev, costs_prev = self.prev
            fine_depth, jbu_ms = run_jbu(level.images[0], np.ascontiguousarray(planes_prev[..., 3]), with_ms=True)     # RunJBU, ACMMP.cpp:1071
            self.t.gpu_ms += jbu_ms      # the reference's own cudaEvent figure of JBU::CudaRun (kernel + D2H copy)
            self.obj = RefACMMP(level.images, level.cams, seed=self.seed, hierarchy=True,
                                coarse_normals=np.ascontiguousarray(planes_prev[..., :3]),
                                coarse_costs=np.ascontiguousarray(costs_prev), fine_depth=fine_depth)
        return self._finish("photometric", finest, time.perf_counter() - t0)

    def prior(self, level, params, masks, finest=False):
        t0 = time.perf_counter()
        self.obj.set_prior(params, masks)
        return self._finish("prior", finest, time.perf_counter() - t0)

    def geom(self, level, multi, neighbour_depths, finest=False, last=False, device_ptrs=None):
        own_planes, own_costs = self.last
        if self.obj is not None:
            self.obj.close()
        t0 = time.perf_counter()
        dm = [np.ascontiguousarray(own_planes[..., 3])] + list(neighbour_depths)
        self.obj = RefACMMP(level.images, level.cams, seed=self.seed, geom=True, multi_geom=multi, depth_maps=dm,
                            prev_planes=own_planes, prev_costs=own_costs)
        return self._finish("geom", finest, time.perf_counter() - t0)

    def end(self):
        if self.obj is not None:
            self.obj.close()
            self.obj = None
