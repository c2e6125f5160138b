#!/usr/bin/env python
"""bench.py -- headline benchmark of the PatchMatch hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C2|C3|C4]

Workloads (BASELINE.json configs):
  C2 (default; configs[1]): 3200x2130 reference view, 10 source views, the FULL multi-scale ACMMP schedule for one
      reference view = 3 pyramid levels x (photometric, planar prior, geometric consistency x 2) = 30 checkerboard
      iterations + 12 initialisations + 2 JBU (reference main.cpp:417-476).  A step = ONE reference view per GPU.
  C4 (configs[3]): equirectangular 4096x2048 panoramas (processed at 3200x1600: the reference caps the longer side at
      3200, ACMMP.h:36), 8 source views, the same schedule through the fork's spherical projection.  A step = one view.
  C3 (configs[2]): a 64-view 3200x2130 scene, 10 source views each, reference views dealt round-robin to the GPUs.  A step
      = the WHOLE scene, level by level like the reference's pipeline loop (acmmp_b200/scene.py): every rank runs the
      photometric + prior stage of the views it owns, the ranks all-gather the depth maps over NCCL, geometric round 0,
      all-gather, geometric round 1.  Strong scaling: the scene is fixed, N varies.

  value  : depth maps / s with inputs resident: (views per step over all ranks) / (sum of the CUDA-event times of every
           kernel of the step + the all-gathers, max over ranks)
  e2e    : the same metric through the host-buffer C ABI, wall clock bracketed by barrier + synchronize, max over ranks.
           Inside the timed region: H2D (pinned) of every level's images, of the prior and of the stand-in neighbour
           depth maps; D2H of the results the host needs; the stages in between hand their state over on the device.
           (C3: also the host part of the planar prior -- the Delaunay triangulation -- on worker threads.)
  N > 1  : one process per GPU (torchrun).  C2 / C4: rank r owns reference view r of the scene (weak scaling); after the
           prior stage and after the first geometric stage the ranks all-gather their depth maps over NCCL and use them as
           neighbour depth maps wherever a source view is another rank's reference view (the only collective of the path)
  --impl reference : the UNMODIFIED reference kernels + host set-up (oracle/_ref/libacmmp_ref.so, its own CUDA build for
           sm_100) through the same schedule on one B200, one view per step (C3: a one-view sample of the scene).  The
           reference has no CPU PatchMatch; BASELINE.json:north_star names this build as the timed baseline.

C2 / C4: the CPU planar-prior stage (Delaunay etc., reference ACMMP.cpp:904-1011) is out of scope for both arms: it runs
once per level during warm-up and its output is re-used in the timed steps; its time is reported separately
(`prior_cpu_s`).  `e2e_driver` (C2, N = 1) is the prior-INCLUSIVE number: the C++ driver `lib/acmmp_b200 --resident 1
--gpu-prior 1` on an 11-view dense folder of the same shape, seconds per view of its own wall clock (image files in, .dmb
files out).  Data are synthetic (acmmp_b200/synth.py), neighbour depth maps that no rank computes are rendered stand-ins.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "acmmp-spherical_b200"))
sys.path.insert(0, str(ROOT))

import numpy as np

UNIT = "depth maps/s"
CONFIGS = {
    "C2": dict(model="pinhole", width=3200, height=2130, focal=2800.0, n_src=10, scene_views=16, seed=2,
               metric="depth maps/s @3200x2130, 10 src views (full ACMMP)",
               workload="C2: 3200x2130 reference view, 10 source views, 3 pyramid levels x (photometric + planar prior + "
                        "2 x geometric consistency), one reference view per step per GPU"),
    "C4": dict(model="sphere", width=4096, height=2048, n_src=8, scene_views=9, seed=4,
               metric="depth maps/s @4096x2048 equirectangular (processed at 3200x1600), 8 src views (full ACMMP)",
               workload="C4: 4096x2048 equirectangular reference view (processed at 3200x1600, the reference's 3200 px cap), "
                        "8 source views, 3 pyramid levels x (photometric + planar prior + 2 x geometric consistency), one "
                        "reference view per step per GPU"),
    "C3": dict(model="pinhole", width=3200, height=2130, focal=2800.0, n_src=10, scene_views=64, seed=3, ring=True,
               metric="depth maps/s, 64-view 3200x2130 scene, 10 src views (full ACMMP), views sharded per GPU",
               workload="C3: 64 reference views 3200x2130 on a ring, 10 source views each, whole scene per step, level by level "
                        "(photometric + prior, all-gather, geometric 0, all-gather, geometric 1), views dealt round-robin to the GPUs"),
}


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, smmax, power, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smmax) if smmax else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def algorithmic_samples_per_pass(w, h, nsrc):
    """14 hypotheses x (N-1) views x 36 taps per pixel visit, half the pixels per pass."""
    return 14 * nsrc * 36 * (w * h // 2)


def load_peaks():
    peaks = dict(hbm_gbs=6650.0, source="fallback")
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            peaks = dict(hbm_gbs=float(json.load(open(p))["hbm_gbs"]), source="MEASURED_PEAKS.json")
        except Exception:
            pass
    t = ROOT / "profiles" / "tex_peak_b200.json"
    peaks["tex_gfetch_s"] = float(json.load(open(t))["r32f_bilinear_gfetch_per_s"]) if t.exists() else 1139.0
    return peaks


# ------------------------------------------------------------------------------------------
def make_scene(cfg, render_ids):
    from acmmp_b200 import synth
    if cfg["model"] == "pinhole":
        return synth.make_pinhole_scene(n_views=cfg["scene_views"], width=cfg["width"], height=cfg["height"], focal=cfg["focal"],
                                        seed=cfg["seed"], n_src=cfg["n_src"], ring=cfg.get("ring", False), render_ids=render_ids)
    return synth.make_sphere_scene(n_views=cfg["scene_views"], width=cfg["width"], height=cfg["height"], seed=cfg["seed"],
                                   n_src=cfg["n_src"], render_ids=render_ids)


def pin(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def make_levels(cfg, rank):
    from acmmp_b200 import pipeline
    scene = make_scene(cfg, [rank])
    levels = pipeline.build_levels(scene, rank)
    ids = [rank] + list(scene.pairs[rank][1])
    # the step's host inputs live in pinned memory (contract: H2D from pinned host memory)
    for L in levels:
        L.images = [pin(im) for im in L.images]
        L.neighbour_depths = [pin(d) for d in L.neighbour_depths]
    return scene, levels, ids


class Exchange:
    """Depth-map exchange between the stages (north_star: the only collective).  With N ranks every rank
    exports its view's depth map on the device, the maps are all-gathered over NCCL, and every source view that
    is some rank's reference view reads that rank's fresh map straight from the gathered device buffer; the other
    source views use the step's stand-in maps (inputs: copied host -> device inside the timed region)."""

    def __init__(self, world, rank, local_rank, ids):
        self.world, self.rank, self.ids = world, rank, ids
        self.bytes = 0
        self.ms = 0.0
        self.h2d = 0
        if world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            self.dev = torch.device("cuda", local_rank)
            self.keep = []

    def neighbour_depths(self, li, level, backend):
        if self.world == 1:
            return list(level.neighbour_depths), None
        from acmmp_b200 import shard
        torch, dist = self.torch, self.dist
        if not hasattr(self, "ex"):
            self.ex = shard.DepthExchange(dist, self.world)
        H, W = level.images[0].shape
        mine = torch.empty((H, W), dtype=torch.float32, device=self.dev)
        backend.own_depth_to(mine.data_ptr())                      # device -> device, waits for the stage
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        gathered = self.ex.all_gather(mine)                        # NCCL over NVLink: the path's only collective
        t1.record()
        # rank r owns reference view r (one view per rank per step): round 0 of the plan
        plan = shard.neighbour_sources(self.ids[1:], 0, self.world)
        standins, ptrs = [], []
        for k, (kind, who) in enumerate(plan):
            d = level.neighbour_depths[k]
            if kind == "gathered" and d.shape == (H, W):
                ptrs.append((gathered[who].data_ptr(), W, H))
            else:
                t = torch.from_numpy(d).to(self.dev, non_blocking=True)      # pinned host -> device (step input)
                standins.append(t)
                self.h2d += d.nbytes
                ptrs.append((t.data_ptr(), d.shape[1], d.shape[0]))
        torch.cuda.synchronize()
        self.ms += t0.elapsed_time(t1)
        self.bytes = self.ex.bytes
        self.keep = [mine, gathered, standins]                     # alive until the next exchange
        return None, ptrs


def run_step(levels, backend, prior_cache, exch):
    """One reference view through the whole schedule, with the depth exchange before each geometric stage."""
    from acmmp_b200.pipeline import run_view
    return run_view(levels, backend, prior_cache, exchange=exch)


# ------------------------------------------------------------------------------------------
def cpu_baseline_port(levels):
    """The CPU restatement (oracle/acmmp_oracle.c) on a bounded sample of the same workload: one black
    photometric pass at the coarsest level (all source views)."""
    from oracle import cpu_oracle
    L = levels[0]
    H, W = L.images[0].shape
    nsrc = len(L.images) - 1
    st = cpu_oracle.random_init(L.images, L.cams, seed=1234, with_costs=False)
    st["costs"] = np.full((H, W), 1.0, np.float32)
    st["views"] = np.zeros((H, W), np.uint32)
    t0 = time.perf_counter()
    cpu_oracle.checkerboard_pass(L.images, L.cams, st, 0, 0)
    dt = time.perf_counter() - t0
    samples = algorithmic_samples_per_pass(W, H, nsrc)
    rate = samples / dt
    # samples of one whole view: per level 20 passes (+ ~4 % for the initialisations, ignored)
    per_view = sum(20 * algorithmic_samples_per_pass(l.images[0].shape[1], l.images[0].shape[0], nsrc) for l in levels)
    return dict(value=rate / per_view, unit=UNIT, cores=cpu_oracle.num_threads(), kind="port",
                sample=f"one black photometric checkerboard pass at the coarsest level ({W}x{H}, {nsrc} source views) "
                       f"in {dt:.1f} s = {rate:.3g} NCC samples/s, extrapolated to the {per_view:.3g} samples of one view")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank, local_rank, world = dist_env()
    if world != a.gpus and world > 1:
        a.gpus = world
    # fd 1 carries the ONE JSON line and nothing else.  The b200 arm's other prints go to stderr; the reference arm's
    # go to /dev/null: the unmodified reference prints from inside its timed calls (`iteration:` ACMMP.cu:1542, `depthe
    # range` ACMMP.cpp:647, one flushed `wrong!` per NaN pixel in RunJBU ACMMP.cpp:1102-1104) -- tens of MB per run, which
    # in round 1 pushed the driver's capture past its limit and cost the scaling run its evidence.
    json_fd = os.dup(1)
    if a.impl == "reference":
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        os.close(devnull)
    else:
        os.dup2(2, 1)

    if a.impl == "reference" and rank != 0:
        return 0            # the reference is single-GPU: rank 0 alone runs and prints it

    import torch
    if not torch.cuda.is_available():
        os.write(json_fd, (json.dumps({"impl": a.impl, "error": "no CUDA device: this benchmark has no CPU fallback"}) + "\n").encode())
        return 1
    torch.cuda.set_device(local_rank)
    use_dist = world > 1 and a.impl == "b200"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from acmmp_b200 import pipeline
    scene, levels, ids = make_levels(rank if a.impl == "b200" else 0)
    exch = Exchange(world if use_dist else 1, rank, local_rank, ids)
    prior_cache = {}

    ctx = None
    if a.impl == "b200":
        from acmmp_b200 import Context
        ctx = Context(local_rank)        # one context (and its buffer pool) serves every view this rank processes

    def new_backend():
        if a.impl == "reference":
            from oracle.ref_pipeline import ReferenceBackend      # the checker / baseline: never on the b200 arm
            return ReferenceBackend(local_rank, seed=1234)
        return pipeline.B200Backend(local_rank, seed=1234, ctx=ctx)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()

    prior_cpu_s = 0.0
    for _ in range(a.warmup):
        b = new_backend()
        run_step(levels, b, prior_cache, exch)
        b.end()
        prior_cpu_s += b.t.prior_cpu_s          # only the first warm-up step computes it

    # the prior (computed once in warm-up) is a step input like the images: pinned host memory
    for k, (pp, mm) in list(prior_cache.items()):
        prior_cache[k] = (torch.from_numpy(np.ascontiguousarray(pp)).pin_memory().numpy(),
                          torch.from_numpy(np.ascontiguousarray(mm)).pin_memory().numpy())

    sampler = ClockSampler(local_rank)
    sampler.start()
    exch.ms, exch.bytes, exch.h2d = 0.0, 0, 0
    if hasattr(exch, "ex"):
        exch.ex.bytes = 0
    barrier()
    t0 = time.perf_counter()
    tot = pipeline.StageTimes()
    result = None
    for _ in range(a.steps):
        b = new_backend()
        result = run_step(levels, b, prior_cache, exch)
        b.end()
        tot.gpu_ms += b.t.gpu_ms; tot.wall_s += b.t.wall_s; tot.h2d_bytes += b.t.h2d_bytes; tot.d2h_bytes += b.t.d2h_bytes
        tot.launches += b.t.launches; tot.passes += b.t.passes
        for k, v in b.t.pass_ms.items():
            tot.pass_ms.setdefault(k, []).extend(v)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    gpu_ms_step = (tot.gpu_ms + exch.ms) / a.steps
    wall_step = wall / a.steps
    if use_dist:
        tt = torch.tensor([gpu_ms_step, wall_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gpu_ms_step, wall_step = float(tt[0]), float(tt[1])
    n_gpus = world if use_dist else 1

    # quality of what was computed (not part of the metric): finest level vs ground truth
    planes, costs = result
    gt = scene.depths_gt[ids[0]]
    within = float((np.abs(planes[..., 3] - gt) / gt <= 0.01)[8:-8, 8:-8].mean())

    if rank == 0:
        peaks = load_peaks()
        Hf, Wf = levels[-1].images[0].shape
        alg = algorithmic_samples_per_pass(Wf, Hf, N_SRC)
        pass_ms = {k: float(np.mean(v)) for k, v in tot.pass_ms.items()}
        photometric_ms = pass_ms.get("photometric")
        line = {
            "metric": METRIC, "value": n_gpus / (gpu_ms_step / 1e3), "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": gpu_ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 3200x2130 reference view, 10 source views, 3 pyramid levels x (photometric + planar prior + "
                                   "2 x geometric consistency), one reference view per step per GPU",
                       "levels": [list(l.images[0].shape[::-1]) for l in levels], "n_src": N_SRC,
                       "l2": "working set per pass (planes 109 MB + 11 x 27 MB images) exceeds the 126 MB L2",
                       "timing": "value: sum of CUDA-event kernel times; e2e: wall clock of the host-buffer API calls",
                       "neighbour_depths": "other ranks' maps via NCCL all-gather where available, else rendered stand-ins",
                       "cpu_prior_stage": "run once in warm-up, reused (out of scope, timed separately)"},
            "clocks": clocks,
            "e2e": {"value": n_gpus / wall_step, "unit": UNIT, "h2d_bytes_per_step": (tot.h2d_bytes + exch.h2d) // a.steps,
                    "d2h_bytes_per_step": tot.d2h_bytes // a.steps, "ms_per_step": wall_step * 1e3},
            "gpu_launches": tot.launches,
            "ms_per_checkerboard_pass": pass_ms,
            "prior_cpu_s_per_view": prior_cpu_s,
            "depth_within_1pct_of_ground_truth": within,
            "nccl_allgather": {"ms_per_step": exch.ms / a.steps, "bytes_per_step": exch.bytes // a.steps},
        }
        if photometric_ms:
            ach = alg / (photometric_ms * 1e-3) / 1e9
            traffic = None
            tp = ROOT / "profiles" / "pass_traffic.json"
            if tp.exists():
                traffic = json.load(open(tp)).get("dram_bytes_per_launch_full_res")
            line["roofline"] = {
                "bound": "tex", "kernel": "k_pass (photometric, finest level)", "achieved": ach, "peak": peaks["tex_gfetch_s"],
                "unit": "Gsample/s", "frac": ach / peaks["tex_gfetch_s"], "traffic": traffic,
                "algorithmic_samples_per_launch": alg,
                "note": "the pass is texture/FP32 bound, not HBM bound (SURVEY.md 8(d)): achieved = algorithmic NCC samples "
                        "(14 hypotheses x 10 views x 36 taps x pixels/2) / mean launch time; peak = measured R32F bilinear fetch "
                        "rate of this pool's B200 (profiles/tex_peak_b200.json)",
                "hbm": {"algorithmic_bytes_per_launch": 190 * (Wf * Hf // 2), "achieved_gbs": 190 * (Wf * Hf // 2) / (photometric_ms * 1e-3) / 1e9,
                        "peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["source"]},
            }
        if a.impl == "reference":
            line["impl"] = "reference"
            line["cpu_baseline"] = {"value": line["value"], "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": "the reference's own CUDA build (sm_100) of the same step on this B200: the "
                                              "reference has no CPU PatchMatch path (north_star)"}
            line["e2e"]["h2d_bytes_per_step"] = 0
            line["e2e"]["d2h_bytes_per_step"] = 0
        elif not a.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_port(levels)
            except Exception as e:      # the checker is optional for the number, never for the product
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
