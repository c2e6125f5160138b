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
(`prior_cpu_s`).  `e2e_driver` (C2 / C4, N = 1) is the prior-INCLUSIVE number: the C++ driver `lib/acmmp_b200 --resident 1
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
    peaks = dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            m = json.load(open(p))
            peaks = dict(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), source="MEASURED_PEAKS.json")
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
def cpu_baseline_port(levels, n_src):
    """The CPU restatement (oracle/acmmp_oracle.c) on a bounded sample of the same workload: one black
    photometric pass at the coarsest level (all source views)."""
    from oracle import cpu_oracle
    L = levels[0]
    H, W = L.images[0].shape
    st = cpu_oracle.random_init(L.images, L.cams, seed=1234, with_costs=False)
    st["costs"] = np.full((H, W), 1.0, np.float32)
    st["views"] = np.zeros((H, W), np.uint32)
    t0 = time.perf_counter()
    cpu_oracle.checkerboard_pass(L.images, L.cams, st, 0, 0)
    dt = time.perf_counter() - t0
    samples = algorithmic_samples_per_pass(W, H, n_src)
    rate = samples / dt
    # samples of one whole view: per level 20 passes (+ ~4 % for the initialisations, ignored)
    per_view = sum(20 * algorithmic_samples_per_pass(l.images[0].shape[1], l.images[0].shape[0], n_src) for l in levels)
    return dict(value=rate / per_view, unit=UNIT, cores=cpu_oracle.num_threads(), kind="port",
                sample=f"one black photometric checkerboard pass at the coarsest level ({W}x{H}, {n_src} source views) "
                       f"in {dt:.1f} s = {rate:.3g} NCC samples/s, extrapolated to the {per_view:.3g} samples of one view")


def roofline_of(cfg, Wf, Hf, pass_ms, peaks):
    """Roofline object of the dominant kernel (k_pass, photometric, finest level).  Unit of work: the NCC sample
    (pixel x hypothesis x source view x tap).  PINHOLE is bound by the texture pipe (1 bilinear fetch per
    sample: peak = the measured R32F fetch rate); SPHERE by the special-function unit (6 MUFU-class operations per sample:
    sqrt, 3 rcp, rsqrt, floor; peak = 16 per clock per SM x 148 SMs x max clock / 6)."""
    photometric_ms = pass_ms.get("photometric")
    if not photometric_ms:
        return None
    n_src = cfg["n_src"]
    alg = algorithmic_samples_per_pass(Wf, Hf, n_src)
    ach = alg / (photometric_ms * 1e-3) / 1e9
    facts = {}
    fp = ROOT / "profiles" / "k_pass_ncu_facts.json"
    if fp.exists():
        facts = json.load(open(fp)).get(cfg["model"], {})
    hbm_bytes = 190 * (Wf * Hf // 2)
    common = {
        "kernel": f"k_pass<{cfg['model'].upper()}, photometric>, finest level", "traffic": facts.get("dram_bytes_per_launch"),
        "algorithmic_samples_per_launch": alg, "executed_samples_per_launch": facts.get("executed_lane_fetches_per_launch"),
        "executed_warp_instructions_per_launch": facts.get("executed_warp_instructions_per_launch"),
        "tex_pipe_pct": facts.get("tex_data_pipe_pct"), "issue_slots_pct": facts.get("issue_slots_pct"),
        "xu_pipe_pct": facts.get("xu_pipe_pct"), "ncu_source": facts.get("source"),
        "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_bytes / (photometric_ms * 1e-3) / 1e9,
                "peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["source"]},
    }
    static = ("traffic / executed samples / instructions / pipe utilisation: static, from the ncu capture named in ncu_source (same kernel "
              "build, same shape)")
    if cfg["model"] == "pinhole":
        peak = peaks["tex_gfetch_s"]
        return {"bound": "tex", "achieved": ach, "peak": peak, "unit": "Gsample/s", "frac": ach / peak, **common,
                "note": f"not HBM bound (SURVEY.md 8(d)): achieved = algorithmic NCC samples (14 hypotheses x {n_src} views x 36 taps x "
                        "pixels/2) / mean launch time (CUDA events, this run); peak = measured R32F bilinear fetch rate of this pool's B200 "
                        f"(profiles/tex_peak_b200.json); {static} -- zero-weight views are legitimately skipped, so executed < algorithmic"}
    # SPHERE.  The reference algorithm's roofline is the special-function unit: 6 MUFU-class operations per sample (sqrt, 3 rcp,
    # rsqrt, floor) = 1.29 ps per sample.  With the zero-weight taps pruned (the default) only ~24 % of the samples are taken
    # and the per-hypothesis work that remains (weights, the tap mask, 14 plane set-ups) makes the kernel ISSUE bound: that is
    # the roofline reported; the SFU figures of the all-taps algorithm ride along.
    sfu_peak = 16 * 148 * peaks["sm_max_mhz"] * 1e6 / 6 / 1e9
    issue_peak = 4 * 148 * peaks["sm_max_mhz"] * 1e6 / 1e9
    inst = facts.get("executed_warp_instructions_per_launch")
    issue = inst / (photometric_ms * 1e-3) / 1e9 if inst else None
    return {"bound": "issue", "achieved": issue, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": issue / issue_peak if issue else None,
            **common,
            "sfu": {"peak_gsample_s": sfu_peak, "algorithmic_gsample_s": ach, "algorithmic_over_peak": ach / sfu_peak,
                    "executed_gsample_s": (common["executed_samples_per_launch"] or 0) / (photometric_ms * 1e-3) / 1e9,
                    "peak_how": f"16 special-function results per clock per SM x 148 SMs x {peaks['sm_max_mhz']:.0f} MHz / 6 per sample "
                                "(nominal; no measured SFU peak)",
                    "note": "algorithmic_over_peak > 1 is work removal, not a faster unit: taps whose FP32 bilateral weight is below 2^-24 "
                            "of the window sum are not sampled.  The all-taps kernel against this roofline: sphere_tap_pruning.sfu_frac_all_taps"},
            "note": "not HBM bound (SURVEY.md 8(d)).  bound = instruction issue, the busiest unit of the default (tap-pruned) kernel in ncu: "
                    "achieved = executed warp instructions per launch (static, ncu) / mean launch time (CUDA events, this run); peak = 4 "
                    f"schedulers x 148 SMs x {peaks['sm_max_mhz']:.0f} MHz.  {static}"}


def driver_leg(cfg, scene, ids, device):
    """The product's own end to end: the C++ driver (`lib/acmmp_b200 --resident 1 --gpu-prior 1`) on a dense folder of the
    step's 11 views (each using the other ten as source views): image files in, .dmb files out, planar prior INCLUDED
    (support points + plane fit + rasteriser on the device, Delaunay on a host thread).  Returns seconds per view."""
    import shutil
    import tempfile
    from acmmp_b200 import synth
    driver = ROOT / "acmmp-spherical_b200" / "lib" / "acmmp_b200"
    if not driver.exists():
        return None
    sub = synth.Scene(scene.model, [scene.images[i] for i in ids], [scene.cams[i] for i in ids], [scene.depths_gt[i] for i in ids],
                      [(k, [j for j in range(len(ids)) if j != k][: cfg["n_src"]]) for k in range(len(ids))],
                      [scene.Rs[i] for i in ids], [scene.ts[i] for i in ids], [scene.Ks[i] for i in ids], scene.quads)
    # images + four .dmb files per view + the fused PLY: ~5 GB at C2; RAM-backed when there is room for it
    shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (12 << 30)
    tmp = tempfile.mkdtemp(prefix="acmmp_bench_driver_", dir="/dev/shm" if shm_ok else None)
    try:
        synth.write_dense_folder(sub, tmp, pgm=True)
        t0 = time.perf_counter()
        r = subprocess.run([str(driver), tmp, "--seed", "1234", "--resident", "1", "--gpu-prior", "1", "--device", str(device)],
                           capture_output=True, text=True, timeout=900)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": (r.stdout[-300:] + r.stderr[-300:])}
        d = json.loads(r.stdout.strip().splitlines()[-1])
        n = d["views"]
        trace = [l for l in r.stderr.splitlines() if l.startswith("[acmmp trace]")]          # ACMMP_TRACE=1 in the environment
        return {**({"trace": trace} if trace else {}),"views": n, "src_views": cfg["n_src"], "s_per_view": d["wall_s"] / n, "value": n / d["wall_s"], "unit": UNIT,
                "kernel_s_per_view": d["kernel_ms"] / 1e3 / n, "fusion_s": d.get("fusion_s"), "fusion_points": d.get("fusion_points"),
                "s_per_view_after_cuda_startup": (d["wall_s"] - d.get("setup_s", 0.0)) / n, "prior_s_per_view": d["prior_cpu_s"] / n, "process_wall_s": wall,
                "breakdown_s": {k: d[k] for k in ("setup_s", "load_s", "views_s", "ctx_s", "upload_s", "run_s", "support_s", "prior_dev_s", "export_s", "output_s", "join_s", "sweep1_s", "geom_s") if k in d},
                "what": "lib/acmmp_b200 <dense_folder> --resident 1 --gpu-prior 1: wall clock inside the process from pair.txt to the "
                        "last .dmb file, all levels and stages of every view, planar prior included (setup_s = CUDA start-up of the "
                        "process, a per-process constant of 1 - 2.5 s on these boxes, is part of s_per_view)"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------
def run_c3(a, cfg, rank, local_rank, world, use_dist, json_fd):
    """C3: the whole 64-view scene per step, views sharded over the ranks (acmmp_b200/scene.py)."""
    import torch
    import torch.distributed as dist
    from acmmp_b200 import scene as sc, shard
    n_views = cfg["scene_views"]
    # every rank renders the views v = rank (mod world) and shares them through /dev/shm: each rank needs every view as a
    # source view of its own reference views (ring, 10 nearest)
    from acmmp_b200 import synth
    full = make_scene(cfg, [])                                  # cameras + pairs, nothing rendered
    share = Path("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") / f"acmmp_bench_c3_{os.environ.get('MASTER_PORT', '0')}"
    share.mkdir(exist_ok=True)
    mine = shard.views_of(rank, n_views, world)
    for v in mine:
        img, dep = synth._render(full.quads, synth.MODEL_PINHOLE, full.Rs[v], full.ts[v], cfg["width"], cfg["height"], K=full.Ks[v])
        np.save(share / f"img_{v}.npy", img)
        np.save(share / f"dep_{v}.npy", dep)
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    full.images = [np.load(share / f"img_{v}.npy") for v in range(n_views)]
    gt_mine = {v: np.load(share / f"dep_{v}.npy") for v in mine}
    if use_dist:
        dist.barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(share, ignore_errors=True)
    levels = sc.SceneLevels.build(full, pin=pin)
    dev = torch.device("cuda", local_rank)

    def alloc(shape):
        t = torch.empty(shape, dtype=torch.float32, device=dev)
        return t, t.data_ptr()

    def all_gather(table):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if use_dist:
            dist.all_gather_into_tensor(table.all.view(-1), table.mine.view(-1))       # NCCL over NVLink: the path's only collective
        else:
            table.all[0].copy_(table.mine)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    results = {}

    def on_result(v, planes, costs):
        gt = gt_mine[v]
        results[v] = float((np.abs(planes[..., 3] - gt) / gt <= 0.01)[8:-8, 8:-8].mean())

    workers, tables = {}, []          # one context per owned view and the two map tables live through all steps (warm pools)

    def step():
        return sc.run_scene(levels, full.pairs, rank, world, lambda: sc.GpuWorker(local_rank, 1234), alloc, all_gather, on_result=on_result,
                            workers=workers, tables=tables)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    tot = sc.SceneTimes()
    for _ in range(a.steps):
        t = step()
        tot.gpu_ms += t.gpu_ms; tot.exchange_ms += t.exchange_ms; tot.exchange_bytes += t.exchange_bytes; tot.prior_host_s += t.prior_host_s
        tot.h2d_bytes += t.h2d_bytes; tot.d2h_bytes += t.d2h_bytes; tot.launches += t.launches; tot.passes += t.passes
        for k, v in t.pass_ms.items():
            tot.pass_ms.setdefault(k, []).extend(v)
    barrier()
    wall_step = (time.perf_counter() - t0) / a.steps
    clocks = sampler.stop()
    gpu_ms_step = (tot.gpu_ms + tot.exchange_ms) / a.steps
    quality = float(np.mean(list(results.values()))) if results else float("nan")
    if use_dist:
        tt = torch.tensor([gpu_ms_step, wall_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gpu_ms_step, wall_step = float(tt[0]), float(tt[1])
        qq = torch.tensor([quality * len(mine), float(len(mine)), float(tot.launches), float(tot.h2d_bytes), float(tot.d2h_bytes)], dtype=torch.float64, device="cuda")
        dist.all_reduce(qq, op=dist.ReduceOp.SUM)
        quality, launches, h2d, d2h = float(qq[0] / qq[1]), int(qq[2]), int(qq[3]), int(qq[4])
        nranks = dist.get_world_size()
    else:
        launches, h2d, d2h, nranks = tot.launches, tot.h2d_bytes, tot.d2h_bytes, 1
    if rank == 0:
        peaks = load_peaks()
        Hf, Wf = levels.images[-1][0].shape
        pass_ms = {k: float(np.mean(v)) for k, v in tot.pass_ms.items()}
        line = {
            "metric": cfg["metric"], "value": n_views / (gpu_ms_step / 1e3), "unit": UNIT, "n_gpus": nranks, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": gpu_ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "views": n_views, "views_per_rank": len(mine), "n_src": cfg["n_src"],
                       "levels": [list(l[0].shape[::-1]) for l in levels.images],
                       "l2": "working set per pass (planes 109 MB + 11 x 27 MB images) exceeds the 126 MB L2",
                       "timing": "value: sum of CUDA-event kernel times + all-gathers, max over ranks; e2e: wall clock incl. H2D of "
                                 "every image, the host Delaunay of the planar prior (worker threads) and D2H of every final map",
                       "exchange": "2 NCCL all-gathers per level of [views_per_rank, H, W] float32 depth maps",
                       "nccl_nranks": nranks},
            "clocks": clocks,
            "e2e": {"value": n_views / wall_step, "unit": UNIT, "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps,
                    "ms_per_step": wall_step * 1e3},
            "gpu_launches": launches, "ms_per_checkerboard_pass": pass_ms,
            "prior_host_s_not_hidden_per_step": tot.prior_host_s / a.steps,
            "depth_within_1pct_of_ground_truth": quality,
            "nccl_allgather": {"ms_per_step": tot.exchange_ms / a.steps, "bytes_per_step": tot.exchange_bytes // a.steps},
        }
        rl = roofline_of(cfg, Wf, Hf, pass_ms, peaks)
        if rl:
            line["roofline"] = rl
        os.write(json_fd, (json.dumps(line) + "\n").encode())
        print(f"[bench] C3: NCCL communicator nranks={nranks}, {n_views} views, {len(mine)} per rank", file=sys.stderr)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-driver-leg", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    cfg = CONFIGS[a.config]
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
    try:
        if a.config == "C3" and a.impl == "b200":
            return run_c3(a, cfg, rank, local_rank, world, use_dist, json_fd)
        return run_views(a, cfg, rank, local_rank, world, use_dist, json_fd)
    finally:
        if use_dist:
            dist.barrier()
            dist.destroy_process_group()


def run_views(a, cfg, rank, local_rank, world, use_dist, json_fd):
    """C2 / C4 (and the reference arm of every config): one reference view per step per GPU."""
    import torch
    if use_dist:
        import torch.distributed as dist
    from acmmp_b200 import pipeline
    if a.config == "C3":                       # reference arm: a one-view sample of the scene
        cfg = dict(cfg, scene_views=16, ring=False)
    n_src = cfg["n_src"]
    scene, levels, ids = make_levels(cfg, rank if a.impl == "b200" else 0)
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
        prior_cache[k] = (pin(pp), pin(mm))

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
    nranks = 1
    if use_dist:
        tt = torch.tensor([gpu_ms_step, wall_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gpu_ms_step, wall_step = float(tt[0]), float(tt[1])
        nranks = dist.get_world_size()
    n_gpus = world if use_dist else 1

    # quality of what was computed (not part of the metric): finest level vs ground truth
    planes, costs = result
    gt = scene.depths_gt[ids[0]]
    if gt.shape != planes.shape[:2]:
        import cv2
        gt = cv2.resize(gt, (planes.shape[1], planes.shape[0]), interpolation=cv2.INTER_NEAREST)
    within = float((np.abs(planes[..., 3] - gt) / gt <= 0.01)[8:-8, 8:-8].mean())

    if rank == 0:
        peaks = load_peaks()
        Hf, Wf = levels[-1].images[0].shape
        pass_ms = {k: float(np.mean(v)) for k, v in tot.pass_ms.items()}
        line = {
            "metric": cfg["metric"], "value": n_gpus / (gpu_ms_step / 1e3), "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": gpu_ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"],
                       "levels": [list(l.images[0].shape[::-1]) for l in levels], "n_src": n_src,
                       "l2": "working set per pass (planes + source images of the finest level, > 300 MB) exceeds the 126 MB L2",
                       "timing": "value: sum of CUDA-event kernel times; e2e: wall clock of the host-buffer API calls",
                       "neighbour_depths": "other ranks' maps via NCCL all-gather where available, else rendered stand-ins",
                       "cpu_prior_stage": "run once in warm-up, reused (out of scope, timed separately); e2e_driver includes it",
                       "nccl_nranks": nranks},
            "clocks": clocks,
            "e2e": {"value": n_gpus / wall_step, "unit": UNIT, "h2d_bytes_per_step": (tot.h2d_bytes + exch.h2d) // a.steps,
                    "d2h_bytes_per_step": tot.d2h_bytes // a.steps, "ms_per_step": wall_step * 1e3},
            "gpu_launches": tot.launches,
            "ms_per_checkerboard_pass": pass_ms,
            "prior_cpu_s_per_view": prior_cpu_s,
            "depth_within_1pct_of_ground_truth": within,
            "nccl_allgather": {"ms_per_step": exch.ms / a.steps, "bytes_per_step": exch.bytes // a.steps},
        }
        rl = roofline_of(cfg, Wf, Hf, pass_ms, peaks)
        if rl:
            line["roofline"] = rl
        if a.impl == "reference":
            line["impl"] = "reference"
            line["config"]["timing"] = ("value: CUDA events around the reference's RunPatchMatch (kernels + its device-wide syncs + its "
                                        "device->host copy of the result, ACMMP.cu:1506-1556) + the reference's own CUDA-event figure of "
                                        "JBU::CudaRun; e2e: wall clock of one ACMMP object per stage with .dmb hand-over, like main.cpp")
            if a.config == "C3":
                line["config"]["sample"] = "one reference view of the scene per step (the reference is single-GPU and sequential over views)"
            line["cpu_baseline"] = {"value": line["value"], "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": "the reference's own CUDA build (sm_100) of the same step on this B200: the "
                                              "reference has no CPU PatchMatch path (north_star)"}
            line["e2e"]["h2d_bytes_per_step"] = 0
            line["e2e"]["d2h_bytes_per_step"] = 0
        else:
            if not a.no_cpu_baseline:
                try:
                    line["cpu_baseline"] = cpu_baseline_port(levels, n_src)
                except Exception as e:      # the checker is optional for the number, never for the product
                    line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
            if cfg["model"] == "sphere":
                # the same photometric stage at the finest level with EVERY window tap sampled (acmmp_set_sphere_tap_pruning 0):
                # the default skips taps whose bilateral weight is below 2^-24 of the window's weight sum
                try:
                    ctx.set_sphere_tap_pruning(0.0)
                    ctx.reset_modes()
                    ctx.set_views(levels[-1].images, levels[-1].cams)
                    ctx.run_patch_match(download=False)
                    tm = ctx.timings()
                    unpruned = tm["pass_sum_ms"] / max(tm["n_pass"], 1)
                    ctx.set_sphere_tap_pruning(2.0 ** -24)
                    line["sphere_tap_pruning"] = {
                        "relative_weight_threshold": 2.0 ** -24, "photometric_pass_ms_all_taps": unpruned,
                        "photometric_pass_ms_default": pass_ms.get("photometric"),
                        "sfu_frac_all_taps": (algorithmic_samples_per_pass(Wf, Hf, n_src)
                                              / (unpruned * 1e-3) / 1e9) / line["roofline"]["sfu"]["peak_gsample_s"] if line.get("roofline") else None,
                        "note": "default: taps below the threshold are not sampled (within the reference's own one-ulp sensitivity, "
                                "tests/test_gpu_parity.py::test_sphere_tap_pruning_*); sfu_frac_all_taps = the all-taps kernel's algorithmic "
                                "samples per second over the 1.29 ps/sample special-function roofline"}
                except Exception as e:
                    line["sphere_tap_pruning"] = {"error": str(e)[:200]}
            if not a.no_driver_leg and n_gpus == 1 and a.config in ("C2", "C4"):
                try:
                    ctx.close()
                    line["e2e_driver"] = driver_leg(cfg, scene, ids, local_rank)
                except Exception as e:
                    line["e2e_driver"] = {"error": str(e)[:300]}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
        if use_dist:
            print(f"[bench] NCCL communicator nranks={nranks}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
