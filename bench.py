#!/usr/bin/env python
"""bench.py -- frames/s of the ELAS stereo hot path (disparity + point cloud) at 1242x375 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]           one JSON line (rank 0)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                          the reference's own serial ELAS on the host cores

Workload (BASELINE.json configs[1]): a batch of synthetic rectified KITTI-shape pairs (1242x375, textured planes of
known disparity, csrc/synth.cpp) per GPU, frames sharded over GPUs with no collective (weak scaling: the per-GPU
batch is fixed).  One "step" = one pass of the whole path (descriptor -> support matches -> host Delaunay -> planes /
grid / raster -> dense matching -> L/R check -> speckle removal -> gap interpolation -> adaptive mean -> median ->
u8 conversion + reprojectTo3D) over the batch, through the C-ABI of include/elas_b200.h.

  value      frames/s with the inputs resident in HBM when the timed region starts (results stay in HBM)
  e2e        frames/s through svb_batch_run_host(): pinned HOST buffers in, host buffers out (disparity + double3
             point cloud, like generatePointCloud returns), H2D and D2H copies inside the timed region
  roofline   the kernel with the largest share of the step, its algorithmic bytes over its mean
             launch duration measured with CUDA events in the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the reference's serial ELAS (oracle/_ref, Makefile flags) + the C restatement of projectParallel (oracle/project_port.c),
             frame-parallel over the host cores, on a bounded sample of the same frames

torch is used for process-group plumbing only (barrier, max over ranks); every kernel is the library's own.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 1242, 375
N = W * H
METRIC = "frames/s at 1242x375 (disparity + point cloud)"
UNIT = "frames/s"
# config.workload is the same string in both arms (--impl b200 / --impl reference): it names the workload, everything that differs
# between the arms (batch per step, sample size) sits in the other config keys
WORKLOAD = ("synthetic KITTI-shape 1242x375 stereo pairs (BASELINE configs[1]), pipeline preset (MIDDLEBURY + postprocess_only_left + "
            "filter_adaptive_mean), disparity + point cloud")

# Q / XR / XT of data/calibration/kitti_2011_09_26.yml at 1242x375 (tests/golden/golden_meta.json, cv2.stereoRectify)
Q_KITTI = [[1.0, 0.0, 0.0, -738.7995529174805], [0.0, 1.0, 0.0, -254.75721931457520], [0.0, 0.0, 0.0, 1027.8551581758902],
           [0.0, 0.0, 1.8616160699568378, -0.0]]


def load_pkg():
    from __graft_entry__ import load_package

    return load_package().binding


def synth_lib():
    """The synthetic-input generator as its own tiny host library (csrc/synth.cpp alone): the CPU legs generate their inputs
    with it, so a process that times the reference never maps the product's libelas_b200.so."""
    import ctypes as C
    import glob

    return C.CDLL(glob.glob(os.path.join(ROOT, "low-cost*", "lib", "libsvb_synth.so"))[0])


# ---- algorithmic HBM bytes per frame of each stage (DESIGN.md "Kernels and rooflines") ---------------
def stage_bytes(p, d):
    both = 1 if p.postprocess_only_left else 2
    lattice_rows = d["ch"] - 1
    return {
        "descriptor": 2 * N + 2 * 16 * N,                      # read both images, write both descriptor images
        "support_match": 2 * 2 * lattice_rows * W * 16 * 2,     # descriptor rows v-2, v+2 of every lattice row, both images, own+other
        "dense_match": 2 * 16 * N + 2 * 4 * N + 2 * 4 * N,      # both descriptor images, both owner maps, both raw maps
        "raster": 2 * 4 * N,                                     # owner maps written once
        "lr_check": 2 * 4 * N + (1 if p.postprocess_only_left else 2) * 4 * N,
        "remove_small_segments": both * 2 * 4 * N,
        "gap_interpolation": both * 2 * 4 * N,
        "adaptive_mean": both * 2 * 4 * N,
        "median": both * 2 * 4 * N,
        "reproject": 4 * N + 24 * N,
        # adaptive mean + median + final map + u8 map + point cloud in one kernel: gap-filled map in, final map / u8 / double3 out
        "post_fused": 4 * N + 4 * N + N + 24 * N,
    }


# What actually bounds each stage's kernels (ncu --set full captures under profiles/): only the stages labelled "hbm"
# are meant to be read against the HBM peak; the others carry their HBM figure for completeness.
STAGE_BOUND = {
    "descriptor": "hbm + alu (tile staging)", "support_match": "int_alu (VABSDIFF4, half-rate pipe; 16 of 19 ALU instructions per hypothesis are the SAD)", "support_filter": "latency (one CTA per frame)",
    "delaunay_device": "latency (one CTA per triangulation, sequential merges near the root)", "planes": "latency", "grid": "latency",
    "raster": "l2 atomics", "dense_match": "issue + int_alu", "lr_check": "hbm", "remove_small_segments": "issue (shared-memory union-find)",
    "gap_interpolation": "issue (validity bit words; the map is read once)", "adaptive_mean": "fp32 issue", "median": "fp32 issue", "reproject": "hbm + fp64",
    "post_fused": "fp32 / fp64 issue (about 350 instructions per pixel)", "h2d_triangles": "pcie", "d2h_support": "pcie",
}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arms -------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One host process: the reference's serial Elas::process (+ projectParallel restated in C, oracle/project_port.c) on its share of frames."""
    frames, slanted_mask, fast = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    from oracle.ref import RefElas
    import parity

    lib = synth_lib()
    ref = RefElas(fast=fast)
    p = ref.pipeline_params()
    L = np.zeros((H, W), np.uint8)
    R = np.zeros((H, W), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    t_total = 0.0
    for f in frames:
        lib.svb_synth_pair(int(f), W, H, int(f) & slanted_mask, vp(L), vp(R))  # input generation, not timed
        t0 = time.perf_counter()
        D1, _, _ = ref.process(p, L, R)
        parity.reproject_oracle(D1, Q_KITTI, np.eye(3), np.zeros(3))
        t_total += time.perf_counter() - t0
    return t_total


def cpu_frames_per_s(n_frames_per_core, cores, first_frame=0, fast=True):
    """Frame-parallel reference over `cores` host processes; returns (frames/s aggregate, wall seconds)."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    jobs = [(list(range(first_frame + c * n_frames_per_core, first_frame + (c + 1) * n_frames_per_core)), 1, fast) for c in range(cores)]
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [([], 1, fast)] * cores)  # warm: import, load the libraries
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    return cores * n_frames_per_core / wall, wall


def cpu_single_process(n_frames, omp):
    """The reference's serial (one core) or OpenMP variant in THIS process: frames/s over n_frames synthetic frames."""
    import ctypes as C

    from oracle.ref import RefElas
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity

    lib = synth_lib()
    try:
        ref = RefElas(fast=True, omp=omp)
    except (FileNotFoundError, OSError):
        return None
    p = ref.pipeline_params()
    L = np.zeros((H, W), np.uint8)
    R = np.zeros((H, W), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    t_total = 0.0
    for f in range(-1, n_frames):  # frame -1: warm-up, not timed
        lib.svb_synth_pair(max(f, 0), W, H, max(f, 0) & 1, vp(L), vp(R))
        t0 = time.perf_counter()
        D1, _, _ = ref.process(p, L, R)
        parity.reproject_oracle(D1, Q_KITTI, np.eye(3), np.zeros(3))
        if f >= 0:
            t_total += time.perf_counter() - t0
    return n_frames / t_total


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_core = max(1, args.ref_frames_per_core)
    vals = []
    for i in range(args.warmup + args.steps):
        fps, wall = cpu_frames_per_s(per_core, cores, first_frame=i * cores * per_core, fast=True)
        if i >= args.warmup:
            vals.append((fps, wall))
    fps = float(np.mean([v[0] for v in vals])) if vals else 0.0
    ms = float(np.mean([v[1] for v in vals]) * 1e3) if vals else 0.0
    sample = "%d frames per step (%d per core x %d cores), synthetic 1242x375, pipeline preset" % (cores * per_core, per_core, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 SAD + f32 filters + f64 reprojection",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": cores * per_core, "bounded_sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample + "; oracle/_ref/libelas_ref_fast.so = reference serial ELAS with the reference Makefile's "
                                            "flags (-O2 -ffast-math), one process per core, + projectParallel restated in C (oracle/project_port.c)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # this arm must not touch the product: true would mean libelas_b200.so got mapped into the process that timed the reference
        "product_library_mapped": any("libelas_b200" in ln for ln in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
    }
    print(json.dumps(line))
    return 0


# ---- GPU arm ----------------------------------------------------------------------------------------------------
def make_batch(svb, n, first, threads):
    L = np.zeros((n, H, W), np.uint8)
    R = np.zeros((n, H, W), np.uint8)
    from concurrent.futures import ThreadPoolExecutor

    def job(i):
        svb.synth_pair(first + i, W, H, (first + i) & 1, L[i], R[i])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(job, range(n)))
    return L, R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per step (BASELINE configs[1]: 1024)")
    ap.add_argument("--chunk", type=int, default=32, help="frames per kernel launch")
    ap.add_argument("--e2e-batch", type=int, default=512, help="frames per GPU per end-to-end step, the same at every N (page-locking "
                    "14 GB of result buffers per rank for 1024 frames takes a while when 8 ranks do it at once)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target wall time of the cpu_baseline leg")
    ap.add_argument("--ref-frames-per-core", type=int, default=8, help="--impl reference: frames per host core and step (about 0.35 s each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-stream", type=int, default=0)
    ap.add_argument("--delaunay-threads", type=int, default=0)
    ap.add_argument("--pin", type=int, default=0, help="N > 1: give every rank its own slice of the host cores (sched_setaffinity); "
                    "measured SLOWER on the 8-GPU box (105.8k vs 128.5k frames/s): the launcher and host-stage threads then fight the workers")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    svb = load_pkg()
    from elas_b200 import sharding

    group = sharding.Group(dist, device="cuda" if dist is not None else None)
    if svb.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; the product has no CPU path")
    cores = host_cores()
    threads_per_rank = max(2, cores // max(world, 1))
    pinned_to = None
    if world > 1 and args.pin and hasattr(os, "sched_setaffinity"):
        # one process per GPU: the Delaunay workers of a rank (created below, inheriting the mask) stay on their own cores
        allowed = sorted(os.sched_getaffinity(0))
        per = len(allowed) // world
        if per >= 1:
            pinned_to = allowed[local_rank * per:(local_rank + 1) * per]
            try:
                os.sched_setaffinity(0, pinned_to)
            except OSError:
                pinned_to = None

    p = svb.default_params(svb.PIPELINE)
    ctx = svb.Context(p, W, H, chunk=args.chunk, device=local_rank)
    ctx.set_calibration(np.array(Q_KITTI))
    ctx.set_delaunay_threads(args.delaunay_threads if args.delaunay_threads > 0 else min(threads_per_rank, 64))
    ctx.set_stage_timing(True)
    if args.single_stream:
        ctx.set_single_stream(True)

    # every rank gets its own, distinct frames (frame index = rank * batch + i): 2 * batch * N bytes resident
    t0 = time.time()
    first_frame, _ = sharding.frame_range(rank, world, args.batch)
    L, R = make_batch(svb, args.batch, first_frame, threads_per_rank)
    t_gen = time.time() - t0
    ctx.batch_upload(L, R)
    flags = svb.OUT_DISPARITY | svb.OUT_POINTS

    def barrier():
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    max_over_ranks, sum_over_ranks = group.max, group.sum

    # ---- resident-input steps ---------------------------------------------------------------------------------
    def timed_steps(n_warm, n_steps, with_clocks):
        for _ in range(n_warm):
            ctx.batch_run(args.batch, flags)
        sampler = ClockSampler(local_rank) if (rank == 0 and with_clocks) else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        barrier()
        acc = {"stage_ms": {}, "launches": 0, "delaunay_ms": 0.0, "delaunay_wall": 0.0, "gpu_ms": 0.0}
        tw0 = time.perf_counter()
        for _ in range(n_steps):
            ctx.batch_run(args.batch, flags)  # synchronises on its own end event; gpu_ms_total = CUDA-event time of the step
            st = ctx.stats()
            acc["gpu_ms"] += st["gpu_ms_total"]
            acc["launches"] += st["kernel_launches"]
            acc["delaunay_ms"] += st["delaunay_ms_total"]
            acc["delaunay_wall"] += st["delaunay_ms_wall"]
            for k, v in st["stage_ms"].items():
                acc["stage_ms"][k] = acc["stage_ms"].get(k, 0.0) + v
        barrier()
        acc["wall_ms"] = (time.perf_counter() - tw0) * 1e3
        acc["clocks"] = sampler.finish() if sampler else None
        return acc

    # (1) the headline: one stream per lane, kernels of neighbouring chunks overlap
    main_run = timed_steps(args.warmup, args.steps, True)
    clocks = main_run["clocks"]
    launches = main_run["launches"]
    delaunay_ms, delaunay_wall, wall_ms = main_run["delaunay_ms"], main_run["delaunay_wall"], main_run["wall_ms"]
    # device time of the K steps (CUDA events around each step, summed), max over ranks
    t_ms = max_over_ranks(main_run["gpu_ms"])
    frames_total = sum_over_ranks(float(args.batch * args.steps))
    value = frames_total / (t_ms * 1e-3)
    st_last = ctx.stats()
    # (2) the same K steps with every lane on ONE stream: kernels never overlap, so the CUDA events that bracket each
    #     stage measure that stage alone -- these are the durations the roofline is computed from
    if not args.single_stream:
        ctx.set_single_stream(True)
        exact_run = timed_steps(1, args.steps, False)
        ctx.set_single_stream(False)
    else:
        exact_run = main_run
    stage_ms = exact_run["stage_ms"]
    value_single_stream = sum_over_ranks(float(args.batch * args.steps)) / (max_over_ranks(exact_run["gpu_ms"]) * 1e-3)

    # ---- pixel-disparity evaluations (SURVEY.md 8d ii): one untimed step with the counting kernels -------------------
    ctx.set_eval_counting(True)
    ctx.batch_run(args.batch, flags)
    n_support_hyp, n_dense_hyp = ctx.eval_counts()
    ctx.set_eval_counting(False)
    support_hyp = sum_over_ranks(float(n_support_hyp)) / (args.batch * world)
    dense_hyp = sum_over_ranks(float(n_dense_hyp)) / (args.batch * world)

    # ---- end to end from pinned host buffers ----------------------------------------------------------------
    nb = min(args.e2e_batch if args.e2e_batch > 0 else args.batch, args.batch)
    hl = svb.PinnedArray((nb, H, W), np.uint8)
    hr = svb.PinnedArray((nb, H, W), np.uint8)
    hD = svb.PinnedArray((nb, H, W), np.float32)
    hP = svb.PinnedArray((nb, N, 3), np.float64)
    hl.array[:] = L[:nb]
    hr.array[:] = R[:nb]
    ctx.set_stage_timing(False)
    for _ in range(2):
        ctx.batch_run_host(hl.array, hr.array, flags, hD.array, hP.array)
    barrier()
    e2e_ms = 0.0
    e2e_steps = max(3, args.steps)
    for _ in range(e2e_steps):
        tt = time.perf_counter()
        ctx.batch_run_host(hl.array, hr.array, flags, hD.array, hP.array)
        _ = float(hD.array[nb - 1, H // 2, W // 2])  # the step's result is read on the host
        e2e_ms += (time.perf_counter() - tt) * 1e3
    barrier()
    e2e_t = max_over_ranks(e2e_ms)
    e2e_value = sum_over_ranks(float(nb * e2e_steps)) / (e2e_t * 1e-3)
    valid_frac = float((hD.array[0] >= 0).mean())
    for a in (hl, hr, hD, hP):
        a.free()

    # ---- the reference's own data: frames 0 and 7 of datasets/kitti_mini (committed gray fixtures), tiled to the batch ------
    kitti = None
    kpath = os.path.join(ROOT, "tests", "golden", "kitti_gray.npz")
    if os.path.exists(kpath):
        z = np.load(kpath)
        if z["L0"].shape == (H, W):
            npairs = len([k for k in z.files if k.startswith("L")])
            nk = (min(args.batch, 512) // npairs) * npairs
            Lk = np.ascontiguousarray(np.stack([z["L%d" % (i % npairs)] for i in range(nk)]))
            Rk = np.ascontiguousarray(np.stack([z["R%d" % (i % npairs)] for i in range(nk)]))
            ctx.batch_upload(Lk, Rk)
            ctx.batch_run(len(Lk), flags)
            barrier()
            k_ms, k_steps = 0.0, 2
            for _ in range(k_steps):
                ctx.batch_run(len(Lk), flags)
                k_ms += ctx.stats()["gpu_ms_total"]
            st_k = ctx.stats()
            barrier()
            k_value = sum_over_ranks(float(len(Lk) * k_steps)) / (max_over_ranks(k_ms) * 1e-3)
            kitti = {"value": k_value, "unit": UNIT, "per_gpu": k_value / world, "n_gpus": world,
                     "frames_per_gpu": len(Lk), "steps": k_steps, "support_points_per_frame": st_k["support_points"] / max(1, st_k["frames"]),
                     "host_delaunay_ms_per_frame_cpu": st_k["delaunay_ms_total"] / max(1, st_k["frames"]),
                     "delaunay_lists_device": st_k["delaunay_lists_device"], "delaunay_lists_host": st_k["delaunay_lists_host"],
                     "data": "all %d stereo pairs of datasets/kitti_mini (tests/golden/kitti_gray.npz) repeated; inputs resident, same outputs as "
                             "`value`; whole-job aggregate like `value` (weak scaling: efficiency = per_gpu at N / per_gpu at 1)" % npairs}

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dims = {"ch": ctx.ch, "cw": ctx.cw}
    sb = stage_bytes(p, dims)
    kernel_stages = {k: v for k, v in stage_ms.items() if k in sb and v > 0.2e-3 * args.batch * args.steps}
    nlaunch_per_stage = args.steps * ((args.batch + args.chunk - 1) // args.chunk)
    top = max(kernel_stages, key=kernel_stages.get) if kernel_stages else None
    roof = None
    per_stage = {}
    for k, v in kernel_stages.items():
        gbs = sb[k] * args.batch * args.steps / (v * 1e-3) / 1e9
        per_stage[k] = {"us_per_frame": 1e3 * v / (args.batch * args.steps), "GBps": gbs, "frac": gbs / peak, "bound": STAGE_BOUND.get(k)}
    for k, v in stage_ms.items():
        if k not in per_stage and v > 0:
            per_stage[k] = {"us_per_frame": 1e3 * v / (args.batch * args.steps), "bound": STAGE_BOUND.get(k)}
    # empty stages (adaptive_mean / median / reproject when the fused tail runs) only hold event overhead: leave them out
    per_stage = {k: v for k, v in per_stage.items() if v["us_per_frame"] >= 0.2 or k in ("d2h_support", "h2d_triangles")}
    if top:
        ach = per_stage[top]["GBps"]
        # dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernels per launch of `ncu_frames_per_launch` frames, from the
        # committed `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), scaled to this chunk size
        traffic, ncu_pipes = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if top in tj.get("stages", {}):
                traffic = tj["stages"][top]["dram_bytes_per_launch"] * args.chunk / tj["frames_per_launch"]
                ncu_pipes = {k: tj["stages"][top][k] for k in ("alu_pipe_pct", "issue_active_pct") if k in tj["stages"][top]}
        # the dominant kernel is a SAD kernel: its real limiter is the half-rate VABSDIFF4 pipe (16 lanes / clock / SM sub-partition), so
        # next to the HBM figure the line carries the fraction of that floor: hypotheses x SAD instructions per hypothesis / pipe rate
        sad_floor = None
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
        pipe_rate = 16.0 * 4 * 148 * sm_hz  # VABSDIFF4 lane-operations per second on the whole GPU
        if top == "support_match":
            floor_s = support_hyp * 16 / pipe_rate  # per frame: every hypothesis occupies one lane of 16 SAD instructions
            sad_floor = {"floor_us_per_frame": floor_s * 1e6, "actual_us_per_frame": per_stage[top]["us_per_frame"],
                         "frac_of_floor": floor_s * 1e6 / per_stage[top]["us_per_frame"],
                         "definition": "hypotheses per frame x 16 VABSDIFF4.U8.ACC per hypothesis / (16 lanes per clock per SM sub-partition x 592 "
                                       "sub-partitions x measured SM clock)"}
        elif top == "dense_match":
            floor_s = dense_hyp * 4 / pipe_rate
            sad_floor = {"floor_us_per_frame": floor_s * 1e6, "actual_us_per_frame": per_stage[top]["us_per_frame"],
                         "frac_of_floor": floor_s * 1e6 / per_stage[top]["us_per_frame"],
                         "definition": "hypotheses per frame x 4 VABSDIFF4.U8.ACC per hypothesis / pipe rate"}
        roof = {"kernel": top, "bound": "hbm", "bound_actual": STAGE_BOUND.get(top), "sad_floor": sad_floor,
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_launch": sb[top] * args.chunk,
                "avg_launch_ms": kernel_stages[top] / nlaunch_per_stage,
                "share_of_step": kernel_stages[top] / sum(stage_ms.values()),
                # what actually bounds this kernel (same committed ncu capture): the two SAD kernels are integer-ALU bound, DESIGN.md 5
                "ncu_pipes": ncu_pipes,
                "note": "durations from the single-stream timed steps (CUDA events on the launching stream bracket each stage; "
                        "no other kernel of ours runs concurrently), see DESIGN.md"}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ctx.close()
        per_core = max(1, int(round(args.cpu_seconds / 0.35)))  # ~0.33 s per frame per core for the reference
        per_core = min(per_core, 64)
        fps, wall = cpu_frames_per_s(per_core, cores, first_frame=0, fast=True)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "%d frames of the same synthetic workload (%d per core), reference serial ELAS (oracle/_ref, -O2 -ffast-math = "
                         "reference Makefile flags) + projectParallel restated in C, one process per core, %.1f s wall" % (per_core * cores, per_core, wall),
               # the two single-process forms the reference's own binaries take (make serial=1 / make omp=1), 8 frames each
               "serial_one_core_fps": cpu_single_process(8, omp=False),
               "openmp_one_process_fps": cpu_single_process(8, omp=True),
               "openmp_note": "src/omp_includes/elas with -fopenmp; its num_threads(2)/(3) clauses cap it at about 3 cores"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32 SAD + f32 filters + f64 reprojection", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_gpu": args.batch, "frames_per_launch": args.chunk, "lanes": int(os.environ.get("SVB_LANES", "8")), "single_stream": bool(args.single_stream),
                       "l2": "inputs larger than L2 (%.0f MB of images, %.0f MB of descriptors per step)" % (2 * N * args.batch / 1e6,
                                                                                                            32 * N * args.batch / 1e6),
                       "parallelism": "frame-batch data parallel x%d, no collective" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * N * nb, "d2h_bytes_per_step": (4 + 24) * N * nb,
                    "frames_per_step": nb, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "pixel_disparity_evals": {
                "support_hypotheses_per_frame": support_hyp, "dense_hypotheses_per_frame": dense_hyp,
                "evals_per_s": (support_hyp + dense_hyp) * value, "sad16_per_s": (4.0 * support_hyp + dense_hyp) * value,
                "note": "hypotheses the reference algorithm evaluates for exactly these frames (support: 4 x 16-byte SAD each, forward + "
                        "backward pass; dense: one 16-byte SAD each), counted on the device by an untimed step (svb_set_eval_counting)"},
            "roofline": roof,
            "kitti_mini": kitti,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "stages": per_stage,
            "stages_overlapped_us_per_frame": {k: 1e3 * v / (args.batch * args.steps) for k, v in main_run["stage_ms"].items() if v > 0},
            "value_single_stream": value_single_stream,
            "host_delaunay": {"ms_per_frame_cpu": delaunay_ms / (args.batch * args.steps), "wall_ms_per_step": delaunay_wall / args.steps,
                              "lists_device": st_last["delaunay_lists_device"], "lists_host": st_last["delaunay_lists_host"],
                              "note": "the divide-and-conquer runs on the device (k_delaunay.cu, stage delaunay_device); the host stage only "
                                      "triangulates lists the device hands back (duplicate coordinates, > 4096 points)",
                              "threads": min(threads_per_rank, 64) if args.delaunay_threads <= 0 else args.delaunay_threads},
            "wall_ms_per_step": wall_ms / args.steps,
            "rank0_pinned_to_cores": pinned_to,
            "support_points_per_frame": st_last["support_points"] / max(1, st_last["frames"]),
            "valid_fraction_frame0": valid_frac,
            "input_generation_s": t_gen,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
