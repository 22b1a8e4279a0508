#!/usr/bin/env python
"""bench.py -- views/s and Mpix/s of pix2face + aggregate_projected_images on synthetic surveys.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload = BASELINE.json config c2: 2 000 000-face canopy mesh, 500 Metashape-calibrated cameras 5472 x 3648,
10-class float32 softmax scores.  One "step" = `--views-per-step` (500 = the whole survey) views through the hot path:
camera projection + tiled z-buffer rasterization (pix2face) + per-face aggregation, issued as batches of `--batch`
views.  The mesh is replicated on every GPU and cameras are sharded by contiguous blocks (weak scaling: every rank
times K steps of its own shard); the per-face float64 sums and int32 counts of all ranks are combined with ONE NCCL
all-reduce (counts packed behind the sums) at the end of the timed region, followed by the mean + argmax epilogue.
Rank 0 prints ONE JSON line:

`value`       device-resident: score images already in HBM (a ring of distinct 0.8 GB buffers, >> L2), reference-parity
              (last-pixel) aggregation.
`roofline`    for the dominant kernel (k_raster_tiles), timed with CUDA events on its own stream inside the library.
              The parity mode moves ~0.1 GB of HBM traffic per 20-Mpx view and is bound by instruction issue, so its
              roofline is an ISSUE roofline (warp instructions per second against SMs x 4 schedulers x clock, the
              instruction count per launch coming from the committed ncu capture, profiles/inst_counts.json); the
              HBM fraction of the same kernel is kept under `roofline.hbm`.  The dense `pixel_sum` mode, which
              streams every score, carries an HBM roofline.
`e2e`         the same metric through the reference-facing Python API (TexturedPhotogrammetryMesh.
              aggregate_projected_images on a SegmentorPhotogrammetryCameraSet) with the score images in HOST memory
              (geograypher_b200.host_array); the rows the aggregation needs cross PCIe inside the timed region and the
              per-face results are copied back.  `e2e_pinned` (page-locked arrays), `e2e_index_u8` (class-index images,
              the LookUpSegmentor contract) and `e2e_pageable` (ordinary NumPy arrays) are the same call on other
              input kinds (N = 1 only).
`pixel_sum`, `c3_strong`, `c4`, `c5`   sub-records: dense mode; the 500-view survey split over the N ranks (strong
              scaling, all-reduce + epilogue + device-to-host copy inside the timing, parity-checked against a
              single-GPU run); render_flat to uint8 label rasters; the 20M-face one-hot-vote survey.
`cpu_baseline` / `--impl reference`   the CPU restatement of the reference's path (oracle/: OpenMP C rasterizer with a
              per-view culling pass on all host cores + the reference's literal single-threaded NumPy aggregation) on
              a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "views/s, pix2face+aggregate (Mpix/s in extras)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "tiny"])
    ap.add_argument("--mode", default="last_pixel", choices=["last_pixel", "pixel_sum"])
    ap.add_argument("--views-per-step", type=int, default=500)
    ap.add_argument("--batch", type=int, default=10, help="views per kernel launch")
    ap.add_argument("--e2e-views", type=int, default=500, help="views per rank in the end-to-end (host buffer) runs")
    ap.add_argument("--cpu-views", type=int, default=12, help="views in the CPU-baseline sample")
    ap.add_argument("--skip", default="", help="comma-separated legs to skip: pixel_sum,c3,c4,c5,e2e,e2e_extra,cpu")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.15)  # let the first sample land before the timed region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs that are local to GPU `index` (pinned host buffers are then allocated on the
    NUMA node behind the GPU's PCIe root, which matters for the zero-copy reads of the end-to-end leg)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def committed_number(filename, key):
    """A per-launch figure of the dominant kernel taken from a committed ncu capture of this command
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum; profiles/inst_counts.json:
    smsp__inst_executed.sum); None when no capture matches."""
    try:
        return json.loads((ROOT / "profiles" / filename).read_text()).get(key)
    except Exception:
        return None


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_survey(name, max_cameras=None):
    from geograypher_b200 import synthetic as syn

    verts, faces, c2ws, cfg = syn.make_survey(name, max_cameras)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    return verts, faces, c2ws, cfg, origin


def shard(n_items, rank, world):
    per = -(-n_items // world)
    return list(range(min(n_items, rank * per), min(n_items, (rank + 1) * per)))


# ----------------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference): used by cpu_baseline and by --impl reference
# ----------------------------------------------------------------------------------------------------------
def cpu_predictions(cfg, n_distinct=2):
    import torch

    W, H = cfg.image_size
    out = []
    for i in range(n_distinct):
        gen = torch.Generator().manual_seed(1000 + i)
        g = torch.randn((1, cfg.n_classes, 43, 64), generator=gen) * 1.5
        logits = torch.nn.functional.interpolate(g, size=(H, W), mode="bilinear", align_corners=True)[0]
        out.append(torch.softmax(logits, dim=0).permute(1, 2, 0).contiguous().numpy())
    return out


def host_threads():
    """All host cores, regardless of OMP_NUM_THREADS (torchrun sets it to 1)."""
    return max(1, os.cpu_count() or 1)


def cpu_views_per_second(verts, faces, c2ws, cfg, origin, view_ids, preds):
    """Oracle port of pix2face + aggregate_projected_images over `view_ids`; returns (seconds per view list,
    raster seconds, aggregate seconds)."""
    from oracle import oracle as ora

    W, H = cfg.image_size
    v32 = (verts - origin).astype(np.float32)
    F = len(faces)
    counts = np.zeros(F)
    summed = None
    t_r = t_a = 0.0
    per_view = []
    for j, k in enumerate(view_ids):
        t0 = time.perf_counter()
        cam = ora.make_camera(c2ws[k], cfg.f, cfg.cx, cfg.cy, W, H, origin=origin)
        p2f = ora.rasterize(v32, faces, cam, nthreads=host_threads())
        t1 = time.perf_counter()
        proj = ora.project_image(p2f, preds[j % len(preds)], F)  # meshes.py:1988-2001
        summed = proj.astype(float) if summed is None else np.nansum([summed, proj], axis=0)  # :2056-2062
        counts += np.any(np.isfinite(proj), axis=1).astype(int)  # :2064-2067
        t2 = time.perf_counter()
        t_r += t1 - t0
        t_a += t2 - t1
        per_view.append(t2 - t0)
    return per_view, t_r, t_a


CPU_NOTE = ("OpenMP C rasterizer (-O3 -march=native, per-view culling pass) on {cores} threads + single-threaded NumPy "
            "aggregation (the reference's aggregation is single-threaded NumPy)")


def run_reference(args):
    """--impl reference: the CPU restatement, one view per step, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    verts, faces, c2ws, cfg, origin = build_survey(args.config)
    preds = cpu_predictions(cfg)
    W, H = cfg.image_size
    n = args.warmup + args.steps
    ids = [(7 + 3 * i) % len(c2ws) for i in range(n)]
    per_view, t_r, t_a = cpu_views_per_second(verts, faces, c2ws, cfg, origin, ids, preds)
    timed = per_view[args.warmup:]
    total = float(sum(timed))
    value = len(timed) / total
    cores = host_threads()
    sample = (f"{len(timed)} views of {args.config}: each CPU step is ONE view of the {args.views_per_step}-view step the "
              f"GPU arm times; ") + CPU_NOTE.format(cores=cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the SAME config record as the GPU arm (the workload is what is compared); how much of it one CPU step covers
        # is a property of this arm and lives in cpu_baseline.sample
        "config": workload_config(args, cfg, args.views_per_step, args.batch),
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample,
                         "raster_s_per_view": t_r / n, "aggregate_s_per_view": t_a / n},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mpix_per_s": value * W * H / 1e6, "gpu_launches": 0,
        "note": "reference = CPU restatement of geograypher's path (oracle/): the reference's own rasterizer is "
                "un-vendored VTK/PyTorch3D and cannot be installed offline",
    }
    print(json.dumps(line))


def workload_config(args, cfg, views_per_step, batch):
    W, H = cfg.image_size
    return {
        "workload": f"{cfg.name}: {cfg.n_faces} faces, {cfg.n_cameras} cameras {W}x{H}, {cfg.n_classes}-class "
                    f"float32 softmax scores",
        "mode": args.mode, "views_per_step": views_per_step, "views_per_launch": batch, "pixels_per_view": W * H,
        "cache": "inputs larger than L2 (ring of distinct prediction buffers, >= 0.8 GB each)",
    }


# ----------------------------------------------------------------------------------------------------------
# GPU path
# ----------------------------------------------------------------------------------------------------------
class Timer:
    """barrier + synchronize, CUDA events on the current stream, max over ranks."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def __enter__(self):
        t = self.torch
        if self.world > 1:
            self.dist.barrier()
        t.cuda.synchronize()
        self.ev = [t.cuda.Event(enable_timing=True)]
        self.ev[0].record()
        return self

    def mark(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append(e)

    def __exit__(self, *exc):
        t = self.torch
        self.mark()
        t.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        ms = [self.ev[i].elapsed_time(self.ev[i + 1]) for i in range(len(self.ev) - 1)]
        ms.append(self.ev[0].elapsed_time(self.ev[-1]))
        v = t.tensor(ms, dtype=t.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(v, op=self.dist.ReduceOp.MAX)
        v = v.tolist()
        self.phases_ms, self.total_ms = v[:-1], v[-1]
        return False


def packed_allreduce(torch, dist, d_sum, d_count, pack):
    """ONE collective for both accumulators: the int32 counts ride behind the float64 sums as float64 (exact: counts
    are far below 2^53).  `pack` is a preallocated (F*(C+1),) float64 buffer."""
    F, C = d_sum.shape
    pack[: F * C].copy_(d_sum.reshape(-1))
    pack[F * C:].copy_(d_count)
    dist.all_reduce(pack)
    d_sum.copy_(pack[: F * C].view(F, C))
    d_count.copy_(pack[F * C:])


def run_ours(args):
    import torch
    import torch.distributed as dist

    import geograypher_b200 as gg
    from geograypher_b200 import _lib
    from geograypher_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    skip = set(filter(None, args.skip.split(",")))

    verts, faces, c2ws, cfg, origin = build_survey(args.config)
    W, H = cfg.image_size
    C, F, P = cfg.n_classes, len(faces), W * H
    B = args.batch
    VPS = args.views_per_step
    n_batches = -(-VPS // B)
    my_cams = shard(len(c2ws), rank, world)
    peak, peak_src = measured_peak_gbs()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    # The host images of the end-to-end legs are allocated first, while the host's memory is still unfragmented: the
    # GPU reads scattered rows out of them over PCIe.
    e2e_host = None
    if "e2e" not in skip:
        e2e_host = []
        for i in range(min(8, args.e2e_views)):  # gg.host_array: host memory that the GPU reads in place over PCIe
            a = gg.host_array((H, W, C), np.float32, device=local_rank)
            torch.from_numpy(a).copy_(syn.softmax_predictions_device(my_cams[i % len(my_cams)], H, W, C, dev))
            e2e_host.append(a)
        torch.cuda.synchronize()

    ctx = _lib.Context(local_rank)
    ctx.set_mesh(torch.from_numpy((verts - origin).astype(np.float32)).to(dev), torch.from_numpy(faces).to(dev))
    w2c = [np.linalg.inv(T) for T in c2ws]
    all_cams = [_lib.make_camera(w2c[k], cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for k in range(len(c2ws))]

    # prediction ring, resident in HBM, generated outside the timed region
    # (view k always reads ring[k % B], on every rank: the multi-rank result can be checked against a single-GPU run)
    ring = [syn.softmax_predictions_device(i, H, W, C, dev) for i in range(B)]
    d_sum = torch.zeros((F, C), dtype=torch.float64, device=dev)
    d_count = torch.zeros((F,), dtype=torch.int32, device=dev)
    pack = torch.empty((F * (C + 1),), dtype=torch.float64, device=dev) if world > 1 else None

    def run_views(ids, mode, check=False):
        """aggregate the views `ids` (indices into the survey's cameras) in batches of B"""
        for s in range(0, len(ids), B):
            part = ids[s:s + B]
            ctx.project_aggregate([all_cams[k] for k in part], [ring[k % B] for k in part], _lib.PRED_F32, C, mode, 0,
                                  d_sum, d_count, check=check)

    def epilogue(want_avg=True):
        ctx.drain()  # the accumulators are written on the library's internal streams
        if world > 1:
            packed_allreduce(torch, dist, d_sum, d_count, pack)
        return ctx.finalize(d_sum, d_count, want_avg=want_avg)

    def step_ids(i):
        return [my_cams[(i * VPS + j) % len(my_cams)] for j in range(VPS)]

    def timed_run(run_mode):
        """W warm-up steps, then K timed steps + all-reduce + finalize, bracketed by barrier + synchronize; the
        time is the max over ranks."""
        d_sum.zero_()
        d_count.zero_()
        run_views(step_ids(0)[:2 * B], run_mode, check=True)  # a scratch overflow grows the scratch and replays here
        for i in range(args.warmup):
            run_views(step_ids(i), run_mode)
        ctx.sync()
        stats = ctx.last_batch_stats(B)
        d_sum.zero_()
        d_count.zero_()
        ctx.profile(True)
        ctx.profile_read(reset=True)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        with Timer(torch, dist, world, dev) as tm:
            for i in range(args.steps):
                run_views(step_ids(args.warmup + i), run_mode)
            epilogue()
        elapsed_s = tm.total_ms / 1e3
        clocks = sampler.stop() if rank == 0 else None
        ctx.sync()  # surfaces a scratch overflow of the timed batches
        prof = ctx.profile_read(reset=True)
        ctx.profile(False)
        views = args.steps * VPS * world
        raster_ms, raster_launches = prof["raster_tiles"]
        f_v = float(stats[:, 1].mean())
        avg_ms = raster_ms / max(raster_launches, 1)
        mode_name = "pixel_sum" if run_mode == _lib.MODE_PIXEL_SUM else "last_pixel"
        key = f"{args.config}:{mode_name}:{B}"
        # algorithmic bytes per view.  Stage 1+2: 12 V_v + 12 F_v with V_v ~ F_v / 2; the face-ID raster
        # (4 P) is NOT written by the fused paths and is not credited.  Dense stage 3: s*C*P_hit scores +
        # read-modify-write of the float64 sums and int32 counts of the touched faces.
        b12 = 12.0 * (f_v / 2.0) + 12.0 * f_v
        if run_mode == _lib.MODE_PIXEL_SUM:
            px_added = float(d_count.sum().item()) / views  # only pixels that hit the mesh contribute scores
            bytes_per_view = b12 + 4.0 * C * px_added + (16.0 * C + 8.0) * f_v
        else:
            bytes_per_view = b12
        hbm_achieved = bytes_per_view * B / (avg_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
        hbm = {"bound": "hbm", "kernel": "k_raster_tiles", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s",
               "frac": hbm_achieved / peak, "traffic": committed_number("traffic.json", key), "peak_source": peak_src,
               "algorithmic_bytes_per_launch": bytes_per_view * B, "avg_launch_ms": avg_ms}
        if run_mode == _lib.MODE_PIXEL_SUM:
            roofline = hbm
        else:
            inst = committed_number("inst_counts.json", key)
            clock_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
            issue_peak = sm_count * 4 * clock_mhz * 1e6 / 1e9  # G warp-instructions / s
            achieved = inst / (avg_ms * 1e-3) / 1e9 if (inst and raster_ms > 0) else None
            roofline = {"bound": "issue", "kernel": "k_raster_tiles", "achieved": achieved, "peak": issue_peak,
                        "unit": "Gwarp-inst/s", "frac": (achieved / issue_peak) if achieved else None,
                        "warp_instructions_per_launch": inst, "avg_launch_ms": avg_ms,
                        "peak_source": f"{sm_count} SMs x 4 schedulers x {clock_mhz:.0f} MHz (SM clock sampled under load)",
                        "traffic": hbm["traffic"], "hbm": hbm,
                        "note": "reference-parity (last_pixel) aggregation needs ~0.02 GB of HBM traffic per 20-Mpx view: "
                                "the rasterizer is bound by instruction issue, not by HBM (DESIGN.md section 5); "
                                "instruction count per launch from the committed ncu capture of THIS kernel "
                                "(profiles/inst_counts.json), time measured live with the binning and resolve kernels "
                                "of the neighbouring batches sharing the SMs (0.70 under ncu, kernel alone); the "
                                "numerator is the kernel's own instruction count, so a leaner kernel lowers this "
                                "fraction while raising views/s"}
            # all kernels of the step together: what the software pipeline as a whole makes of the issue slots
            step_inst = committed_number("inst_counts.json", key + ":step")
            if step_inst and elapsed_s > 0:
                step_rate = step_inst * args.steps * n_batches / elapsed_s / 1e9
                roofline["step_issue"] = {
                    "bound": "issue", "kernels": "reset, cull, setup, reserve, fill, k_raster_tiles, winner reset, resolve",
                    "achieved": step_rate, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": step_rate / issue_peak,
                    "warp_instructions_per_batch": step_inst,
                    "note": "committed ncu instruction counts of every kernel of a batch (profiles/inst_counts.json, "
                            "key ...:step) x batches per rank / the timed region (all-reduce and epilogue included)"}
        return {"value": views / elapsed_s, "elapsed_s": elapsed_s, "roofline": roofline, "clocks": clocks,
                "stage_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0},
                "launches": int(sum(v[1] for v in prof.values())), "faces_per_view": f_v,
                "faces_observed": int((d_count > 0).sum().item())}

    headline_mode = {"last_pixel": _lib.MODE_LAST_PIXEL, "pixel_sum": _lib.MODE_PIXEL_SUM}[args.mode]
    other_mode = _lib.MODE_PIXEL_SUM if headline_mode == _lib.MODE_LAST_PIXEL else _lib.MODE_LAST_PIXEL
    other_name = "pixel_sum" if headline_mode == _lib.MODE_LAST_PIXEL else "last_pixel"
    other = timed_run(other_mode) if other_name not in skip else None
    main = timed_run(headline_mode)

    # ---- strong scaling: exactly the 500-view survey over the N ranks, parity-checked ------------------------------
    c3 = None
    if "c3" not in skip:
        c3 = run_c3_strong(args, torch, dist, _lib, ctx, world, rank, dev, len(c2ws), run_views, d_sum, d_count, pack, F, C)

    del ring  # 8 GB
    torch.cuda.empty_cache()
    c4 = run_c4(args, torch, dist, _lib, syn, ctx, world, rank, dev, verts, faces, all_cams, cfg, peak, peak_src) \
        if "c4" not in skip else None
    if c4 is not None and world == 1 and "e2e" not in skip and "e2e_extra" not in skip:
        try:  # a secondary leg: a failure is recorded in the line instead of costing it
            c4["e2e"] = run_c4_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, rank)
        except Exception as e:  # noqa: BLE001
            c4["e2e"] = {"error": f"{type(e).__name__}: {e}"}
    e2e = e2e_idx = e2e_page = e2e_pin = None
    if "e2e" not in skip:
        e2e = run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, e2e_host, "host_array_f32")
        if world == 1 and "e2e_extra" not in skip:
            e2e_idx = run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, None, "pinned_index_u8")
            e2e_pin = run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, e2e_host, "pinned_f32")
            e2e_page = run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, e2e_host, "pageable_f32")
    e2e_host = None

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and "cpu" not in skip:
        ids = [(7 + 3 * i) % len(c2ws) for i in range(args.cpu_views)]
        per_view, t_r, t_a = cpu_views_per_second(verts, faces, c2ws, cfg, origin, ids, cpu_predictions(cfg))
        cpu = {"value": len(per_view) / float(sum(per_view)), "unit": "views/s", "cores": host_threads(), "kind": "port",
               "sample": f"{len(per_view)} views of {args.config}; " + CPU_NOTE.format(cores=host_threads()),
               "raster_s_per_view": t_r / len(per_view), "aggregate_s_per_view": t_a / len(per_view)}

    # ---- the large survey (its 20M-face mesh replaces the c2 mesh on the device) ---------------------------------
    ctx.close()
    del ctx, d_sum, d_count, pack
    torch.cuda.empty_cache()
    c5 = run_c5(args, torch, dist, _lib, syn, world, rank, local_rank, dev) if "c5" not in skip else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * main["elapsed_s"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, VPS, B),
            "mpix_per_s": main["value"] * P / 1e6, "roofline": main["roofline"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": main["launches"], "clocks": main["clocks"], "stage_ms": main["stage_ms"],
            "faces_per_view": main["faces_per_view"], "faces_observed": main["faces_observed"],
            other_name: None if other is None else {
                "value": other["value"], "unit": "views/s", "mpix_per_s": other["value"] * P / 1e6,
                "ms_per_step": 1e3 * other["elapsed_s"] / args.steps, "roofline": other["roofline"],
                "stage_ms": other["stage_ms"],
                "note": "same workload with every pixel adding its scores (GG_MODE_PIXEL_SUM, not the reference's "
                        "semantics): the fused rasterizer epilogue streams the (H,W,C) float32 scores from HBM"
                        if other_name == "pixel_sum" else "reference-parity mode"},
            "c3_strong": c3, "c4": c4, "c5": c5, "e2e_index_u8": e2e_idx, "e2e_pinned": e2e_pin, "e2e_pageable": e2e_page,
            "parity_check": (c3 or {}).get("parity_check"),
            "accumulators": "float64 sums + int32 counts" + ("; one NCCL all-reduce (counts packed behind the sums) at "
                                                              "the end" if world > 1 else ""),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c3_strong(args, torch, dist, _lib, ctx, world, rank, dev, n_cams, run_views, d_sum, d_count, pack, F, C):
    """BASELINE config 3: the SAME 500-view survey split over the N ranks.  Timed: every rank's share of the views,
    the collective, the mean + argmax epilogue and the device-to-host copy of averages, sums and counts.  One GPU: the
    copy goes into page-locked buffers.  N > 1 (geograypher_b200.distributed.finalize_sharded): ONE reduce-scatter of
    the packed accumulators, every rank finishes its slice of the faces and copies it over its own PCIe link into a
    host block shared by the node's ranks; rank 0 holds the complete result when the timed region ends."""
    from geograypher_b200 import distributed as ggd

    mine = shard(n_cams, rank, world)
    mode = _lib.MODE_LAST_PIXEL
    sharded = world > 1
    if sharded:
        try:
            res = ggd.SharedHostResult.get(F, C, None, 0)  # created (and page-locked) once, outside the timed region
            lo, hi = ggd.face_slice(F, rank, world)
            h_avg, h_sum, h_cnt = (torch.from_numpy(a[lo:hi]) for a in (res.avg, res.sums, res.counts))
        except OSError:  # no room in /dev/shm, or it cannot be page-locked here (every rank raises): all-reduce instead
            sharded = False
    if not sharded:
        host = [torch.empty((F, C), dtype=torch.float64, pin_memory=True) for _ in range(2)] if rank == 0 else None
        host_cnt = torch.empty((F,), dtype=torch.int32, pin_memory=True) if rank == 0 else None

    def once():
        d_sum.zero_()
        d_count.zero_()
        with Timer(torch, dist, world, dev) as tm:
            run_views(mine, mode)
            ctx.drain()
            tm.mark()
            if sharded:
                s_sum, s_cnt = ggd.reduce_scatter_accumulators(d_sum, d_count)
                tm.mark()
                avg, argmax = ctx.finalize(s_sum, s_cnt)
                tm.mark()
                k = hi - lo
                h_avg.copy_(avg[:k], non_blocking=True)
                h_sum.copy_(s_sum[:k], non_blocking=True)
                h_cnt.copy_(s_cnt[:k].double(), non_blocking=True)
            else:
                if world > 1:
                    packed_allreduce(torch, dist, d_sum, d_count, pack)
                tm.mark()
                avg, argmax = ctx.finalize(d_sum, d_count)
                tm.mark()
                if rank == 0:
                    host[0].copy_(avg, non_blocking=True)
                    host[1].copy_(d_sum, non_blocking=True)
                    host_cnt.copy_(d_count, non_blocking=True)
        ctx.sync()
        return tm

    once()  # warm-up (NCCL channels, pinned staging)
    tm = min((once() for _ in range(3)), key=lambda t: t.total_ms)
    names = ["views", "reduce_scatter" if sharded else "allreduce", "finalize", "d2h"]  # (chosen before the timed runs)
    out = {"value": n_cams / (tm.total_ms / 1e3), "unit": "views/s", "scaling": "strong", "views": n_cams,
           "views_per_rank": len(mine), "total_ms": tm.total_ms,
           "phase_ms": {k: round(v, 3) for k, v in zip(names, tm.phases_ms)},
           "value_without_d2h": n_cams / (sum(tm.phases_ms[:3]) / 1e3),
           "collective": ("one reduce_scatter of F*(C+1) float64 (counts packed behind the sums of the same face slice); "
                          "every rank copies its slice of the result into a host block shared by the node's ranks"
                          if sharded else ("none (1 GPU)" if world == 1 else
                                           "one all_reduce of F*(C+1) float64 (counts packed behind the sums); rank 0 "
                                           "copies the result to the host (the shared host block was unavailable)")),
           "note": "best of 3; max over ranks; finalize's NaN marking of unseen faces is part of the D2H'd sums"}
    # parity: a single GPU aggregates the whole survey alone; the result the ranks assembled on the host must agree
    if world > 1:
        ok = torch.ones(1, device=dev)
        if rank == 0:
            got = (res.avg, res.sums, res.counts) if sharded else (host[0].numpy(), host[1].numpy(),
                                                                   host_cnt.numpy().astype(np.float64))
            got = [np.array(a, copy=True) for a in got]
            d_sum.zero_()
            d_count.zero_()
            run_views(list(range(n_cams)), mode)
            one_avg, _ = ctx.finalize(d_sum, d_count)
            ctx.sync()
            same_cnt = np.array_equal(got[2], d_count.double().cpu().numpy())
            close = (np.allclose(got[1], d_sum.cpu().numpy(), rtol=1e-12, atol=0.0, equal_nan=True)
                     and np.allclose(got[0], one_avg.cpu().numpy(), rtol=1e-12, atol=0.0, equal_nan=True))
            ok[0] = 1.0 if (same_cnt and close) else 0.0
        dist.broadcast(ok, src=0)
        if ok.item() != 1.0:
            raise RuntimeError("parity check failed: the result assembled by the ranks differs from the single-GPU run")
        out["parity_check"] = "ok"
        out["parity_note"] = (f"rank 0 re-ran all {n_cams} views alone: counts identical, float64 sums and averages within "
                              f"1e-12 relative of the result the {world} ranks "
                              + ("assembled in the shared host block" if sharded else "all-reduced"))
    return out


def run_c4(args, torch, dist, _lib, syn, ctx, world, rank, dev, verts, faces, all_cams, cfg, peak, peak_src):
    """BASELINE config 4: render_flat of per-face polygon labels to uint8 label rasters (no coupling between views:
    replicas only; every rank renders its shard of the cameras)."""
    W, H = cfg.image_size
    P, B = W * H, args.batch
    mine = shard(len(all_cams), rank, world)
    tex = torch.from_numpy(syn.voronoi_face_labels(verts, faces)).to(dev)
    out = torch.empty((B, H, W, 1), dtype=torch.uint8, device=dev)
    n_views = min(len(mine), max(B, args.views_per_step) // B * B)
    ids = mine[:n_views]

    def run():
        for s in range(0, len(ids), B):
            part = ids[s:s + B]
            ctx.rasterize_render_flat([all_cams[k] for k in part], tex, out_dtype=_lib.OUT_U8, out=out[:len(part)],
                                      check=False)

    ctx.rasterize_render_flat([all_cams[k] for k in ids[:B]], tex, out_dtype=_lib.OUT_U8, out=out)  # checked warm-up
    run()
    ctx.sync()
    stats = ctx.last_batch_stats(B)
    ctx.profile(True)
    ctx.profile_read(reset=True)
    with Timer(torch, dist, world, dev) as tm:
        run()
    ctx.sync()
    prof = ctx.profile_read(reset=True)
    ctx.profile(False)
    views = n_views * world
    raster_ms, launches = prof["raster_tiles"]
    avg_ms = raster_ms / max(launches, 1)
    f_v = float(stats[:, 1].mean())
    bytes_per_view = 1.0 * P + 12.0 * (f_v / 2.0) + 12.0 * f_v  # B4 fused: uint8 out, no ID raster
    achieved = bytes_per_view * B / (avg_ms * 1e-3) / 1e9
    labelled = float((out > 0).float().mean().item())
    inst = committed_number("inst_counts.json", f"c4:render_flat:{B}")
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    issue_peak = sm_count * 4 * 1965.0 * 1e6 / 1e9  # G warp-instructions / s at the B200's 1965 MHz
    issue = ({"bound": "issue", "achieved": inst / (avg_ms * 1e-3) / 1e9, "peak": issue_peak, "unit": "Gwarp-inst/s",
              "frac": inst / (avg_ms * 1e-3) / 1e9 / issue_peak, "warp_instructions_per_launch": inst,
              "peak_source": f"{sm_count} SMs x 4 schedulers x 1965 MHz"} if inst and avg_ms > 0 else None)
    return {"value": views / (tm.total_ms / 1e3), "unit": "views/s", "mpix_per_s": views / (tm.total_ms / 1e3) * P / 1e6,
            "views": views, "total_ms": tm.total_ms, "scaling": "weak (replicas only)",
            "workload": f"c4: render_flat of per-face labels (200 Voronoi polygons, 10 classes, 20% unlabelled) to "
                        f"{W}x{H} uint8 rasters, fused gg_rasterize_render_flat, {B} views per launch, rasters stay in HBM",
            "roofline": {"bound": "hbm", "kernel": "k_raster_tiles<GATHER>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": committed_number("traffic.json", f"c4:render_flat:{B}"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_view * B, "avg_launch_ms": avg_ms,
                         "issue": issue,
                         "note": "the kernel is issue bound like the aggregation's rasterizer (see `issue`: instruction "
                                 "count from the committed ncu capture, time measured live); the uint8 raster is its "
                                 "only sizeable HBM traffic"},
            "stage_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0}, "labelled_pixel_fraction": labelled}


def run_c4_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, rank):
    """BASELINE config 4 end to end: TexturedPhotogrammetryMesh.render_flat through the reference-facing API, every
    render delivered to HOST memory inside the timed region (page-locked blocks filled on a copy stream while the next
    batch is rendered; the consumer looks at each image and drops it).  Two output types: uint8 label rasters (the
    save_renders cast rule applied on the GPU: one byte per pixel crosses PCIe) and the reference's own contract,
    float64 (h, w, d) arrays (8 bytes per pixel and channel).  Replicas only: every rank renders its own cameras."""
    W, H = cfg.image_size
    P, B = W * H, args.batch
    mine = shard(len(c2ws), rank, world)
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}

    def cams_of(n):
        return gg.PhotogrammetryCameraSet(cam_to_world_transforms=[c2ws[mine[i % len(mine)]] for i in range(n)],
                                          intrinsic_params_per_sensor_type=intr)

    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), device=dev.index, views_per_batch=B, log_level="WARNING")
    mesh.set_texture(syn.voronoi_face_labels(verts, faces), is_vertex_texture=False)

    def timed(n, batch_size, **kw):
        for img in mesh.render_flat(cams_of(3 * batch_size), batch_size=batch_size, **kw):  # warm-up: page-locks the blocks
            pass
        cams = cams_of(n)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got, probe = 0, 0.0
        for img in mesh.render_flat(cams, batch_size=batch_size, **kw):
            got += 1
            probe += float(np.nan_to_num(img[H // 2, W // 2, 0]))  # the consumer reads the host copy
            del img
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert got == n
        return float(dt.item()), probe

    n_u8 = min(args.e2e_views, 500)
    dt_u8, _ = timed(n_u8, B, out_dtype="uint8")
    n_f64 = min(args.e2e_views, 40)
    dt_f64, _ = timed(n_f64, 2)
    del mesh
    torch.cuda.empty_cache()
    return {"value": n_u8 * world / dt_u8, "unit": "views/s", "views": n_u8 * world, "seconds": dt_u8,
            "h2d_bytes_per_view": 0, "d2h_bytes_per_view": P, "d2h_gbs": n_u8 * world * P / dt_u8 / 1e9,
            "api": f"TexturedPhotogrammetryMesh.render_flat(cameras, batch_size={B}, out_dtype='uint8')",
            "float64": {"value": n_f64 * world / dt_f64, "unit": "views/s", "views": n_f64 * world, "seconds": dt_f64,
                        "d2h_bytes_per_view": 8 * P, "d2h_gbs": n_f64 * world * 8 * P / dt_f64 / 1e9,
                        "api": "TexturedPhotogrammetryMesh.render_flat(cameras, batch_size=2): the reference's "
                               "contract, float64 (h, w, d) arrays"},
            "note": "every render is copied to host memory inside the timed region (the device-to-host copy of a batch "
                    "overlaps the rendering of the next); both legs are bound by the PCIe link's device-to-host rate "
                    "(d2h_gbs), not by the rasterizer"}


def run_c5(args, torch, dist, _lib, syn, world, rank, local_rank, dev):
    """BASELINE config 5: 20M-face mesh, 2000 rig cameras 8192 x 5460, one-hot voting from class-index images
    (TexturedPhotogrammetryMeshIndexPredictions semantics, GG_MODE_VOTE), cameras sharded over the ranks, one
    all-reduce of the vote accumulators."""
    verts, faces, c2ws, cfg, origin = build_survey("c5")
    W, H = cfg.image_size
    C, F, P, B = cfg.n_classes, len(faces), W * H, 8
    ctx = _lib.Context(local_rank)
    ctx.set_mesh(torch.from_numpy((verts - origin).astype(np.float32)).to(dev), torch.from_numpy(faces).to(dev))
    mine = shard(len(c2ws), rank, world)
    n_views = min(len(mine), 240) // B * B
    ids = mine[:n_views]
    cams = {k: _lib.make_camera(np.linalg.inv(c2ws[k]), cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for k in ids}
    host_preds = [syn.class_index_image(ids[i], H, W, C) for i in range(B)]
    preds = [torch.from_numpy(h).to(dev) for h in host_preds]
    d_sum = torch.zeros((F, C), dtype=torch.float64, device=dev)
    d_count = torch.zeros((F,), dtype=torch.int32, device=dev)
    pack = torch.empty((F * (C + 1),), dtype=torch.float64, device=dev) if world > 1 else None

    def run(sub, check=False):
        for s in range(0, len(sub), B):
            part = sub[s:s + B]
            ctx.project_aggregate([cams[k] for k in part], preds[:len(part)], _lib.PRED_U8, C, _lib.MODE_VOTE, 0, d_sum,
                                  d_count, check=check)

    run(ids[:2 * B], check=True)
    run(ids)
    ctx.sync()
    stats = ctx.last_batch_stats(B)
    if world > 1:  # first use of a 1.8 GB message: let NCCL set its channels / registrations up outside the timing
        packed_allreduce(torch, dist, d_sum, d_count, pack)
    d_sum.zero_()
    d_count.zero_()
    ctx.profile(True)
    ctx.profile_read(reset=True)
    with Timer(torch, dist, world, dev) as tm:
        run(ids)
        ctx.drain()
        tm.mark()
        if world > 1:
            packed_allreduce(torch, dist, d_sum, d_count, pack)
        tm.mark()
        ctx.finalize(d_sum, d_count, want_avg=False)
    ctx.sync()
    prof = ctx.profile_read(reset=True)
    ctx.profile(False)
    views = n_views * world
    value = views / (tm.total_ms / 1e3)
    out = {"value": value, "unit": "views/s", "mpix_per_s": value * P / 1e6, "views": views, "views_per_rank": n_views,
           "total_ms": tm.total_ms, "scaling": "weak",
           "phase_ms": {k: round(v, 3) for k, v in zip(["views", "allreduce", "finalize"], tm.phases_ms)},
           "workload": f"c5: {F} faces, {cfg.n_cameras} rig cameras {W}x{H} (nadir + 4 obliques per station), uint8 "
                       f"class-index images with 2% ignored pixels, one-hot votes (GG_MODE_VOTE), {B} views per launch",
           "faces_per_view": float(stats[:, 1].mean()), "tile_entries_per_view": float(stats[:, 2].mean()),
           "faces_observed": int((d_count > 0).sum().item()), "votes": int(torch.nansum(d_sum).item()),
           "accumulator_bytes": F * C * 8 + F * 4,
           "stage_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0}}
    ctx.close()
    del ctx, d_sum, d_count, pack, preds
    torch.cuda.empty_cache()
    if world == 1 and "e2e" not in args.skip.split(",") and "e2e_extra" not in args.skip.split(","):
        try:  # a secondary leg: a failure is recorded in the line instead of costing it
            out["e2e"] = run_c5_e2e(args, torch, cfg, verts, faces, c2ws, ids, dev, host_preds)
            same = out["e2e"]["faces_observed"] == out["faces_observed"] and out["e2e"]["votes"] == out["votes"]
            out["e2e"]["parity_with_resident_leg"] = "ok" if same else "MISMATCH"  # same views, same images
        except Exception as e:  # noqa: BLE001
            out["e2e"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def run_c5_e2e(args, torch, cfg, verts, faces, c2ws, ids, dev, host_preds):
    """BASELINE config 5 end to end: TexturedPhotogrammetryMeshIndexPredictions.aggregate_projected_images through
    the reference-facing API -- (H, W) uint8 class-index images in page-locked HOST memory (8 distinct ones), one
    vote per visible face and view, results returned as the reference's scipy CSR arrays (built on the device; only
    the CSR arrays cross PCIe).  Timed: the whole call."""
    import geograypher_b200 as gg

    W, H = cfg.image_size
    C, B = cfg.n_classes, 8
    host = []
    for h in host_preds:  # the images of the device-resident leg, now in page-locked host memory
        t = torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
        t.copy_(torch.from_numpy(h))
        host.append(t.numpy())
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=[c2ws[k] for k in ids], intrinsic_params_per_sensor_type=intr)
    seg = gg.SegmentorPhotogrammetryCameraSet(
        cams, gg.ArraySegmentor([host[i % len(host)] for i in range(len(ids))], num_classes=C))
    mesh = gg.TexturedPhotogrammetryMeshIndexPredictions((verts, faces), device=dev.index, views_per_batch=B,
                                                         log_level="WARNING")
    mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(2 * B))), n_classes=C)  # warm-up (uploads the mesh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    avg, info = mesh.aggregate_projected_images(seg, n_classes=C)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    counts, summed = info["projection_counts"], info["summed_projections"]
    d2h = sum(a.data.nbytes + a.indices.nbytes + a.indptr.nbytes for a in (avg, counts, summed))
    votes = int(summed.sum())
    del mesh
    torch.cuda.empty_cache()
    return {"value": len(ids) / dt, "unit": "views/s", "views": len(ids), "seconds": dt,
            "h2d_bytes_per_view": votes / max(len(ids), 1) * 32.0, "d2h_bytes": d2h, "faces_observed": int(counts.nnz),
            "votes": votes, "input": "pinned_index_u8", "distinct_host_images": len(host),
            "api": "TexturedPhotogrammetryMeshIndexPredictions.aggregate_projected_images(SegmentorPhotogrammetryCameraSet, "
                   "n_classes) -> scipy.sparse.csr_array results",
            "note": "class-index images stay in page-locked host memory; the GPU reads one byte (a 32-byte PCIe sector) "
                    "per visible face and view in place; the CSR results (average, counts, sums) are assembled on the "
                    "device and copied once"}


def run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, dev, world, host, kind):
    """aggregate_projected_images through the reference-facing API, prediction images in host memory.
    kind: host_array_f32 ((H,W,C) float32 in gg.host_array memory -- host-resident, mapped into the GPU -- read in
    place), pinned_f32 (the same in page-locked memory), pinned_index_u8 ((H,W) uint8 class-index images, the
    LookUpSegmentor contract, expanded on the GPU), pageable_f32 (ordinary NumPy arrays)."""
    W, H = cfg.image_size
    C = cfg.n_classes
    from geograypher_b200 import distributed as ggd

    n_views = args.e2e_views  # per rank (weak scaling, like the device-resident leg): the rank's cameras, cycled
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
    all_ids = []
    for r in range(world):  # every rank describes the WHOLE job (rank r owning the r-th contiguous block)
        block = shard(len(c2ws), r, world)
        all_ids += [block[i % len(block)] for i in range(n_views)]
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=[c2ws[k] for k in all_ids],
                                      intrinsic_params_per_sensor_type=intr)
    if kind == "pinned_index_u8":
        host = []
        for i in range(16):
            t = torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
            t.copy_(torch.from_numpy(syn.class_index_image(i, H, W, C)))
            host.append(t.numpy())
        segmentor = gg.ArraySegmentor([host[i % len(host)] for i in range(len(all_ids))], num_classes=C, one_hot=True)
        elem_bytes, row_elems = 1, 1
    elif kind == "pinned_f32":
        src = host
        host = []
        for i in range(len(src)):  # the same images in page-locked (cudaHostAlloc) memory
            t = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True)
            t.copy_(torch.from_numpy(src[i]))
            host.append(t.numpy())
        segmentor = gg.ArraySegmentor([host[i % len(host)] for i in range(len(all_ids))], num_classes=C)
        elem_bytes, row_elems = 4, C
    elif kind == "pageable_f32":
        pinned = host
        host = [np.array(pinned[i % len(pinned)], copy=True) for i in range(16)]  # ordinary (pageable) arrays
        for i, h in enumerate(host):
            h[0, 0, 0] += 1e-3 * i  # distinct contents
        segmentor = gg.ArraySegmentor([host[i % len(host)] for i in range(len(all_ids))], num_classes=C)
        elem_bytes, row_elems = 4, C
    else:
        segmentor = gg.ArraySegmentor([host[i % len(host)] for i in range(len(all_ids))], num_classes=C)
        elem_bytes, row_elems = 4, C
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, segmentor)
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), device=dev.index, views_per_batch=args.batch,
                                         log_level="WARNING")
    mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(min(2 * args.batch, n_views)))))  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:  # camera-sharded, one all-reduce, the result is copied to the host once (rank 0)
        avg, info = ggd.aggregate_projected_images_distributed(mesh, seg, dst_rank=0)
    else:
        avg, info = mesh.aggregate_projected_images(seg)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    seen = torch.tensor([float(info["projection_counts"].sum()) if info else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(seen, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    steps = -(-n_views // args.views_per_step)
    F = len(faces)
    seen_per_view = float(seen.item()) / max(n_views * world, 1)
    row_bytes = -(-row_elems * elem_bytes // 32) * 32  # PCIe reads are sector-granular
    if kind in ("pageable_f32", "pinned_f32"):  # (6.4 GB of page-locked images take the host-gather route too)
        h2d = seen_per_view * (row_elems * elem_bytes) * n_views / steps  # the host gathers the rows, then uploads them
    else:
        h2d = seen_per_view * row_bytes * n_views / steps
    del mesh
    torch.cuda.empty_cache()
    notes = {
        "host_array_f32": "float32 (H,W,C) score images stay in HOST memory (geograypher_b200.host_array: managed memory "
                          "whose preferred location is the host, mapped into the GPU); last-pixel aggregation needs one "
                          "row per visible face, which a small kernel fetches over PCIe in place while the next batch is "
                          "rasterized (h2d bytes = rows fetched, sector-granular); the per-face float64 averages, sums and "
                          "counts are copied back at the end",
        "pinned_f32": "the same images in page-locked (cudaHostAlloc) memory, 6.4 GB of them: on this host the "
                      "scattered-read rate from page-locked memory falls with the pinned footprint "
                      "(profiles/r02_pcie_rows_footprint.txt: 284 -> 100 M 40-byte rows/s from 1.3 to 10 GB; read in "
                      "place this leg measured ~3 400 views/s), so above 4 GiB of distinct page-locked images the API "
                      "takes the host-gather route of the pageable leg instead "
                      "(TexturedPhotogrammetryMesh.pinned_direct_limit_bytes)",
        "pinned_index_u8": "(H,W) uint8 class-index images in pinned host memory (what LookUpSegmentor yields before "
                           "its one-hot expansion), expanded on the GPU exactly like Segmentor.inds_to_one_hot; one "
                           "byte per visible face crosses PCIe",
        "pageable_f32": "float32 (H,W,C) score images in ordinary NumPy arrays (16 distinct images): the GPU lists (face, "
                        "last pixel) per view for batch k+1 while the host's cores pick the rows of batch k "
                        "(gg_project_winners || gg_gather_rows_host -> gg_accumulate_rows)"}
    return {"value": n_views * world / dt, "unit": "views/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int((F * C * 8 * 2 + F * 8) / steps), "views": n_views * world, "seconds": dt,
            "input": kind, "distinct_host_images": len(host),
            "api": ("TexturedPhotogrammetryMesh.aggregate_projected_images(SegmentorPhotogrammetryCameraSet)" if world == 1 else
                    "geograypher_b200.distributed.aggregate_projected_images_distributed(mesh, SegmentorPhotogrammetryCameraSet, dst_rank=0)"),
            "note": notes[kind] + ("; cameras sharded over the ranks, accumulators all-reduced (NCCL), result copied to "
                                   "the host by rank 0" if world > 1 else "")}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
