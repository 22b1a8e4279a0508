#!/usr/bin/env python
"""bench.py -- views/s and Mpix/s of pix2face + aggregate_projected_images on synthetic surveys.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2] [--mode last_pixel]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one batch of `--views-per-step` views of the configuration through the hot path: camera projection +
tiled z-buffer rasterization (pix2face) + per-face aggregation of the views' class-score images.  The mesh is
replicated on every GPU, cameras are sharded by contiguous blocks, and the per-face float64 sums / int32 counts of
all ranks are combined with one NCCL all-reduce at the end of the timed region, followed by the mean + argmax
epilogue (gg_finalize).  Rank 0 prints ONE JSON line.

`value`      device-resident: prediction images already in HBM (a ring of distinct buffers, > L2 in total).
`e2e`        the same metric through the reference-facing Python API (TexturedPhotogrammetryMesh.
             aggregate_projected_images on a SegmentorPhotogrammetryCameraSet) with the prediction images in pinned
             HOST memory: every view's H2D copy and the final D2H of the per-face averages are inside the timing.
`roofline`   for the dominant kernel (k_raster_tiles), timed with CUDA events on its own stream inside the library.
`cpu_baseline` / `--impl reference`  the CPU restatement of the reference's path (oracle/: OpenMP C rasterizer on all
             host cores + the reference's literal single-threaded NumPy aggregation) on a bounded sample.
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c5", "tiny"])
    ap.add_argument("--mode", default="last_pixel", choices=["last_pixel", "pixel_sum"])
    ap.add_argument("--views-per-step", type=int, default=10)
    ap.add_argument("--e2e-views", type=int, default=500, help="views per rank in the end-to-end (host buffer) run")
    ap.add_argument("--cpu-views", type=int, default=6, help="views in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--single-mode", action="store_true", help="time only --mode (skip the other aggregation mode)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

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
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
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


def measured_traffic(config, mode_name, views_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from a committed `ncu --set full` capture of
    this command (profiles/traffic.json), per launch; None when no capture matches."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text()).get(f"{config}:{mode_name}:{views_per_launch}")
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


def build_survey(name):
    from geograypher_b200 import synthetic as syn

    verts, faces, c2ws, cfg = syn.make_survey(name)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    return verts, faces, c2ws, cfg, origin


def shard(n_items, rank, world):
    per = -(-n_items // world)
    return list(range(rank * per, min(n_items, (rank + 1) * per)))


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


def cpu_views_per_second(verts, faces, c2ws, cfg, origin, view_ids, preds, on_step=None):
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
        if on_step:
            on_step(j)
    return per_view, t_r, t_a


def run_reference(args):
    """--impl reference: the CPU restatement, one view per step, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as ora

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
    sample = (f"{len(timed)} views of {args.config} (1 view per step), OpenMP C rasterizer on {cores} threads + "
              f"single-threaded NumPy aggregation as in the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, 1),
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample,
                         "raster_s_per_view": t_r / n, "aggregate_s_per_view": t_a / n},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mpix_per_s": value * W * H / 1e6, "gpu_launches": 0,
        "note": "reference = CPU restatement of geograypher's path (oracle/): the reference's own rasterizer is "
                "un-vendored VTK/PyTorch3D and cannot be installed offline",
    }
    print(json.dumps(line))


def workload_config(args, cfg, views_per_step):
    W, H = cfg.image_size
    return {
        "workload": f"{cfg.name}: {cfg.n_faces} faces, {cfg.n_cameras} cameras {W}x{H}, {cfg.n_classes}-class "
                    f"float32 softmax scores",
        "mode": args.mode, "views_per_step": views_per_step, "pixels_per_view": W * H,
        "cache": "inputs larger than L2 (ring of distinct prediction buffers, >= 0.8 GB each)",
    }


# ----------------------------------------------------------------------------------------------------------
# GPU path
# ----------------------------------------------------------------------------------------------------------
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

    verts, faces, c2ws, cfg, origin = build_survey(args.config)
    W, H = cfg.image_size
    C, F, P = cfg.n_classes, len(faces), W * H
    B = args.views_per_step
    mode = _lib.MODE_LAST_PIXEL

    my_cams = shard(len(c2ws), rank, world)
    # The page-locked host images of the end-to-end leg are allocated first, while the host's memory is still
    # unfragmented: the GPU reads scattered rows out of them over PCIe, and that runs measurably slower from pinned
    # buffers that were allocated late in a process (after the legs below) than from these.
    e2e_host = None
    if not args.no_e2e:
        e2e_host = []
        for i in range(min(4, args.e2e_views)):
            t = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True)
            t.copy_(syn.softmax_predictions_device(my_cams[i % len(my_cams)], H, W, C, dev))
            e2e_host.append(t.numpy())
        torch.cuda.synchronize()
    ctx = _lib.Context(local_rank)
    ctx.set_mesh(torch.from_numpy((verts - origin).astype(np.float32)).to(dev), torch.from_numpy(faces).to(dev))
    w2c = [np.linalg.inv(T) for T in c2ws]
    gg_cams = {k: _lib.make_camera(w2c[k], cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for k in my_cams}

    # prediction ring, resident in HBM, generated outside the timed region
    ring = [syn.softmax_predictions_device(my_cams[i % len(my_cams)], H, W, C, dev) for i in range(B)]
    d_sum = torch.zeros((F, C), dtype=torch.float64, device=dev)
    d_count = torch.zeros((F,), dtype=torch.int32, device=dev)

    def step(i, check=False):
        ids = [my_cams[(i * B + j) % len(my_cams)] for j in range(B)]
        ctx.project_aggregate([gg_cams[k] for k in ids], ring, _lib.PRED_F32, C, mode, 0, d_sum, d_count, check=check)

    def epilogue():
        ctx.drain()  # the accumulators are written on the library's internal streams
        if world > 1:
            dist.all_reduce(d_sum)
            dist.all_reduce(d_count)
        return ctx.finalize(d_sum, d_count)

    peak, peak_src = measured_peak_gbs()

    def timed_run(run_mode):
        """W warm-up steps, then K timed steps + all-reduce + finalize, bracketed by barrier + synchronize; the
        time is the max over ranks.  Returns a dict with the numbers of this mode."""
        nonlocal mode
        mode = run_mode
        d_sum.zero_()
        d_count.zero_()
        for i in range(args.warmup):
            step(i, check=True)  # a scratch overflow grows the scratch and replays; the timed steps do not check
        ctx.sync()
        stats = ctx.last_batch_stats(B)
        d_sum.zero_()
        d_count.zero_()
        ctx.profile(True)
        ctx.profile_read(reset=True)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            step(args.warmup + i)
        epilogue()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        elapsed_s = float(elapsed_ms.item()) / 1e3
        clocks = sampler.stop() if rank == 0 else None
        ctx.sync()  # surfaces a scratch overflow of the last batch
        prof = ctx.profile_read(reset=True)
        ctx.profile(False)
        views = args.steps * B * world
        raster_ms, raster_launches = prof["raster_tiles"]
        f_v = float(stats[:, 1].mean())
        # algorithmic bytes per view.  Stage 1+2: 12 V_v + 12 F_v (+ 4 P only when the raster is written,
        # which the fused paths do not do).  Dense stage 3: s*C*P scores + read-modify-write of the float64 sums and
        # int32 counts of the touched faces.
        b12 = 12.0 * (f_v / 2.0) + 12.0 * f_v
        if run_mode == _lib.MODE_PIXEL_SUM:
            # only pixels that hit the mesh contribute scores: count them from the per-face pixel counts
            px_added = float(d_count.sum().item()) / (args.steps * B * world)
            bytes_per_view = b12 + 4.0 * C * px_added + (16.0 * C + 8.0) * f_v
        else:
            bytes_per_view = b12 + 4.0 * P  # B12 of SURVEY 8d: the figure for pix2face, IDs written once
        avg_ms = raster_ms / max(raster_launches, 1)
        achieved = bytes_per_view * B / (avg_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
        mode_name = "pixel_sum" if run_mode == _lib.MODE_PIXEL_SUM else "last_pixel"
        roofline = {"bound": "hbm", "kernel": "k_raster_tiles", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": measured_traffic(args.config, mode_name, B),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_per_view * B, "avg_launch_ms": avg_ms}
        return {"value": views / elapsed_s, "elapsed_s": elapsed_s, "roofline": roofline, "clocks": clocks,
                "stage_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0},
                "launches": int(sum(v[1] for v in prof.values())), "faces_per_view": f_v,
                "faces_observed": int((d_count > 0).sum().item())}

    headline_mode = {"last_pixel": _lib.MODE_LAST_PIXEL, "pixel_sum": _lib.MODE_PIXEL_SUM}[args.mode]
    other_mode = _lib.MODE_PIXEL_SUM if headline_mode == _lib.MODE_LAST_PIXEL else _lib.MODE_LAST_PIXEL
    other = timed_run(other_mode) if not args.single_mode else None
    main = timed_run(headline_mode)
    value, elapsed_s, roofline, clocks = main["value"], main["elapsed_s"], main["roofline"], main["clocks"]
    stage_ms, launches, f_v, observed = main["stage_ms"], main["launches"], main["faces_per_view"], main["faces_observed"]
    if headline_mode == _lib.MODE_LAST_PIXEL:
        roofline["note"] = ("reference-parity (last_pixel) aggregation needs ~0.1 GB of HBM traffic per 20-Mpx view: the "
                            "rasterizer is instruction-issue bound, not HBM bound (DESIGN.md section 5); the dense "
                            "pixel_sum mode, which streams every score, is reported under 'pixel_sum'")

    # ---- end to end through the public API with host buffers ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, my_cams, dev, world, e2e_host)

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as ora

        ids = [(7 + 3 * i) % len(c2ws) for i in range(args.cpu_views)]
        per_view, t_r, t_a = cpu_views_per_second(verts, faces, c2ws, cfg, origin, ids, cpu_predictions(cfg))
        cpu = {"value": len(per_view) / float(sum(per_view)), "unit": "views/s", "cores": host_threads(),
               "kind": "port",
               "sample": f"{len(per_view)} views of {args.config}; OpenMP C rasterizer on all cores + single-threaded "
                         f"NumPy aggregation (the reference's aggregation is single-threaded NumPy)",
               "raster_s_per_view": t_r / len(per_view), "aggregate_s_per_view": t_a / len(per_view)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, B),
            "mpix_per_s": value * P / 1e6, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "stage_ms": stage_ms,
            "faces_per_view": f_v, "faces_observed": observed,
            ("pixel_sum" if headline_mode == _lib.MODE_LAST_PIXEL else "last_pixel"): None if other is None else {
                "value": other["value"], "unit": "views/s", "mpix_per_s": other["value"] * P / 1e6,
                "ms_per_step": 1e3 * other["elapsed_s"] / args.steps, "roofline": other["roofline"],
                "stage_ms": other["stage_ms"],
                "note": "same workload with every pixel adding its scores (GG_MODE_PIXEL_SUM, not the reference's "
                        "semantics): the fused rasterizer epilogue streams the (H,W,C) float32 scores from HBM"
                        if headline_mode == _lib.MODE_LAST_PIXEL else "reference-parity mode"},
            "accumulators": "float64 sums + int32 counts; one NCCL all-reduce at the end" if world > 1 else
                            "float64 sums + int32 counts",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, gg, syn, torch, dist, cfg, verts, faces, c2ws, my_cams, dev, world, host):
    """aggregate_projected_images through the reference-facing API, prediction images in pinned host memory."""
    W, H = cfg.image_size
    C = cfg.n_classes
    from geograypher_b200 import distributed as ggd

    n_views = args.e2e_views  # per rank (weak scaling, like the device-resident leg): the rank's cameras, cycled
    n_host = len(host)
    intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
    # every rank describes the WHOLE job (world x n_views cameras, rank r owning the r-th contiguous block)
    all_ids = []
    for r in range(world):
        block = shard(len(c2ws), r, world)
        all_ids += [block[i % len(block)] for i in range(n_views)]
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=[c2ws[k] for k in all_ids],
                                      intrinsic_params_per_sensor_type=intr)
    seg = gg.SegmentorPhotogrammetryCameraSet(
        cams, gg.ArraySegmentor([host[i % n_host] for i in range(len(all_ids))], num_classes=C))
    mesh = gg.TexturedPhotogrammetryMesh((verts, faces), device=dev.index, views_per_batch=args.views_per_step,
                                         log_level="WARNING")
    mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(min(2, n_views)))))  # warm-up: mesh upload etc.
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:  # camera-sharded, one all-reduce, the result is copied to the host once (rank 0)
        timings = {} if os.environ.get("GG_BENCH_DEBUG") else None
        avg, info = ggd.aggregate_projected_images_distributed(mesh, seg, dst_rank=0, timings=timings)
        if timings is not None:
            print(f"[e2e rank {dist.get_rank()}] " + ", ".join(f"{k} {v:.3f}s" for k, v in timings.items()), file=sys.stderr, flush=True)
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
    zero_copy = all(torch.from_numpy(h).is_pinned() for h in host)
    seen_per_view = float(seen.item()) / max(n_views * world, 1)
    row_bytes = -(-C * 4 // 32) * 32  # PCIe reads are sector-granular
    h2d = seen_per_view * row_bytes * n_views / steps if zero_copy else H * W * C * 4 * n_views / steps
    return {"value": n_views * world / dt, "unit": "views/s",
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int((F * C * 8 * 2 + F * 8) / steps),
            "views": n_views * world, "seconds": dt,
            "api": ("TexturedPhotogrammetryMesh.aggregate_projected_images(SegmentorPhotogrammetryCameraSet)" if world == 1 else
                    "geograypher_b200.distributed.aggregate_projected_images_distributed(mesh, SegmentorPhotogrammetryCameraSet, dst_rank=0)"),
            "note": ("float32 (H,W,C) score images stay in pinned HOST memory; last-pixel aggregation needs one row per "
                     "visible face, which a staging kernel fetches over PCIe through unified addressing, all rows of a batch in parallel (h2d bytes = "
                     "rows actually fetched, estimated from the per-face counts); the per-face float64 averages, sums "
                     "and counts are copied back at the end" if zero_copy else
                     "float32 (H,W,C) score images uploaded from host memory every view") +
                    ("; cameras sharded over the ranks, accumulators all-reduced (NCCL), result copied to the host by "
                     "rank 0" if world > 1 else "")}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
