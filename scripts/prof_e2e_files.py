#!/usr/bin/env python
"""aggregate_projected_images with predictions DECODED from files (LookUpSegmentor: 20-Mpx class-index PNGs / .npy),
with and without the read-ahead of utils/prefetch.py.  Usage: python scripts/prof_e2e_files.py > profiles/r02_e2e_files.txt"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import geograypher_b200 as gg  # noqa: E402
from geograypher_b200 import synthetic as syn  # noqa: E402
from geograypher_b200.utils.prefetch import default_threads  # noqa: E402

verts, faces, c2ws, cfg, origin = bench.build_survey("c2")
W, H = cfg.image_size
C, n, distinct = cfg.n_classes, 96, 8
root = Path(tempfile.mkdtemp(prefix="gg_files_", dir="/dev/shm" if Path("/dev/shm").is_dir() else None))
(root / "imgs").mkdir()
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
names = [root / "imgs" / f"{i % distinct:03d}.JPG" for i in range(n)]  # 96 views cycle through 8 prediction files
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=[c2ws[k] for k in range(n)], image_filenames=names,
                                  image_folder=root / "imgs", intrinsic_params_per_sensor_type=intr)
print(f"c2: {len(faces)} faces, {n} views {W}x{H}, {C} classes, uint8 class-index predictions; host threads available: "
      f"{default_threads()} used for read-ahead")
for fmt in ("png", "npy"):
    folder = root / f"preds_{fmt}"
    folder.mkdir()
    t0 = time.perf_counter()
    for i in range(distinct):
        img = syn.class_index_image(i, H, W, C, ignore_frac=0.0)
        if fmt == "png":
            from PIL import Image

            Image.fromarray(img).save(folder / f"{i:03d}.png", compress_level=1)
        else:
            np.save(folder / f"{i:03d}.npy", img)
    size = sum(p.stat().st_size for p in folder.iterdir()) / distinct / 1e6
    print(f"-- {fmt}: {size:.1f} MB per file (written in {time.perf_counter() - t0:.1f} s)")
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.LookUpSegmentor(root / "imgs", folder, num_classes=C))
    ref = None
    for threads in (0, None):
        mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=8, prefetch_threads=threads, log_level="WARNING")
        mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(16))))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        avg, info = mesh.aggregate_projected_images(seg)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = "" if ref is None else f"  same result: {bool(np.array_equal(np.nan_to_num(avg), np.nan_to_num(ref)))}"
        ref = avg if ref is None else ref
        print(f"   read-ahead {'off' if threads == 0 else 'on '}: {n / dt:8.1f} views/s ({1e3 * dt / n:6.1f} ms per view){same}")
        del mesh
import shutil

shutil.rmtree(root, ignore_errors=True)
