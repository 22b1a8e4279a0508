#!/usr/bin/env python
"""Where the c5 end-to-end leg (vote aggregation through the API, CSR results) spends its time: the accumulation of the
views (row fetch from page-locked class-index images) and the assembly + delivery of the CSR results, per step of
_votes_to_csr.  Usage: python scripts/prof_c5_e2e.py > profiles/r02_e2e_c5.txt"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import geograypher_b200 as gg  # noqa: E402
from geograypher_b200 import _lib, synthetic as syn  # noqa: E402

verts, faces, c2ws, cfg, origin = bench.build_survey("c5")
W, H = cfg.image_size
C, B, n = cfg.n_classes, 8, 240
host = []
for i in range(B):
    t = torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
    t.copy_(torch.from_numpy(syn.class_index_image(i, H, W, C)))
    host.append(t.numpy())
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=[c2ws[k] for k in range(n)], intrinsic_params_per_sensor_type=intr)
seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i % B] for i in range(n)], num_classes=C))
mesh = gg.TexturedPhotogrammetryMeshIndexPredictions((verts, faces), device=0, views_per_batch=B, log_level="WARNING")
mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(2 * B))), n_classes=C)


def clock(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label:44s} {1e3 * (time.perf_counter() - t0):8.1f} ms")
    return out


for rep in range(2):
    print(f"-- call {rep}")
    d_sum, d_count, _ = clock("accumulate 240 views (rows over PCIe)", lambda: mesh._accumulate_views(seg, 1, _lib.MODE_VOTE, n_channels=C, pix2face_kwargs={}))
    res = clock("_votes_to_csr (device assembly + delivery)", lambda: mesh._votes_to_csr(d_sum, d_count))
    nz = clock("  non-zero pattern + nonzero()", lambda: (d_sum != 0).nonzero(as_tuple=True))
    big = clock("  1 GB device -> pageable host", lambda: torch.empty(2**27, dtype=torch.float64, device="cuda").cpu())
    print("   nnz", res[0].nnz, "observed faces", res[1].nnz, "bytes", sum(a.data.nbytes + a.indices.nbytes + a.indptr.nbytes for a in res))
    del d_sum, d_count, res, nz, big
    clock("whole call", lambda: mesh.aggregate_projected_images(seg, n_classes=C))
