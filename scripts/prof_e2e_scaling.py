"""End-to-end time as a function of the number of views (slope = per-batch cost, intercept = fixed cost)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn
verts, faces, c2ws, cfg = syn.make_survey("c2")
W, H = cfg.image_size; C = cfg.n_classes
dev = torch.device("cuda", 0)
host = []
for i in range(4):
    t = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True); t.copy_(syn.softmax_predictions_device(i, H, W, C, dev)); host.append(t.numpy())
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=10, log_level="WARNING")
def make(n):
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[:n], intrinsic_params_per_sensor_type=intr)
    return gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i % 4] for i in range(n)], num_classes=C))
mesh.aggregate_projected_images(make(20))
import geograypher_b200.meshes.meshes as mm
for n in (100, 300, 500, 500):
    seg = make(n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d_sum, d_count, _ = mesh._accumulate_views(seg, 1, mm._lib.MODE_LAST_PIXEL, pix2face_kwargs={})
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx = mesh._get_context()
    avg, _ = ctx.finalize(d_sum, d_count, want_avg=True, want_argmax=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out = mesh._to_host(avg, d_sum, d_count.double())
    t3 = time.perf_counter()
    print(f"n={n}: accumulate {1e3*(t1-t0):.1f} ms, finalize {1e3*(t2-t1):.1f} ms, to_host {1e3*(t3-t2):.1f} ms")
