"""Per-stage GPU time of the end-to-end (pinned host predictions) aggregation with the software pipeline OFF: every
kernel runs alone, so the stage times are the kernels' own (development aid)."""
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
n = 200
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[100:100 + n], intrinsic_params_per_sensor_type=intr)
seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i % 4] for i in range(n)], num_classes=C))
mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=10, log_level="WARNING")
mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(20))))
ctx = mesh._get_context()
for pipe in (True, False):
    ctx.set_pipeline(pipe)
    ctx.profile(True); ctx.profile_read()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    avg, info = mesh.aggregate_projected_images(seg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pipeline", pipe, "views/s", n / dt, "rows/view", info["projection_counts"].sum() / n)
    print({k: (round(v[0], 2), v[1]) for k, v in ctx.profile_read().items() if v[1]})
