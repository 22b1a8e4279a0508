import sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn
verts, faces, c2ws, cfg = syn.make_survey("c2")
W,H = cfg.image_size; C = cfg.n_classes
dev = torch.device("cuda",0)
# usage: prof_e2e.py [pinned|pageable|upload] [n_views]
where = sys.argv[1] if len(sys.argv) > 1 else "pinned"
host=[]
for i in range(4):
    t = torch.empty((H,W,C), dtype=torch.float32, pin_memory=(where == "pinned")); t.copy_(syn.softmax_predictions_device(i,H,W,C,dev)); host.append(t.numpy())
torch.cuda.synchronize()
n=int(sys.argv[2]) if len(sys.argv) > 2 else 500
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[:n], intrinsic_params_per_sensor_type=intr)
seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i%4] for i in range(n)], num_classes=C))
mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=10, log_level="WARNING", sparse_host_gather=(where != "upload"))
mesh.aggregate_projected_images(seg.get_subset_cameras([0,1]))
torch.cuda.synchronize()
t0=time.perf_counter()
pr = cProfile.Profile(); pr.enable()
avg, info = mesh.aggregate_projected_images(seg)
torch.cuda.synchronize()
pr.disable()
dt=time.perf_counter()-t0
print(where, "views/s", n/dt, "total s", dt)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
