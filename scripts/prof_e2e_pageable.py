"""Where the pageable-image route spends its time (development aid): wall clock per phase of a 500-view aggregation."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn, _lib
from geograypher_b200.meshes import meshes as M
verts, faces, c2ws, cfg = syn.make_survey("c2")
W, H = cfg.image_size; C = cfg.n_classes
dev = torch.device("cuda", 0)
host = [syn.softmax_predictions_device(i, H, W, C, dev).cpu().numpy() for i in range(16)]
n = 500
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[:n], intrinsic_params_per_sensor_type=intr)
seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i % 16] for i in range(n)], num_classes=C))
mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=10, log_level="WARNING")
mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(40))))
T = {}
def timed(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); T[name] = T.get(name, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, g)
timed(_lib, "gather_rows_host"); timed(mesh, "_sparse_begin"); timed(mesh, "_sparse_finish"); timed(mesh, "_fetch_prediction")
timed(mesh, "_gg_cameras"); timed(_lib, "pointer_kind")
ev_sync = torch.cuda.Event.synchronize
def es(self):
    t = time.perf_counter(); ev_sync(self); T["event_sync"] = T.get("event_sync", 0.0) + time.perf_counter() - t
torch.cuda.Event.synchronize = es
ss = torch.cuda.Stream.synchronize
def s2(self):
    t = time.perf_counter(); ss(self); T["stream_sync"] = T.get("stream_sync", 0.0) + time.perf_counter() - t
torch.cuda.Stream.synchronize = s2
for rep in range(2):
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    avg, info = mesh.aggregate_projected_images(seg)
    dt = time.perf_counter() - t0
    print("views/s", round(n / dt), "seconds", round(dt, 4), {k: round(v, 4) for k, v in T.items()})
