"""Where the end-to-end (host_array float32 predictions) aggregation of 500 c2 views spends its time: wall clock of the
three phases of the public call (accumulate / finalize / results to host) and per-stage GPU time with the software
pipeline on and off (development aid; profiles/r02_e2e_host_array.txt)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import geograypher_b200 as gg
from geograypher_b200 import synthetic as syn
verts, faces, c2ws, cfg = syn.make_survey("c2")
W, H = cfg.image_size; C = cfg.n_classes
dev = torch.device("cuda", 0)
host = []
for i in range(8):
    a = gg.host_array((H, W, C), np.float32)
    t = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True); t.copy_(syn.softmax_predictions_device(i, H, W, C, dev))
    a[...] = t.numpy(); del t; host.append(a)
n = 500
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[:n], intrinsic_params_per_sensor_type=intr)
seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i % 8] for i in range(n)], num_classes=C))
mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=10, log_level="WARNING")
mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(40))))
ctx = mesh._get_context()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    avg, info = mesh.aggregate_projected_images(seg)
    dt = time.perf_counter() - t0
    print("plain call", rep, "views/s", round(n / dt), "seconds", round(dt, 4))
# phases, with a synchronisation between them (adds the pipeline drain to the first)
acc, fin, toh = mesh._accumulate_views, ctx.finalize, mesh._to_host
T = {}
def timed(name, f):
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); T[name + "_enqueue"] = time.perf_counter() - t
        torch.cuda.synchronize(); T[name] = time.perf_counter() - t
        return r
    return g
mesh._accumulate_views = timed("accumulate", acc); ctx.finalize = timed("finalize", fin); mesh._to_host = timed("to_host", toh)
t0 = time.perf_counter(); mesh.aggregate_projected_images(seg); print("phases (s)", {k: round(v, 4) for k, v in T.items()}, "total", round(time.perf_counter() - t0, 4))
mesh._accumulate_views, ctx.finalize, mesh._to_host = acc, fin, toh
for pipe in (True, False):
    ctx.set_pipeline(pipe)
    ctx.profile(True); ctx.profile_read()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    avg, info = mesh.aggregate_projected_images(seg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pipeline", pipe, "views/s", round(n / dt), "rows/view", info["projection_counts"].sum() / n)
    print({k: (round(v[0], 2), v[1]) for k, v in ctx.profile_read().items() if v[1]})
