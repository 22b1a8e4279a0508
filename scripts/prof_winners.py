"""Host-side cost of the sparse (pageable-array) aggregation path, call by call (development aid)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import geograypher_b200 as gg
from geograypher_b200 import _lib, synthetic as syn
verts, faces, c2ws, cfg = syn.make_survey("c2")
W, H = cfg.image_size; C = cfg.n_classes
dev = torch.device("cuda", 0)
host = [syn.softmax_predictions_device(i, H, W, C, dev).cpu().numpy() for i in range(4)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
intr = {0: dict(f=cfg.f, cx=cfg.cx, cy=cfg.cy, image_width=W, image_height=H, distortion_params={})}
cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=c2ws[:n], intrinsic_params_per_sensor_type=intr)
seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([host[i % 4] for i in range(n)], num_classes=C))
mesh = gg.TexturedPhotogrammetryMesh((verts, faces), views_per_batch=10, log_level="WARNING")
mesh.aggregate_projected_images(seg.get_subset_cameras(list(range(int(sys.argv[2]) if len(sys.argv) > 2 else 10))))
lib = _lib.load()
log = []
def wrap(name):
    fn = getattr(lib, name)
    def inner(*a):
        t0 = time.perf_counter(); r = fn(*a); log.append((name, 1e3 * (time.perf_counter() - t0))); return r
    return inner
class Proxy:
    def __init__(self, lib): self._lib = lib; self._w = {}
    def __getattr__(self, k):
        if k.startswith("gg_"):
            if k not in self._w: self._w[k] = wrap(k)
            return self._w[k]
        return getattr(self._lib, k)
ctx = mesh._get_context(); ctx.lib = Proxy(lib)
orig_sync = torch.cuda.Stream.synchronize
def sync(self):
    t0 = time.perf_counter(); orig_sync(self); log.append(("stream.sync", 1e3 * (time.perf_counter() - t0)))
torch.cuda.Stream.synchronize = sync
t0 = time.perf_counter(); mesh.aggregate_projected_images(seg); dt = time.perf_counter() - t0
print("views/s", n / dt)
import collections
agg = collections.OrderedDict()
for k, v in log: agg.setdefault(k, []).append(v)
for k, v in agg.items(): print(f"{k:28s} n={len(v):4d} total {sum(v):8.2f} ms  max {max(v):7.2f}  first5 {[round(x,2) for x in v[:5]]}")
