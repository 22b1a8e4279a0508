"""Context-level loop with pinned host predictions: GPU time per batch without the Python API around it."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from geograypher_b200 import _lib, synthetic as syn
verts, faces, c2ws, cfg = syn.make_survey("c2")
origin = 0.5 * (verts.min(0) + verts.max(0)); W, H = cfg.image_size; C = cfg.n_classes; F = len(faces)
dev = torch.device("cuda", 0)
host = []
for i in range(4):
    t = torch.empty((H, W, C), dtype=torch.float32, pin_memory=True); t.copy_(syn.softmax_predictions_device(i, H, W, C, dev)); host.append(t)
ctx = _lib.Context(0)
ctx.set_mesh(torch.from_numpy((verts - origin).astype(np.float32)).cuda(), torch.from_numpy(faces).cuda())
cams = [_lib.make_camera(np.linalg.inv(T), cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for T in c2ws]
d_sum = torch.zeros((F, C), dtype=torch.float64, device=dev); d_count = torch.zeros((F,), dtype=torch.int32, device=dev)
B = 10
def run(nb, pipeline=True):
    ctx.set_pipeline(pipeline)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for b in range(nb):
        ctx.project_aggregate(cams[b * B:(b + 1) * B], [host[(b * B + j) % 4] for j in range(B)], _lib.PRED_F32, C, _lib.MODE_LAST_PIXEL, 0, d_sum, d_count, check=False)
    t1 = time.perf_counter()
    ctx.sync(); t2 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t0)
run(5)
for nb in (10, 30, 50, 50):
    enq, tot = run(nb)
    print(f"pipelined  {nb} batches: enqueue {enq:.1f} ms, total {tot:.1f} ms = {tot/nb:.2f} ms/batch")
enq, tot = run(30, False)
print(f"serial     30 batches: enqueue {enq:.1f} ms, total {tot:.1f} ms = {tot/30:.2f} ms/batch")
ctx.set_pipeline(True)
ctx.profile(True); run(30); print({k: round(v[0] / 30, 3) for k, v in ctx.profile_read().items() if v[1]})
