"""GPU sanity + timing of the other BASELINE configs: c4 (render_flat to uint8 label rasters on c2 geometry) and
c5 (20M faces, 8192x5460 rig cameras, vote aggregation).  Usage: python scripts/check_configs.py [c4] [c5]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from geograypher_b200 import _lib, synthetic as syn

def build(name, ncam):
    t = time.time()
    verts, faces, c2ws, cfg = syn.make_survey(name, ncam)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    W, H = cfg.image_size
    ctx = _lib.Context(0)
    ctx.set_mesh(torch.from_numpy((verts - origin).astype(np.float32)).cuda(), torch.from_numpy(faces).cuda())
    cams = [_lib.make_camera(np.linalg.inv(T), cfg.f, cfg.cx, cfg.cy, W, H, origin=origin) for T in c2ws]
    print(f"{name}: F={len(faces)} V={len(verts)} cams={len(cams)} {W}x{H} built in {time.time()-t:.1f}s", flush=True)
    return verts, faces, cfg, ctx, cams

def timeit(fn, n, ctx):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); [fn(i) for i in range(n)]
    ctx.drain()  # gg_project_aggregate works on the library's own streams: make this stream wait for them
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if "c4" in sys.argv:
    verts, faces, cfg, ctx, cams = build("c2", 40)
    W, H = cfg.image_size
    tex = torch.from_numpy(syn.voronoi_face_labels(verts, faces)).cuda()
    B = 10
    p2f = torch.empty((B, H, W), dtype=torch.int32, device="cuda")
    out = torch.empty((B, H, W, 1), dtype=torch.uint8, device="cuda")
    def step(i):
        ctx.rasterize(cams[(i % 4) * B:(i % 4) * B + B], out=p2f, check=False)
        ctx.render_flat(p2f, tex, out_dtype=_lib.OUT_U8, out=out)
    step(0); ctx.sync()
    ms = timeit(step, 8, ctx)
    ctx.sync()
    ref = out.clone()
    print(f"c4 render_flat -> uint8 (two kernels): {ms/B*1e3:.1f} us/view, {B/ms*1e3:.0f} views/s; labelled px frac {(out>0).float().mean().item():.3f}")
    def fused(i):
        ctx.rasterize_render_flat(cams[(i % 4) * B:(i % 4) * B + B], tex, out_dtype=_lib.OUT_U8, out=out, check=False)
    fused(7); ctx.sync()
    assert torch.equal(out, ref)
    ms = timeit(fused, 8, ctx)
    ctx.sync()
    print(f"c4 render_flat -> uint8 (fused): {ms/B*1e3:.1f} us/view, {B/ms*1e3:.0f} views/s")

if "c5" in sys.argv:
    verts, faces, cfg, ctx, cams = build("c5", 40)
    W, H = cfg.image_size
    F, C, B = len(faces), cfg.n_classes, 4
    preds = [torch.from_numpy(syn.class_index_image(i, H, W, C)).cuda() for i in range(B)]
    d_sum = torch.zeros((F, C), dtype=torch.float64, device="cuda"); d_count = torch.zeros((F,), dtype=torch.int32, device="cuda")
    def step(i):
        ctx.project_aggregate(cams[(i % 10) * B:(i % 10) * B + B], preds, _lib.PRED_INDEX_U8, C, _lib.MODE_LAST_PIXEL, 0, d_sum, d_count, check=False)
    step(0); ctx.sync(); print("stats", ctx.last_batch_stats(B).tolist())
    d_sum.zero_(); d_count.zero_()
    ms = timeit(step, 10, ctx)
    ctx.sync()
    ctx.profile(True)
    for i in range(10):
        step(i)
    ctx.sync()
    print("c5 stage ms per batch of", B, "views:", {k: round(v[0] / 10, 3) for k, v in ctx.profile_read().items() if v[1]})
    ctx.profile(False)
    avg, amax = ctx.finalize(d_sum, d_count)
    print(f"c5 one-hot aggregation: {ms/B*1e3:.1f} us/view, {B/ms*1e3:.0f} views/s, {B/ms*1e3*W*H/1e9:.1f} Gpix/s; faces observed {(d_count>0).sum().item()}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB torch")
    # oracle spot check on one oblique rig camera
    from oracle import oracle as ora
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    k = 3
    T = syn.lawnmower_cameras(cfg, cfg.n_cells * cfg.cell_m)[k]
    oc = ora.make_camera(T, cfg.f, cfg.cx, cfg.cy, W, H, origin=origin)
    t = time.time(); ref, _, margin = ora.rasterize(v32, faces, oc, want_depth=True, want_margin=True); print(f"oracle {time.time()-t:.1f}s")
    got = ctx.rasterize([cams[k]]).cpu().numpy()[0]
    diff = got != ref
    print("c5 view", k, "diff px", int(diff.sum()), "outside mask", int((diff & (margin > 1e-5)).sum()), "covered frac", float((ref >= 0).mean()))
