"""Generate tests/golden/*.npz by running the UNMODIFIED reference's own code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference's rasterizer (VTK through pyvista) cannot run here, so ``pix2face`` is supplied through the
reference's own plug-in point -- a subclass that overrides ``pix2face`` (like
derived_meshes.py:642 does) and returns rasters made by oracle/oracle_raster.c.  Everything downstream of
``pix2face`` is the reference's real code:

* ``TexturedPhotogrammetryMesh.project_images`` / ``aggregate_projected_images``  (meshes.py:1944-2084)
* ``TexturedPhotogrammetryMeshIndexPredictions.aggregate_projected_images``       (derived_meshes.py:415-550)
* ``TexturedPhotogrammetryMesh.render_flat``                                       (meshes.py:1858-1942)
* ``SegmentorPhotogrammetryCameraSet`` + ``Segmentor.inds_to_one_hot``           (cameras/segmentor.py, predictors/segmentor.py)
* ``PhotogrammetryCamera`` / ``PhotogrammetryCameraSet``                          (cameras/cameras.py)
* ``find_argmax_nonzero_value``                                                    (utils/indexing.py:9-32)

The inputs are stored next to the outputs so the tests do not depend on the generators staying fixed.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as ora  # noqa: E402
from oracle.ref_stub import import_reference  # noqa: E402
from geograypher_b200 import synthetic as syn  # noqa: E402

import_reference()
from geograypher.cameras import (  # noqa: E402
    PhotogrammetryCamera,
    PhotogrammetryCameraSet,
    SegmentorPhotogrammetryCameraSet,
)
from geograypher.meshes.derived_meshes import TexturedPhotogrammetryMeshIndexPredictions  # noqa: E402
from geograypher.meshes.meshes import TexturedPhotogrammetryMesh  # noqa: E402
from geograypher.predictors import Segmentor  # noqa: E402
from geograypher.utils.indexing import find_argmax_nonzero_value  # noqa: E402

OUT = Path(__file__).resolve().parent


def _cam_index(camera):
    return int(Path(camera.image_filename).stem)


class _PlugIn:
    """The three members that need pyvista in the reference, replaced; nothing else is touched."""

    def _golden_init(self, n_faces, pix2face_stack, face_texture=None):
        self.faces = np.zeros((n_faces, 3), dtype=int)
        self._p2f = pix2face_stack
        self._face_texture = face_texture

    def get_mesh_in_cameras_coords(self, cameras, inplace=False):
        return None

    def pix2face(self, cameras, mesh=None, render_img_scale=1, **kwargs):
        if isinstance(cameras, PhotogrammetryCamera):
            return self._p2f[_cam_index(cameras)]
        return np.stack([self._p2f[_cam_index(c)] for c in cameras.cameras], axis=0)

    def get_texture(self, request_vertex_texture=None, try_verts_faces_conversion=True):
        return self._face_texture


class GoldenMesh(_PlugIn, TexturedPhotogrammetryMesh):
    def __init__(self, *a, **k):
        self._golden_init(*a, **k)


class GoldenIndexMesh(_PlugIn, TexturedPhotogrammetryMeshIndexPredictions):
    def __init__(self, *a, **k):
        self._golden_init(*a, **k)


class MemorySegmentor(Segmentor):
    """Returns in-memory predictions keyed by camera index; one-hot encoding is the reference's."""

    def __init__(self, images, num_classes=None, one_hot=False):
        super().__init__(num_classes=num_classes)
        self.images = images
        self.one_hot = one_hot

    def segment_image(self, image, filename=None, image_scale=1.0, **kwargs):
        img = self.images[int(Path(filename).stem)]
        if self.one_hot:
            return self.inds_to_one_hot(img, num_classes=self.num_classes)
        return img


def build_scene():
    n_cells, W, H, f, cx, cy = 24, 96, 72, 70.0, 1.5, -2.25
    verts, faces = syn.terrain_mesh(n_cells, 1.0, seed=7, crowns=True)
    cfg = syn.SurveyConfig("golden", n_cells, 1.0, 7, True, (1, 3), (7.0, 0.0), (W, H), f, cx, cy, 30.0, 17, 4)
    c2ws = syn.lawnmower_cameras(cfg, n_cells * 1.0, jitter_deg=6.0)
    origin = 0.5 * (verts.min(0) + verts.max(0))
    v32 = (verts - origin).astype(np.float32)
    cams = [ora.make_camera(T, f, cx, cy, W, H, origin=origin) for T in c2ws]
    p2f = ora.pix2face_set(v32, faces, cams)
    return dict(verts=verts, faces=faces, c2ws=np.stack(c2ws), W=W, H=H, f=f, cx=cx, cy=cy, origin=origin,
                pix2face=p2f, n_classes=4)


def main():
    sc = build_scene()
    F = sc["faces"].shape[0]
    n, H, W, C = sc["c2ws"].shape[0], sc["H"], sc["W"], sc["n_classes"]
    p2f = sc["pix2face"]
    assert (p2f == -1).any() and (p2f >= 0).any()

    ref_cams = PhotogrammetryCameraSet(
        cameras=[
            PhotogrammetryCamera(f"/golden/{i:04d}.png", sc["c2ws"][i], sc["f"], sc["cx"], sc["cy"], W, H,
                                 local_to_epsg_4978_transform=np.eye(4))
            for i in range(n)
        ],
        local_to_epsg_4978_transform=np.eye(4),
    )

    # --- camera container pins ------------------------------------------------------------------------
    cam0 = ref_cams[0]
    cam_pins = dict(
        world_to_cam=np.stack([c.world_to_cam_transform for c in ref_cams.cameras]),
        size_s1=np.array(cam0.get_image_size(1.0)),
        size_s07=np.array(cam0.get_image_size(0.7)),
        size_s05=np.array(cam0.get_image_size(0.5)),
        n_channels=np.array(ref_cams.n_image_channels()),
    )

    # --- (1) one-hot class-index predictions through the real Segmentor camera set ------------------
    idx_imgs = [syn.class_index_image(i, H, W, C, block=8, ignore_frac=0.05, seed=3) for i in range(n)]
    seg_set = SegmentorPhotogrammetryCameraSet(ref_cams, MemorySegmentor(idx_imgs, num_classes=C, one_hot=True))
    mesh = GoldenMesh(F, p2f)
    avg1, info1 = mesh.aggregate_projected_images(seg_set)
    onehot0 = seg_set.get_image_by_index(0)

    # --- (2) float softmax predictions with NaN holes -------------------------------------------------
    rng = np.random.default_rng(5)
    soft = [syn.softmax_predictions(i, H, W, C, grid=(5, 7)) for i in range(n)]
    for s in soft:
        s[rng.random((H, W)) < 0.03] = np.nan  # whole-pixel nulls
        s[rng.random((H, W)) < 0.01, 1] = np.nan  # single-channel nulls
    soft[1][:] = np.nan  # a view with no finite pixel at all
    seg_set2 = SegmentorPhotogrammetryCameraSet(ref_cams, MemorySegmentor(soft, num_classes=C))
    avg2, info2 = mesh.aggregate_projected_images(seg_set2)
    projs2 = list(mesh.project_images(seg_set2))
    # single view: the "first projection keeps its NaNs" corner of meshes.py:2056-2057
    avg2s, info2s = mesh.aggregate_projected_images(seg_set2.get_subset_cameras([0]))

    # --- (3) one-hot votes (IndexPredictions) ---------------------------------------------------------
    n_vote_classes = 6
    vote_imgs = []
    for i in range(n):
        v = syn.class_index_image(i, H, W, n_vote_classes, block=6, ignore_frac=0.0, seed=9).astype(np.float64)
        v[rng.random((H, W)) < 0.4] = np.nan
        vote_imgs.append(v)
    vote_imgs[2][:] = np.nan
    seg_set3 = SegmentorPhotogrammetryCameraSet(ref_cams, MemorySegmentor(vote_imgs))
    imesh = GoldenIndexMesh(F, p2f)
    avg3, info3 = imesh.aggregate_projected_images(seg_set3, n_classes=n_vote_classes)

    # --- (4) render_flat ----------------------------------------------------------------------------------
    tex1 = syn.voronoi_face_labels(sc["verts"], sc["faces"], n_sites=12, n_classes=5, nan_frac=0.25, seed=4)
    tex3 = np.random.default_rng(6).uniform(-20, 300, size=(F, 3))
    r1 = np.stack(list(GoldenMesh(F, p2f, tex1).render_flat(ref_cams)))
    r3 = np.stack(list(GoldenMesh(F, p2f, tex3).render_flat(ref_cams)))

    # --- (5) argmax -----------------------------------------------------------------------------------------
    am1 = find_argmax_nonzero_value(avg1)
    am2 = find_argmax_nonzero_value(avg2, keepdims=True)

    np.savez_compressed(
        OUT / "golden_scene.npz",
        verts=sc["verts"], faces=sc["faces"], c2ws=sc["c2ws"], origin=sc["origin"],
        intrinsics=np.array([sc["f"], sc["cx"], sc["cy"], W, H], dtype=np.float64),
        pix2face=p2f.astype(np.int32),
        **{f"cam_{k}": v for k, v in cam_pins.items()},
    )
    np.savez_compressed(
        OUT / "golden_aggregate.npz",
        idx_imgs=np.stack(idx_imgs), onehot0=onehot0,
        avg1=avg1, counts1=info1["projection_counts"], summed1=info1["summed_projections"],
        soft=np.stack(soft), avg2=avg2, counts2=info2["projection_counts"], summed2=info2["summed_projections"],
        projs2=np.stack(projs2), avg2s=avg2s, counts2s=info2s["projection_counts"],
        summed2s=info2s["summed_projections"],
        vote_imgs=np.stack(vote_imgs), n_vote_classes=np.array(n_vote_classes),
        avg3=avg3.toarray(), counts3=info3["projection_counts"].toarray()[:, 0],
        summed3=info3["summed_projections"].toarray(),
        argmax1=am1, argmax2=am2,
    )
    np.savez_compressed(OUT / "golden_render.npz", tex1=tex1, tex3=tex3, render1=r1, render3=r3)
    for fn in ("golden_scene.npz", "golden_aggregate.npz", "golden_render.npz"):
        print(fn, (OUT / fn).stat().st_size, "bytes")


def main_metashape():
    """MetashapeCameraSet parsing + lens model + distortion maps from the reference's own code."""
    import tempfile

    sys.path.insert(0, str(ROOT / "tests"))
    import metashape_fixture as mf
    import pyproj  # the stub: make Transformer.from_crs(...).transform(...) return three arrays

    from geograypher.cameras.derived_cameras import MetashapeCameraSet

    poses = mf.poses(5)
    pyproj.Transformer.from_crs.return_value.transform.side_effect = lambda xx, yy, zz: (xx * 0, yy * 0, zz * 0)
    with tempfile.TemporaryDirectory() as tmp:
        xml = mf.camera_xml(poses, unaligned=(3,), other_component=(4,), extra_uncalibrated_sensor=True)
        path = Path(tmp, "cameras.xml")
        path.write_text(xml)
        cams = MetashapeCameraSet(camera_file=path, image_folder="/mnt/images", original_image_folder="/data/survey/images")
    cam = cams.cameras[0]
    rng = np.random.default_rng(3)
    xp, yp = rng.uniform(0, cam.image_width, 200), rng.uniform(0, cam.image_height, 200)
    xw, yw = cams.ideal_to_warped(cam, xp, yp)

    # a small strongly distorted sensor, like the reference's test (f = 100, k1 = -0.05), scale 1 and 0.5
    small = cams.cameras[1]
    small.f, small.cx, small.cy = 100.0, 1.25, -0.75
    small.image_width, small.image_height, small.image_size = 161, 129, (129, 161)
    small.distortion_params = dict(k1=-0.05, k2=0.004, k3=0.0, k4=0.0, p1=0.001, p2=-0.0005, b1=0.02, b2=-0.01)
    maps = {}
    for scale in (1.0, 0.5):
        cams.make_distortion_map(small, inversion_downsample=1, image_scale=scale)
        key = cams.distortion_key(small.distortion_params, scale)
        maps[f"i2w_{scale}"] = cams._maps_ideal_to_warped[key]
        maps[f"w2i_{scale}"] = cams._maps_warped_to_ideal[key]
    cams._maps_ideal_to_warped.clear(); cams._maps_warped_to_ideal.clear()
    cams.make_distortion_map(small, inversion_downsample=8, image_scale=1.0)
    maps["w2i_ds8_1.0"] = cams._maps_warped_to_ideal[cams.distortion_key(small.distortion_params, 1.0)]

    np.savez_compressed(
        OUT / "golden_metashape.npz",
        xml=np.array(xml), n_cameras=np.array(len(cams)),
        c2w=np.stack([c.cam_to_world_transform for c in cams.cameras]),
        filenames=np.array([str(c.image_filename) for c in cams.cameras]),
        local_to_epsg_4978=cams.get_local_to_epsg_4978_transform(),
        f=np.array(cam.f), cx=np.array(cam.cx), cy=np.array(cam.cy),
        size=np.array([cam.image_width, cam.image_height]),
        dist_keys=np.array(sorted(cam.distortion_params)), dist_vals=np.array([cam.distortion_params[k] for k in sorted(cam.distortion_params)]),
        xp=xp, yp=yp, xw=xw, yw=yw,
        small_params=np.array([small.f, small.cx, small.cy, small.image_width, small.image_height] +
                              [small.distortion_params[k] for k in ("k1", "k2", "k3", "k4", "p1", "p2", "b1", "b2")]),
        key=np.array(cams.distortion_key(small.distortion_params, 0.5)),
        **{k.replace(".", "_"): v.astype(np.float32) for k, v in maps.items()},
    )
    print("golden_metashape.npz", (OUT / "golden_metashape.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    if "--metashape-only" not in sys.argv:
        main()
    main_metashape()
