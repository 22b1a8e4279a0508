"""Read-ahead of per-view predictions (geograypher_b200/utils/prefetch.py) and when the mesh uses it."""
import threading
import time

import numpy as np
import pytest

import geograypher_b200 as gg
from geograypher_b200.utils.prefetch import OrderedPrefetcher


def test_results_in_order_and_concurrent():
    active, peak, lock = [0], [0], threading.Lock()

    def slow(k):
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        time.sleep(0.02)
        with lock:
            active[0] -= 1
        return k * k

    t0 = time.perf_counter()
    with OrderedPrefetcher(slow, 40, threads=8) as pf:
        got = [pf.fetch(k) for k in range(40)]
    dt = time.perf_counter() - t0
    assert got == [k * k for k in range(40)]
    assert peak[0] > 1 and dt < 0.5 * 40 * 0.02  # 0.8 s serial


def test_exception_surfaces_at_its_item_and_close_stops_the_rest():
    calls = []

    def fn(k):
        calls.append(k)
        if k == 3:
            raise ValueError("bad view 3")
        time.sleep(0.005)
        return k

    pf = OrderedPrefetcher(fn, 1000, threads=2, depth=4)
    assert [pf.fetch(k) for k in range(3)] == [0, 1, 2]
    with pytest.raises(ValueError, match="bad view 3"):
        pf.fetch(3)
    pf.close()
    assert max(calls) < 20  # nothing far ahead was started, nothing runs after close()
    n = len(calls)
    time.sleep(0.03)
    assert len(calls) == n


def test_depth_is_bounded_by_bytes():
    started = []

    def fn(k):
        started.append(k)
        return np.zeros(1 << 20, dtype=np.uint8)

    with OrderedPrefetcher(fn, 100, threads=4, depth=32, max_bytes=3 << 20, size_of=lambda a: a.nbytes) as pf:
        pf.fetch(0)
        time.sleep(0.05)
        assert pf.depth == 3 and max(started) <= 4 + 3  # at most the first pool's worth, then 3 items ahead
        pf.fetch(1)
        time.sleep(0.05)
        assert max(started) <= 4 + 3


def test_skipping_items_is_allowed():
    with OrderedPrefetcher(lambda k: k, 50, threads=3) as pf:
        assert pf.fetch(0) == 0 and pf.fetch(10) == 10 and pf.fetch(11) == 11 and pf.fetch(49) == 49
        with pytest.raises(IndexError):
            pf.fetch(50)


def _mesh(**kw):
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=float)
    return gg.TexturedPhotogrammetryMesh((verts, np.array([[0, 1, 2]])), log_level="WARNING", **kw)


def _cams(n=4):
    T = np.eye(4)
    return gg.PhotogrammetryCameraSet(cam_to_world_transforms=[T] * n, intrinsic_params_per_sensor_type={
        0: dict(f=100.0, cx=0.0, cy=0.0, image_width=8, image_height=6, distortion_params={})})


def test_mesh_reads_ahead_only_for_decoded_predictions(tmp_path):
    cams = _cams()
    in_memory = gg.SegmentorPhotogrammetryCameraSet(cams, gg.ArraySegmentor([np.zeros((6, 8, 2), np.float32)] * 4, num_classes=2))
    look_up = gg.SegmentorPhotogrammetryCameraSet(cams, gg.LookUpSegmentor(tmp_path, tmp_path, num_classes=2))

    class Network(gg.Segmentor):  # says nothing about thread safety
        def segment_image(self, image, **kwargs):
            return np.zeros((6, 8, 2), np.float32)

    custom = gg.SegmentorPhotogrammetryCameraSet(cams, Network(num_classes=2))
    mesh = _mesh()
    kinds = {}
    for name, cs in [("in_memory", in_memory), ("look_up", look_up), ("custom", custom), ("files", cams)]:
        f = mesh._prediction_fetcher(cs, 4, 1.0, None, None)
        kinds[name] = isinstance(f, OrderedPrefetcher)
        getattr(f, "close", lambda: None)()
    assert kinds == {"in_memory": False, "look_up": True, "custom": False, "files": True}
    # explicit settings win; a caller-supplied getter and single views are never read ahead
    assert isinstance(_mesh(prefetch_threads=0)._prediction_fetcher(look_up, 4, 1.0, None, None), OrderedPrefetcher) is False
    forced = _mesh(prefetch_threads=3)._prediction_fetcher(custom, 4, 1.0, None, None)
    assert isinstance(forced, OrderedPrefetcher) and forced.threads == 3
    forced.close()
    assert not isinstance(mesh._prediction_fetcher(look_up, 4, 1.0, lambda k: None, None), OrderedPrefetcher)
    assert not isinstance(mesh._prediction_fetcher(look_up, 1, 1.0, None, None), OrderedPrefetcher)


def test_fetcher_returns_what_the_inline_path_returns(tmp_path):
    """LookUpSegmentor over .npy index images: the read-ahead path hands the aggregation the same (array, kind, C)."""
    from geograypher_b200 import _lib

    T = np.eye(4)
    names = [tmp_path / "imgs" / f"{i:03d}.JPG" for i in range(6)]
    cams = gg.PhotogrammetryCameraSet(cam_to_world_transforms=[T] * 6, image_filenames=names, image_folder=tmp_path / "imgs",
                                      intrinsic_params_per_sensor_type={0: dict(f=100.0, cx=0.0, cy=0.0, image_width=8,
                                                                                image_height=6, distortion_params={})})
    (tmp_path / "preds").mkdir()
    rng = np.random.default_rng(0)
    images = [rng.integers(0, 3, size=(6, 8), dtype=np.uint8) for _ in range(6)]
    for n, img in zip(names, images):
        np.save(tmp_path / "preds" / (n.stem + ".npy"), img)
    seg = gg.SegmentorPhotogrammetryCameraSet(cams, gg.LookUpSegmentor(tmp_path / "imgs", tmp_path / "preds", num_classes=3))
    mesh = _mesh()
    getter = seg.get_class_index_image_by_index
    fetch = mesh._prediction_fetcher(seg, 6, 1.0, None, getter)
    assert isinstance(fetch, OrderedPrefetcher)
    try:
        for k in range(6):
            arr, kind, C = fetch(k)
            ref_arr, ref_kind, ref_C = mesh._fetch_prediction(seg, k, 1.0, None, getter)
            assert kind == ref_kind == _lib.PRED_INDEX_U8 and C == ref_C == 3
            np.testing.assert_array_equal(arr, images[k])
            np.testing.assert_array_equal(arr, ref_arr)
    finally:
        fetch.close()
