"""A small Metashape camera-file generator for tests (same schema as Agisoft's XML export; the calibration values are
those of the DJI M3M sensor quoted in the reference's tests/test_derived_cameras.py:19-65)."""
import numpy as np

M3M = dict(width=5280, height=3956, f=3705.4728792737214, cx=11.6738523909, cy=-27.7497969199, b1=0.5262024073,
           b2=-0.3058334293, k1=-0.0919367147, k2=-0.0762807468, k3=0.1162639394, k4=-0.0761413904, p1=-0.0003134847,
           p2=0.0001164035)

CHUNK_ROTATION = ("0.8710699454 0.3070967021 -0.3833128821 -0.4909742007 0.5658364395 -0.6623997719 "
                  "0.0134716110 0.7651932691 0.6436596744")
CHUNK_TRANSLATION = "-2499366.429888 -4256836.520906 4029217.965003"
CHUNK_SCALE = "12.354192447331316"


def camera_xml(poses, labels=None, sensor=M3M, georeferenced=True, extra_uncalibrated_sensor=False,
               unaligned=(), other_component=()):
    """poses: list of 4x4 cam-to-chunk transforms."""
    labels = labels or [f"/data/survey/images/flight1/img_{i:05d}.JPG" for i in range(len(poses))]
    cal = "".join(f"<{k}>{sensor[k]!r}</{k}>" for k in ("f", "cx", "cy", "b1", "b2", "k1", "k2", "k3", "k4", "p1", "p2")
                  if k in sensor)
    sensors = (f'<sensor id="0" label="M3M" type="frame"><resolution width="{sensor["width"]}" height="{sensor["height"]}"/>'
               f'<calibration type="frame" class="adjusted"><resolution width="{sensor["width"]}" height="{sensor["height"]}"/>'
               f"{cal}</calibration></sensor>")
    if extra_uncalibrated_sensor:
        sensors += '<sensor id="1" label="raw" type="frame"><resolution width="640" height="480"/></sensor>'
    transform = ""
    if georeferenced:
        transform = (f'<transform><rotation locked="true">{CHUNK_ROTATION}</rotation><translation locked="true">'
                     f'{CHUNK_TRANSLATION}</translation><scale locked="true">{CHUNK_SCALE}</scale></transform>')
    cams = ""
    for i, (T, label) in enumerate(zip(poses, labels)):
        comp = "1" if i in other_component else "0"
        tr = "" if i in unaligned else "<transform>" + " ".join(repr(float(v)) for v in np.asarray(T).ravel()) + "</transform>"
        cams += f'<camera id="{i}" sensor_id="0" component_id="{comp}" label="{label}">{tr}</camera>'
    return ('<?xml version="1.0" encoding="UTF-8"?><document version="2.0.0"><chunk label="Chunk 1" enabled="true">'
            f'<sensors next_id="2">{sensors}</sensors>'
            f'<components next_id="2" active_id="0"><component id="0" label="Component 1">{transform}</component>'
            '<component id="1" label="Component 2"/></components>'
            f'<cameras next_id="{len(poses)}" next_group_id="1"><group id="0" label="Group 1" type="folder">{cams}</group>'
            "</cameras></chunk></document>")


def poses(n=4, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        T = np.eye(4)
        a = rng.normal(0, 0.02, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        T[:3, :3] = Rz @ np.diag([1.0, -1.0, -1.0]) @ Rx
        T[:3, 3] = [7.0 + 0.5 * i, 7.0 - 0.3 * i, 0.43]
        out.append(T)
    return out
