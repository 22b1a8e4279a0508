"""Metashape camera-file parsing (reference: geograypher/utils/parsing.py:46-157), xml.etree + NumPy only."""
import xml.etree.ElementTree as ET

import numpy as np

from geograypher_b200.constants import PATH_TYPE


def make_4x4_transform(rotation_str: str, translation_str: str, scale_str: str = "1") -> np.ndarray:
    """``[s*R | t; 0 0 0 1]`` from Metashape's strings: 9 row-major rotation entries, 3 translation entries and a
    scale (reference parsing.py:46-70).  Raises ValueError for an improper rotation."""
    rotation = np.array(rotation_str.split(), dtype=float).reshape(3, 3)
    if not np.isclose(np.linalg.det(rotation), 1.0, atol=1e-8, rtol=0):
        raise ValueError(f"Inproper rotation matrix with determinant {np.linalg.det(rotation)}")
    transform = np.eye(4)
    transform[:3, :3] = rotation * float(scale_str)
    transform[:3, 3] = np.array(translation_str.split(), dtype=float)
    return transform


def parse_transform_metashape(camera_file: PATH_TYPE, return_component_id: bool = False):
    """Local (chunk) -> EPSG:4978 transform of the ACTIVE component, or None when the chunk is not georeferenced
    (reference parsing.py:73-111)."""
    components = ET.parse(camera_file).getroot().find("chunk").find("components")
    active_id = components.get("active_id")
    transform = components.find(f"component[@id='{active_id}']").find("transform")
    if transform is None:
        local_to_epsg_4978 = None
    else:
        local_to_epsg_4978 = make_4x4_transform(
            transform.find("rotation").text, transform.find("translation").text, transform.find("scale").text
        )
    return (local_to_epsg_4978, active_id) if return_component_id else local_to_epsg_4978


def parse_sensors(sensors, default_sensor_dict=None):
    """{sensor id: {image_width, image_height, f, cx, cy, distortion_params}} from the <sensors> element; a sensor
    without an adjusted calibration gets the defaults (or None) -- reference parsing.py:114-157."""
    out = {}
    for sensor in sensors:
        d = {"image_width": int(sensor[0].get("width")), "image_height": int(sensor[0].get("height"))}
        calibration = sensor.find("calibration[@class='adjusted']")
        if calibration is None:
            if default_sensor_dict is not None:
                d.update(default_sensor_dict)
            else:
                d = None
        else:
            d["f"] = float(calibration.find("f").text)
            cx, cy = calibration.find("cx"), calibration.find("cy")
            try:
                d["cx"] = float(cx.text) if cx is not None else default_sensor_dict["cx"]
                d["cy"] = float(cy.text) if cy is not None else default_sensor_dict["cy"]
                d["distortion_params"] = {
                    el.tag: float(el.text) for el in calibration if el.tag not in ("resolution", "f", "cx", "cy")
                }
            except (KeyError, TypeError):
                d = None
        out[int(sensor.get("id"))] = d
    return out
