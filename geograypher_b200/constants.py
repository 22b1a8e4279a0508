"""Constants of the reference that the hot path refers to (geograypher/constants.py), restated."""
import typing
from pathlib import Path

PATH_TYPE = typing.Union[str, Path]
# constants.py:27 -- value written for pixels that cannot be represented when renders are cast to uint8
NULL_TEXTURE_INT_VALUE = 0
# constants.py:106-113
EXAMPLE_INTRINSICS = {
    "f": 1000,
    "cx": 0,
    "cy": 0,
    "image_width": 800,
    "image_height": 600,
    "distortion_params": {},
}
# constants.py:130 -- only used by the chunked variant's signature
CHUNKED_MESH_BUFFER_DIST_METERS = 125
# The reference uses pyproj.CRS.from_epsg(4978); pyproj is optional here, so the CRS is identified by its code.
EARTH_CENTERED_EARTH_FIXED_CRS = "EPSG:4978"
LAT_LON_CRS = "EPSG:4326"
CACHE_FOLDER = Path(Path.home(), ".cache", "geograypher_b200")
