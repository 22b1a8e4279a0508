from geograypher_b200.cameras.cameras import PhotogrammetryCamera, PhotogrammetryCameraSet
from geograypher_b200.cameras.derived_cameras import MetashapeCameraSet
from geograypher_b200.cameras.segmentor import SegmentorPhotogrammetryCameraSet

__all__ = ["PhotogrammetryCamera", "PhotogrammetryCameraSet", "MetashapeCameraSet", "SegmentorPhotogrammetryCameraSet"]
