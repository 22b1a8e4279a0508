"""geograypher_b200: B200-native multiview projection (pix2face / aggregate / render_flat) behind
open-forest-observatory/geograypher's TexturedPhotogrammetryMesh and PhotogrammetryCamera(Set) API."""
from geograypher_b200.cameras import (
    MetashapeCameraSet,
    PhotogrammetryCamera,
    PhotogrammetryCameraSet,
    SegmentorPhotogrammetryCameraSet,
)
from geograypher_b200.meshes import (
    TexturedPhotogrammetryMesh,
    TexturedPhotogrammetryMeshChunked,
    TexturedPhotogrammetryMeshIndexPredictions,
)
from geograypher_b200.predictors import ArraySegmentor, LookUpSegmentor, Segmentor
from geograypher_b200.utils.indexing import find_argmax_nonzero_value

__version__ = "0.1.0"
