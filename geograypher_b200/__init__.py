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


def host_array(shape, dtype="float32", device: int = 0):
    """NumPy array in host memory that the GPU reads in place (see ``_lib.host_array``): hold prediction images in
    these (or in page-locked arrays) and ``aggregate_projected_images`` fetches only the rows it needs over PCIe."""
    from geograypher_b200 import _lib

    return _lib.host_array(shape, dtype, device)

__version__ = "0.1.0"
