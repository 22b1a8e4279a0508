from geograypher_b200.meshes.meshes import LocalMesh, TexturedPhotogrammetryMesh
from geograypher_b200.meshes.derived_meshes import (
    TexturedPhotogrammetryMeshChunked,
    TexturedPhotogrammetryMeshIndexPredictions,
)

__all__ = [
    "LocalMesh",
    "TexturedPhotogrammetryMesh",
    "TexturedPhotogrammetryMeshChunked",
    "TexturedPhotogrammetryMeshIndexPredictions",
]
