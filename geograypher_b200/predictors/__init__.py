from geograypher_b200.predictors.segmentor import Segmentor
from geograypher_b200.predictors.derived_segmentors import ArraySegmentor, LookUpSegmentor

__all__ = ["Segmentor", "ArraySegmentor", "LookUpSegmentor"]
