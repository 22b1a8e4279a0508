"""oracle/ref_stub.py -- TEST INFRASTRUCTURE (build container only).

Makes the UNMODIFIED reference at /root/reference importable in the build container, where its heavy
third-party imports (pyvista, vtk, geopandas, shapely, pyproj, rasterio, fiona, skimage, ubelt,
matplotlib, imageio, ...) are missing: those module names resolve to inert ``MagicMock`` packages.  The
in-repo NumPy / SciPy code of the hot path (``project_images``, ``aggregate_projected_images``, the
``IndexPredictions`` variant, ``render_flat``'s gather, ``inds_to_one_hot``,
``find_argmax_nonzero_value``, the camera containers) then runs for real; only the rasterizer call
(``pix2face``, which is VTK) has to be supplied by a subclass override -- the reference's own plug-in point
(derived_meshes.py:642).

Used by ``tests/golden/make_golden.py`` to produce the committed golden vectors.  /root/reference does not
exist on the GPU box, so nothing at test / bench run time imports this.
"""
import importlib.abc
import importlib.machinery
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"

_MISSING = [
    "pyvista", "vtk", "pytorch3d", "geopandas", "shapely", "pyproj", "rasterio", "fiona", "skimage",
    "ubelt", "matplotlib", "imageio", "networkx", "rasterstats", "trimesh", "IPython", "contextily",
    "torchvision", "exifread", "piexif", "tifffile",
]


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


def import_reference():
    """Return the reference's ``geograypher`` package, importing it with stubs on first use."""
    if "geograypher" not in sys.modules:
        missing = []
        for name in _MISSING:
            try:
                __import__(name)
            except Exception:
                missing.append(name)
        _MISSING[:] = missing
        sys.meta_path.insert(0, _StubFinder())
        sys.path.insert(0, REFERENCE_ROOT)
    import geograypher

    return geograypher
