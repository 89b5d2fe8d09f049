"""Import the UNMODIFIED reference (``/root/reference/tobac_flow``) in the build container.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so nothing that runs
there may import this module; it is used by ``make_golden.py`` (to generate the committed golden
vectors) and by the optional ``-m "not gpu"`` cross-checks that skip when the reference is absent.

The reference needs ``xarray``, ``pyproj``, ``skimage`` and its own compiled Cython watershed at import
time; none are installed here and none are on the dense-flow hot path, so they are stubbed
(SURVEY.md Appendix B).  ``cv2.optflow`` (opencv-contrib) is absent too:
``cv2.optflow.createOptFlow_Farneback`` is aliased to ``cv2.FarnebackOpticalFlow_create`` whose
defaults (5, 0.5, False, 13, 10, 5, 1.1, 0) are the ones the contrib factory uses.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "tobac_flow")):
        return False
    try:
        import cv2  # noqa: F401
    except Exception:
        return False
    return True


class DataArray:
    """Tiny ndarray wrapper standing in for ``xarray.DataArray`` (only what detection.py touches)."""

    def __init__(self, data, coords=None, dims=None, t=None):
        import numpy as np

        self.data = np.asarray(data)
        self.coords = coords if coords is not None else {}
        self.dims = dims if dims is not None else ("t", "y", "x")
        if t is not None:
            self.t = t
        elif isinstance(self.coords, dict) and "t" in self.coords:
            self.t = self.coords["t"]

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    def to_numpy(self):
        return self.data

    def compute(self):
        return self

    def __array__(self, dtype=None, copy=None):
        import numpy as np

        return np.asarray(self.data, dtype=dtype)

    def __getitem__(self, item):
        return self.data[item]

    def __ge__(self, other):
        return self.data >= other


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load_reference():
    """Return the reference's ``tobac_flow`` package, imported unmodified under stubs."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not available in this environment")
    import cv2

    if "xarray" not in sys.modules:
        _stub("xarray", DataArray=DataArray, Dataset=DataArray)
    if "pyproj" not in sys.modules:
        _stub("pyproj", Geod=lambda **k: None, Proj=None)
    if "skimage" not in sys.modules:
        _stub("skimage")
        _stub("skimage.segmentation", watershed=None)
        _stub("skimage.feature", peak_local_max=None)
        _stub("skimage.morphology")
        _stub(
            "skimage.morphology._util",
            _validate_connectivity=None,
            _offsets_to_raveled_neighbors=None,
        )
        _stub("skimage.util", crop=None)
        _stub("skimage.segmentation._watershed", _validate_inputs=None)
    _stub("tobac_flow._watershed", watershed_raveled=None)
    if not hasattr(cv2, "optflow"):
        cv2.optflow = types.SimpleNamespace(
            createOptFlow_Farneback=cv2.FarnebackOpticalFlow_create
        )
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tobac_flow  # noqa: F401
    from tobac_flow import flow as _flow  # noqa: F401

    _loaded = sys.modules["tobac_flow"]
    return _loaded
