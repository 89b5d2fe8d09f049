"""Function-level mirror of ``tobac_flow/sobel.py`` on the CUDA path (fused 27-tap gather + Sobel reducer)."""
import numpy as np

from . import _lib
from .flow import Flow, _sobel_ref, _tag

_sobel_func = _tag(_lib.TF_RED_SOBEL)(_sobel_ref(None))                     # sobel.py:70-86
_sobel_func_uphill = _tag(_lib.TF_RED_SOBEL_UPHILL)(_sobel_ref("uphill"))      # sobel.py:32-48
_sobel_func_downhill = _tag(_lib.TF_RED_SOBEL_DOWNHILL)(_sobel_ref("downhill"))  # sobel.py:51-67


def sobel_reducer(direction=None):
    """sobel.py:118-123: anything other than 'uphill' / 'downhill' selects the plain reducer."""
    if direction == "uphill":
        return _sobel_func_uphill
    if direction == "downhill":
        return _sobel_func_downhill
    return _sobel_func


def sobel(data, forward_flow, backward_flow, method="linear", dtype=np.float32, fill_value=np.nan, direction=None):
    """``tobac_flow.sobel.sobel`` (sobel.py:89-143); note the fp32 default here vs ``Flow.sobel``'s ``None``."""
    return Flow(forward_flow, backward_flow).sobel(data, method=method, dtype=dtype, fill_value=fill_value,
                                                   direction=direction)
