"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's dense-flow hot path *above* OpenCV:

* pair normalisation + 8-bit quantisation  — ``tobac_flow/utils/normalisation_utils.py:59-72`` (linear_norm)
  and ``:10-33`` (to_8bit)
* ``calculate_flow`` / ``create_flow``     — ``tobac_flow/flow.py:362-428`` and ``:23-65``
* ``smooth_flow_step``                     — ``tobac_flow/flow.py:530-568``
* the semi-Lagrangian tap gather           — ``tobac_flow/convolve.py:8-86,89-144,147-245,248-348``
* ``Flow.diff`` reducer                    — ``tobac_flow/flow.py:159-191``
* flow-aware Sobel reducers                — ``tobac_flow/sobel.py:7-143``
* nanmean / any reducers used by callers   — ``tobac_flow/detection.py:53-55,187-192,313-320``

Two interchangeable back-ends do the OpenCV part: ``backend="numpy"`` uses the restatements in
``farneback_np`` / ``remap_np`` (always available), ``backend="cv2"`` calls the real OpenCV routines
exactly as the reference does (available wherever the image's opencv-python-headless is importable;
this is what the reference itself executes and is what ``bench.py``'s CPU-baseline legs time).

Pinned by ``tests/test_oracle_ops.py`` against golden vectors produced by the unmodified reference
(``tests/golden/make_golden.py``) and, when /root/reference is present, against the reference live.
"""
import warnings

import numpy as np

from . import farneback_np, remap_np

F32 = np.float32

try:  # pragma: no cover - depends on the image
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None


def have_cv2() -> bool:
    return _cv2 is not None


# ----------------------------------------------------------------------------------------------
# normalisation (normalisation_utils.py:59-72 then :10-33 with vmin=0, vmax=1)
# ----------------------------------------------------------------------------------------------
def pair_to_u8(frame0: np.ndarray, frame1: np.ndarray):
    pair = np.stack([frame0, frame1])
    if pair.dtype != F32:
        pair = pair.astype(np.float64, copy=False)   # numpy keeps float64 and promotes integers: the reference then
    one = pair.dtype.type(1)                         # normalises in float64
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lo = np.nanmin(pair)
        hi = np.nanmax(pair)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        factor = one / (hi - lo) if hi > lo else pair.dtype.type(0)
        scaled = (pair - lo) * factor
        scaled = np.maximum(np.minimum(scaled, 1), 0)  # NaN propagates
        scaled = (scaled - 0) * 255.0  # python float keeps the fp32 array dtype
    ok = np.isfinite(scaled)
    scaled[~ok] = 127
    scaled[0][~ok[0]] = scaled[1][~ok[0]]
    scaled[1][~ok[1]] = scaled[0][~ok[1]]
    q = scaled.astype(np.uint8)
    return q[0], q[1]


# ----------------------------------------------------------------------------------------------
# dense flow for one pair and for a sequence
# ----------------------------------------------------------------------------------------------
def farneback_pair(q0: np.ndarray, q1: np.ndarray, backend: str = "numpy"):
    if backend == "cv2":
        model = _cv2.FarnebackOpticalFlow_create()
        return model.calc(q0, q1, None), model.calc(q1, q0, None)
    return farneback_np.farneback(q0, q1), farneback_np.farneback(q1, q0)


def warp_image(img, flow, method="linear", fill_value=np.nan, dx=0, dy=0, backend="numpy"):
    """One tap plane: img sampled at grid + flow + (dx, dy) (convolve.py:56-84)."""
    px, py = remap_np.warp_positions(flow, dx, dy)
    if backend == "cv2":
        code = {"nearest": _cv2.INTER_NEAREST, "linear": _cv2.INTER_LINEAR,
                "cubic": _cv2.INTER_CUBIC, "lanczos": _cv2.INTER_LANCZOS4}
        if method not in code:
            raise ValueError(f"method must be one of {list(code)}")
        return _cv2.remap(img, np.stack([px, py], -1), None, code[method], None,
                          _cv2.BORDER_CONSTANT, fill_value)
    return remap_np.remap(img, px, py, method, fill_value)


def smooth_flow_step(fwd, bwd, method="linear", backend="numpy"):
    """flow.py:530-568: average each field with the negated, warped opposite field (nanmean)."""
    def blend(a, b):
        w = np.stack([-warp_image(b[..., c], a, method, np.nan, backend=backend) for c in (0, 1)], -1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return np.nanmean([a, w], 0)
    return blend(fwd, bwd), blend(bwd, fwd)


def refine_pair(q0, q1, fwd, bwd, backend="numpy"):
    """flow.py:513-519: vr_model.calc(prev, next, fwd), vr_model.calc(next, prev, bwd)."""
    if backend == "cv2":
        vr = _cv2.VariationalRefinement_create()
        return vr.calc(q0, q1, fwd.copy()), vr.calc(q1, q0, bwd.copy())
    from . import varref_np
    return varref_np.variational_refinement(q0, q1, fwd), varref_np.variational_refinement(q1, q0, bwd)


def calculate_flow(data, smoothing_passes=0, interp_method="linear", backend="numpy", vr_steps=0):
    data = np.asarray(data)
    T = data.shape[0]
    fwd = np.full(data.shape + (2,), np.nan, dtype=F32)
    bwd = np.full(data.shape + (2,), np.nan, dtype=F32)
    for i in range(T - 1):
        q0, q1 = pair_to_u8(data[i], data[i + 1])
        f, b = farneback_pair(q0, q1, backend)
        if vr_steps > 0:
            f, b = refine_pair(q0, q1, f, b, backend)
        for _ in range(smoothing_passes):
            f, b = smooth_flow_step(f, b, interp_method, backend)
        fwd[i], bwd[i + 1] = f, b
    fwd[-1] = -bwd[-1]
    bwd[0] = -fwd[0]
    return fwd, bwd


def create_flow(data, smoothing_passes=0, interp_method="linear", max_value=20, backend="numpy", vr_steps=0):
    fwd, bwd = calculate_flow(data, smoothing_passes, interp_method, backend, vr_steps)
    fwd = np.minimum(np.maximum(fwd, -max_value), max_value)
    bwd = np.minimum(np.maximum(bwd, -max_value), max_value)
    return fwd, bwd


# ----------------------------------------------------------------------------------------------
# semi-Lagrangian tap stack (convolve.py)
# ----------------------------------------------------------------------------------------------
def structure_taps(structure):
    """[(slab, dx, dy)] in output order: slab 0 (t-1), 1 (t), 2 (t+1); row-major (y, x) inside."""
    structure = np.asarray(structure)
    if structure.ndim != 3:
        raise ValueError("structure must have three dimensions")
    if structure.shape[0] != 3:
        raise ValueError("leading dimension of structure must have length 3")
    cy, cx = structure.shape[1] // 2, structure.shape[2] // 2
    taps = []
    for s in range(3):
        ys, xs = np.nonzero(structure[s])
        taps += [(s, int(x) - cx, int(y) - cy) for y, x in zip(ys, xs)]
    return taps


def shift_image(img, dx, dy, fill_value):
    """Integer-offset same-step tap (convolve.py:89-144): out[y, x] = img[y+dy, x+dx], OOB -> fill."""
    h, w = img.shape
    out = np.full((h, w), fill_value, dtype=img.dtype)
    ys0, ys1 = max(0, -dy), min(h, h - dy)
    xs0, xs1 = max(0, -dx), min(w, w - dx)
    if ys1 > ys0 and xs1 > xs0:
        out[ys0:ys1, xs0:xs1] = img[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
    return out


def tap_stack(prev, cur, nxt, fflow, bflow, structure, method, dtype, fill_value, backend="numpy"):
    taps = structure_taps(structure)
    out = np.full((len(taps),) + cur.shape, fill_value, dtype=dtype)
    for n, (slab, dx, dy) in enumerate(taps):
        if slab == 0:
            out[n] = warp_image(prev, bflow, method, fill_value, dx, dy, backend)
        elif slab == 1:
            out[n] = shift_image(cur, dx, dy, fill_value)
        else:
            out[n] = warp_image(nxt, fflow, method, fill_value, dx, dy, backend)
    return out


def convolve(data, fwd, bwd, structure=None, method="linear", dtype=F32, fill_value=np.nan,
             func=None, backend="numpy"):
    if structure is None:
        structure = np.zeros((3, 3, 3), bool)
        structure[1, 1, :] = structure[1, :, 1] = structure[:, 1, 1] = True
    structure = np.asarray(structure)
    assert structure.shape == (3, 3, 3), "Structure input must be a 3x3x3 array"
    data = np.asarray(data)
    T = data.shape[0]
    n = int(np.count_nonzero(structure))
    res = np.full(data.shape if func is not None else (n,) + data.shape, fill_value, dtype=dtype)
    for i in range(T):
        blank = np.full(data[i].shape, fill_value, dtype=dtype)
        prev = data[i - 1] if i > 0 else blank
        nxt = data[i + 1] if i < T - 1 else blank
        stack = tap_stack(prev, data[i], nxt, fwd[i], bwd[i], structure, method, dtype, fill_value, backend)
        if func is not None:
            res[i] = func(stack)
        else:
            res[:, i] = stack
    if func is not None:
        res[np.isnan(data)] = fill_value
    return res


# ----------------------------------------------------------------------------------------------
# reducers
# ----------------------------------------------------------------------------------------------
def diff_reducer(x):
    """flow.py:182-186"""
    with np.errstate(invalid="ignore"):
        return np.nansum([x[2] - x[1], x[1] - x[0]], axis=0) * 1 / np.maximum(
            np.sum([np.isfinite(x[2]), np.isfinite(x[0])], 0), 1)


def diff(data, fwd, bwd, method="linear", dtype=F32, backend="numpy"):
    s = np.zeros((3, 3, 3))
    s[:, 1, 1] = 1
    return convolve(data, fwd, bwd, s, method, dtype, np.nan, diff_reducer, backend)


def nanmean_reducer(x):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.nanmean(x, 0)


def any_reducer(x):
    return np.any(x, axis=0)


def sobel_weights():
    """sobel.py:7-26 — S[t,y,x] = w[t]*w[y]*d[x]; the three gradients are S, S.T(1,2,0), S.T(2,0,1)."""
    w = np.array([1, 2, 1])
    d = np.array([-1, 0, 1])
    S = w[:, None, None] * w[None, :, None] * d[None, None, :]
    return S.ravel(), S.transpose(1, 2, 0).ravel(), S.transpose(2, 0, 1).ravel()


def sobel_reducer(direction=None):
    k0, k1, k2 = (k[:, None, None] for k in sobel_weights())

    def reduce(x):
        with np.errstate(invalid="ignore"):
            if direction == "uphill":
                x = np.fmax(x - x[13], 0)
            elif direction == "downhill":
                x = np.fmin(x - x[13], 0)
            else:
                x = x - x[13]
            acc = np.nansum(x * k0, 0) ** 2
            acc += np.nansum(x * k1, 0) ** 2
            acc += np.nansum(x * k2, 0) ** 2
            return acc ** 0.5
    return reduce


def sobel(data, fwd, bwd, method="linear", dtype=F32, fill_value=np.nan, direction=None,
          backend="numpy"):
    return convolve(data, fwd, bwd, np.ones((3, 3, 3), bool), method, dtype, fill_value,
                    sobel_reducer(direction), backend)
