"""ORACLE (test infrastructure, never shipped, never on the product path).

numpy restatement of ``cv2.remap(src, map32FC2, None, interp, dst, BORDER_CONSTANT, fill)`` as the
reference calls it at ``tobac_flow/convolve.py:65-84`` and ``tobac_flow/utils/flow_utils.py:90-98``.

The arithmetic lives in OpenCV (not vendored in /root/reference; image has 4.13.0).  Restated from
opencv/modules/imgproc/src/imgwarp.cpp (remap, remapNearest, remapBilinear, remapBicubic,
interpolateLinear/Cubic/Lanczos4, remapLanczos4, initInterTab2D):

* float maps are converted to fixed point: s = cvRound(coord * 32) (round half even), integer part
  s >> 5, fraction (s & 31) / 32;  nearest uses cvRound(coord) with no sub-pixel part;
* weights are fp32 table entries wy[k1] * wx[k2]; accumulation in fp32 for fp32 sources and fp64 for
  fp64 sources, left to right;
* BORDER_CONSTANT is evaluated per tap, so an out-of-image tap contributes fill*w even when w == 0
  (with fill = NaN the last row/column of a zero-flow linear warp is NaN).

Pinned against cv2 in ``tests/test_oracle_remap.py`` and against the five ``warp_flow`` cases of the
reference's ``tests/test_flow.py:94-161``.
"""
import numpy as np

F32 = np.float32

INTERP_CODES = {"nearest": 0, "linear": 1, "cubic": 2, "lanczos": 3}


def _round_half_even_i64(v: np.ndarray) -> np.ndarray:
    with np.errstate(invalid="ignore"):
        r = np.rint(v)
    bad = ~np.isfinite(r) | (np.abs(r) > 2.0e9)
    r = np.where(bad, -2147483648.0, r)  # cvRound of NaN/inf/overflow -> INT_MIN ("integer indefinite")
    return r.astype(np.int64)


def _sat_short(v: np.ndarray) -> np.ndarray:
    return np.clip(v, -32768, 32767)


def _cubic_coeffs(frac_idx: np.ndarray) -> np.ndarray:
    """interpolateCubic on x = idx/32, fp32 arithmetic, A = -0.75."""
    A = F32(-0.75)
    x = (frac_idx.astype(F32) * F32(1.0 / 32.0)).astype(F32)
    one = F32(1)
    c0 = ((A * (x + one) - F32(5) * A) * (x + one) + F32(8) * A) * (x + one) - F32(4) * A
    c1 = ((A + F32(2)) * x - (A + F32(3))) * x * x + one
    omx = one - x
    c2 = ((A + F32(2)) * omx - (A + F32(3))) * omx * omx + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(F32)


def remap_linear_replicate(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    """cv2.remap(src32f, map, INTER_LINEAR, BORDER_REPLICATE) (used by cv2.VariationalRefinement): same 1/32-px
    quantisation and weights as below, taps clamped to the image."""
    src = np.asarray(src, dtype=F32)
    H, W = src.shape
    sx = _round_half_even_i64(np.asarray(mapx, F32) * F32(32))
    sy = _round_half_even_i64(np.asarray(mapy, F32) * F32(32))
    ix, iy = _sat_short(sx >> 5), _sat_short(sy >> 5)
    fx = ((sx & 31).astype(F32) * F32(1.0 / 32.0)).astype(F32)
    fy = ((sy & 31).astype(F32) * F32(1.0 / 32.0)).astype(F32)
    x0, x1 = np.clip(ix, 0, W - 1), np.clip(ix + 1, 0, W - 1)
    y0, y1 = np.clip(iy, 0, H - 1), np.clip(iy + 1, 0, H - 1)
    one = F32(1)
    w00, w01, w10, w11 = (one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx
    return (((src[y0, x0] * w00 + src[y0, x1] * w01) + src[y1, x0] * w10) + src[y1, x1] * w11).astype(F32)


def remap(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray, method: str = "linear",
          fill_value=np.nan) -> np.ndarray:
    """Gather ``src`` (H, W) at float32 positions (mapx, mapy) (any common shape)."""
    if method not in INTERP_CODES:
        raise ValueError(f"method must be one of {list(INTERP_CODES)}")
    src = np.asarray(src)
    H, W = src.shape
    mapx = np.asarray(mapx, dtype=F32)
    mapy = np.asarray(mapy, dtype=F32)
    if src.dtype == np.float64:
        wt = np.float64
    elif src.dtype == np.float32:
        wt = F32
    else:
        if method != "nearest":
            raise TypeError("cv2.remap: only nearest interpolation supports integer sources")
        wt = src.dtype
    with np.errstate(invalid="ignore", over="ignore"):
        fill = np.array(fill_value).astype(src.dtype)

    if method == "nearest":
        ix = _sat_short(_round_half_even_i64(mapx))
        iy = _sat_short(_round_half_even_i64(mapy))
        inb = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
        out = src[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)]
        return np.where(inb, out, fill).astype(src.dtype)

    sx = _round_half_even_i64(mapx * F32(32))
    sy = _round_half_even_i64(mapy * F32(32))
    ix = _sat_short(sx >> 5)
    iy = _sat_short(sy >> 5)
    fxi = sx & 31
    fyi = sy & 31
    cv = fill.astype(wt)

    if method == "linear":
        fx = (fxi.astype(F32) * F32(1.0 / 32.0)).astype(F32)
        fy = (fyi.astype(F32) * F32(1.0 / 32.0)).astype(F32)
        wx = [F32(1) - fx, fx]
        wy = [F32(1) - fy, fy]
        acc = None
        for k1 in range(2):
            for k2 in range(2):
                w = (wy[k1] * wx[k2]).astype(F32)
                yy = iy + k1
                xx = ix + k2
                inb = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                v = np.where(inb, src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], fill).astype(wt)
                with np.errstate(invalid="ignore"):
                    term = v * w.astype(wt)
                    acc = term if acc is None else acc + term
        outside = (ix >= W) | (ix + 1 < 0) | (iy >= H) | (iy + 1 < 0)
        return np.where(outside, cv, acc).astype(src.dtype)

    if method == "cubic":
        cx = _cubic_coeffs(fxi)
        cy = _cubic_coeffs(fyi)
        x0 = ix - 1
        y0 = iy - 1
        fast = (x0 >= 0) & (x0 < max(W - 3, 0)) & (y0 >= 0) & (y0 < max(H - 3, 0))
        outside = (x0 >= W) | (x0 + 4 <= 0) | (y0 >= H) | (y0 + 4 <= 0)
        # fast path: sum over rows of (4-term row expression), left to right
        acc_fast = None
        # border path: cv + sum (S - cv) * w over in-bounds taps
        with np.errstate(invalid="ignore"):
            acc_b = cv * np.ones(mapx.shape, dtype=wt)
            for k1 in range(4):
                yy = y0 + k1
                yin = (yy >= 0) & (yy < H)
                row = None
                for k2 in range(4):
                    xx = x0 + k2
                    xin = (xx >= 0) & (xx < W)
                    w = (cy[..., k1] * cx[..., k2]).astype(F32).astype(wt)
                    s = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(wt)
                    t = s * w
                    row = t if row is None else row + t
                    acc_b = np.where(yin & xin, acc_b + (s - cv) * w, acc_b)
                acc_fast = row if acc_fast is None else acc_fast + row
        out = np.where(fast, acc_fast, acc_b)
        return np.where(outside & ~fast, cv, out).astype(src.dtype)

    # lanczos (cv2.INTER_LANCZOS4, remapLanczos4): 8 x 8 taps from (ix - 3, iy - 3), table weights wy[k1] * wx[k2] in fp32;
    # inside: sum += (eight-term row expression, left to right) row by row; near the border: cv + sum (S - cv) * w over
    # the in-bounds taps in row-major order; entirely outside: cv
    tab = lanczos4_table()
    cx = tab[fxi]
    cy = tab[fyi]
    x0 = ix - 3
    y0 = iy - 3
    fast = (x0 >= 0) & (x0 < max(W - 7, 0)) & (y0 >= 0) & (y0 < max(H - 7, 0))
    outside = (x0 >= W) | (x0 + 8 <= 0) | (y0 >= H) | (y0 + 8 <= 0)
    with np.errstate(invalid="ignore", over="ignore"):
        acc_fast = np.zeros(mapx.shape, dtype=wt)
        acc_b = cv * np.ones(mapx.shape, dtype=wt)
        for k1 in range(8):
            yy = y0 + k1
            yin = (yy >= 0) & (yy < H)
            row = None
            for k2 in range(8):
                xx = x0 + k2
                xin = (xx >= 0) & (xx < W)
                w = (cy[..., k1] * cx[..., k2]).astype(F32).astype(wt)
                sv = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(wt)
                t = sv * w
                row = t if row is None else row + t
                acc_b = np.where(yin & xin, acc_b + (sv - cv) * w, acc_b)
            acc_fast = acc_fast + row
    out = np.where(fast, acc_fast, acc_b)
    return np.where(outside & ~fast, cv, out).astype(src.dtype)


_LANCZOS_TAB = None


def lanczos4_table() -> np.ndarray:
    """interpolateLanczos4 for the 32 table fractions (imgwarp.cpp): eight fp32 coefficients per fraction, sines / cosines
    in double, normalised in fp32."""
    global _LANCZOS_TAB
    if _LANCZOS_TAB is not None:
        return _LANCZOS_TAB
    import math
    s45 = 0.70710678118654752440084436210485
    cs = [(1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45)]
    tab = np.zeros((32, 8), F32)
    tab[0, 3] = 1            # x < FLT_EPSILON: the centre tap alone
    for i in range(1, 32):
        x = F32(F32(i) * F32(1.0 / 32.0))
        y0 = -(float(x) + 3) * math.pi * 0.25
        s0, c0 = math.sin(y0), math.cos(y0)
        co = np.zeros(8, F32)
        ssum = F32(0)
        for k in range(8):
            y = -float(F32(F32(x + F32(3)) - F32(k))) * math.pi * 0.25
            co[k] = F32((cs[k][0] * s0 + cs[k][1] * c0) / (y * y))
            ssum = F32(ssum + co[k])
        tab[i] = (co * F32(F32(1) / ssum)).astype(F32)
    _LANCZOS_TAB = tab
    return tab


def warp_positions(flow: np.ndarray, dx: int = 0, dy: int = 0):
    """Sampling position p = fl32(fl32(flow + offset) + grid)   (convolve.py:56-63)."""
    h, w = flow.shape[:2]
    px = (flow[..., 0].astype(F32) + F32(dx)).astype(F32) + np.arange(w, dtype=F32)[None, :]
    py = (flow[..., 1].astype(F32) + F32(dy)).astype(F32) + np.arange(h, dtype=F32)[:, None]
    return px.astype(F32), py.astype(F32)
