"""Pin oracle/remap_np.py against cv2.remap and the reference's warp_flow tests (tests/test_flow.py:94-161)."""
import numpy as np
import pytest

from oracle import remap_np as rm
from oracle import flow_ops


def _warp(img, flow, method="linear"):
    px, py = rm.warp_positions(flow)
    return rm.remap(img, px, py, method, np.nan)


def test_reference_warp_flow_cases():
    # restated from the reference's tests/test_flow.py:94-161 (3x5 arange image)
    img = np.arange(15, dtype=np.float32).reshape(3, 5)
    z = np.zeros((3, 5, 2), np.float32)
    out = _warp(img, z)
    ok = np.isfinite(out)
    assert np.array_equal(out[ok], img[ok]) and ok[:-1, :-1].all()
    assert np.isnan(out[-1]).all() and np.isnan(out[:, -1]).all()
    f = z.copy(); f[..., 0] = 1
    out = _warp(img, f)
    assert np.array_equal(out[:2, :3], img[:2, 1:4])
    f = z.copy(); f[..., 1] = 1
    out = _warp(img, f)
    assert np.array_equal(out[:1, :4], img[1:2, :4])
    f = z.copy(); f[..., 0] = 0.5
    out = _warp(img, f)
    assert np.array_equal(out[:2, :3], img[:2, :3] + 0.5)


def test_quantisation_and_rounding():
    img = np.arange(40, dtype=np.float32).reshape(4, 10)
    y = np.ones(4, np.float32)
    # 1/32 px quantisation: 0.51 and 0.49 both -> 0.5; 1/64 -> 0 (half-even)
    for dx, want in [(0.51, 0.5), (0.49, 0.5), (1 / 64, 0.0), (3 / 64, 2 / 32)]:
        out = rm.remap(img, np.full(4, 3 + dx, np.float32), y, "linear")
        assert np.allclose(out, img[1, 3] + want)
    # nearest: half-even rounding, no quantisation
    out = rm.remap(img, np.array([2.5, 3.5, -0.5, 9.5], np.float32), y, "nearest", -1.0)
    assert list(out) == [img[1, 2], img[1, 4], img[1, 0], -1.0]


@pytest.mark.skipif(not flow_ops.have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", ["nearest", "linear", "cubic", "lanczos"])
@pytest.mark.parametrize("fill", [np.nan, 0.0, -3.5])
def test_bit_exact_vs_cv2(dtype, method, fill):
    import cv2
    code = dict(nearest=cv2.INTER_NEAREST, linear=cv2.INTER_LINEAR, cubic=cv2.INTER_CUBIC, lanczos=cv2.INTER_LANCZOS4)[method]
    rng = np.random.default_rng(11)
    H, W = 61, 203
    src = (rng.standard_normal((H, W)) * 100).astype(dtype)
    src[5, 7] = np.nan
    src[20:22, 30] = np.nan
    for mag in (0.0, 0.7, 3.0, 40.0):
        flow = (rng.standard_normal((H, W, 2)) * mag).astype(np.float32)
        flow[0:5, :, 0] = np.round(flow[0:5, :, 0] * 64) / 64
        flow[5:9, :, 1] = np.round(flow[5:9, :, 1] * 2) / 2
        px, py = rm.warp_positions(flow, 1, -1)
        ref = cv2.remap(src, np.stack([px, py], -1), None, code, None, cv2.BORDER_CONSTANT, fill)
        mine = rm.remap(src, px, py, method, fill)
        assert np.array_equal(ref, mine, equal_nan=True)


@pytest.mark.skipif(not flow_ops.have_cv2(), reason="cv2 not importable")
def test_int32_nearest_vs_cv2():
    import cv2
    rng = np.random.default_rng(3)
    src = rng.integers(0, 1000, (40, 50)).astype(np.int32)
    flow = (rng.standard_normal((40, 50, 2)) * 4).astype(np.float32)
    px, py = rm.warp_positions(flow)
    ref = cv2.remap(src, np.stack([px, py], -1), None, cv2.INTER_NEAREST, None, cv2.BORDER_CONSTANT, 0)
    assert np.array_equal(ref, rm.remap(src, px, py, "nearest", 0))


@pytest.mark.skipif(not flow_ops.have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("shape", [(3, 5), (8, 8), (1, 9), (9, 64)])
def test_lanczos_small_images_vs_cv2(shape):
    """Images smaller than the 8 x 8 kernel and sampling positions with zero fraction (the centre-tap-only table row)."""
    import cv2
    rng = np.random.default_rng(5)
    H, W = shape
    src = rng.standard_normal((H, W)).astype(np.float32)
    mapx = rng.uniform(-6, W + 5, (40, 50)).astype(np.float32)
    mapy = rng.uniform(-6, H + 5, (40, 50)).astype(np.float32)
    mapx[::7, ::5] = np.round(mapx[::7, ::5])
    mapy[::3, ::11] = np.round(mapy[::3, ::11])
    for fill in (np.nan, 0.0):
        ref = cv2.remap(src, np.stack([mapx, mapy], -1), None, cv2.INTER_LANCZOS4, None, cv2.BORDER_CONSTANT, fill)
        assert np.array_equal(ref, rm.remap(src, mapx, mapy, "lanczos", fill), equal_nan=True)


def test_conversion_free_quantisation_identity():
    """The identity behind `quantise_fast` (csrc/gather.cu): for a float32 position p with |32 p| < 2^22 the integer
    cvRound(32 p) (round half to even) sits in the low mantissa bits of fl32(32 p + 1.5 * 2^23) -- 32 p is exact, so the
    fused multiply-add rounds once, exactly as cvRound does; anything out of range, NaN or infinite yields a value that
    fails the kernel's unsigned range test (q < limit <= 32767 * 32)."""
    rng = np.random.default_rng(11)
    p = np.concatenate([
        rng.uniform(-100, 3000, 200000), rng.uniform(-1, 1, 50000),
        (np.arange(-4000, 4000) + 0.5) / 32.0,                 # exact ties of the 1/32 grid
        (np.arange(-4000, 4000) + 0.5) / 32.0 + 1e-6, np.array([0.0, -0.0, 131071.9, -131071.9, 5e-39, -5e-39]),
    ]).astype(np.float32)

    def quantise_fast(x):
        r = (x.astype(np.float64) * 32.0 + 12582912.0).astype(np.float32)      # one rounding, like the FMA
        return r.view(np.int32).astype(np.int64) - 0x4B400000

    want = np.rint(p.astype(np.float64) * 32.0).astype(np.int64)               # 32 p is exact in float32 and float64
    assert np.array_equal(quantise_fast(p), want)
    limit = 32767 * 32
    with np.errstate(all="ignore"):
        bad = np.array([1.4e5, -1.4e5, 4.2e6, -4.2e6, 1e9, -1e9, 3e38, -3e38, np.inf, -np.inf, np.nan], np.float32)
        q = quantise_fast(bad) & 0xFFFFFFFF                                     # the kernel compares as unsigned
    assert (q >= limit).all()
