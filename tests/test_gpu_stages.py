"""GPU stage parity: the Farneback pipeline stage by stage (through the C ABI) against the oracle's restatement of
OpenCV's optflowgf.cpp (oracle/farneback_np.py, itself pinned to cv2 in tests/test_oracle_farneback.py).

Level images: the one-pass kernels (row kernel with combined blur x resize taps for W % 4 == 0, tile kernels otherwise)
agree with the separable two-pass kernels and with the oracle to a few ulp of the 0..255 range (fp32 sums; the tile
kernels add in the two-pass order and give the same bits, the row kernel associates differently).  Polynomial expansion:
fp32 horizontal sums where OpenCV uses fp64 accumulators, so a relative tolerance (2e-5 of the channel's range).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import farneback_np  # noqa: E402
from tobac_flow_b200 import _lib, flow as tflow  # noqa: E402


def _u8_pair(H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 120 + 80 * np.sin(xx / 17.0 + seed) * np.cos(yy / 11.0) + 20 * rng.standard_normal((H, W))
    q0 = np.clip(base, 0, 255).astype(np.uint8)
    q1 = np.clip(np.roll(base, (1, 2), (0, 1)) + 5 * rng.standard_normal((H, W)), 0, 255).astype(np.uint8)
    return q0, q1


# sizes: CONUS-like even chain, odd sizes (scale != 2^k exactly, W % 4 != 0), tiny levels
@pytest.mark.parametrize("H,W", [(375, 625), (300, 500), (339, 170), (129, 257), (64, 1100), (1500, 2500)])
def test_pyramid_levels_fused_vs_two_pass_vs_oracle(H, W):
    q0, q1 = _u8_pair(H, W, H + W)
    plan = farneback_np.level_plan(H, W)
    mine_plan = _lib.level_plan(H, W)
    assert [(l["h"], l["w"]) for l in plan] == mine_plan
    for li, lvl in enumerate(plan):
        fused = tflow.fb_pyramid_level(q0, q1, li)
        two = tflow.fb_pyramid_level(q0, q1, li, two_pass=True)
        assert fused.shape == (2, lvl["h"], lvl["w"])
        if W % 4 == 0:
            assert np.abs(fused - two).max() <= 2e-4, (li, np.abs(fused - two).max())
        else:
            assert np.array_equal(fused, two), (li, np.abs(fused - two).max())
        if H * W <= 400 * 700:      # the numpy oracle is slow on big frames
            for k, q in enumerate((q0, q1)):
                ref = farneback_np.pyramid_level(q, lvl)
                assert np.abs(fused[k] - ref).max() <= 2e-4, (li, k, np.abs(fused[k] - ref).max())


@pytest.mark.parametrize("h,w", [(47, 78), (94, 156), (375, 625), (50, 50), (33, 130), (200, 119)])
def test_polyexp_vs_oracle(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    yy, xx = np.mgrid[0:h, 0:w]
    img = (128 + 90 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 10 * rng.standard_normal((h, w))).astype(np.float32)
    img2 = (rng.uniform(0, 255, (h, w))).astype(np.float32)
    out = tflow.fb_polyexp(np.stack([img, img2]))
    for k, im in enumerate((img, img2)):
        ref = farneback_np.poly_exp(im)
        scale = np.abs(ref).max(axis=(0, 1))
        err = np.abs(out[k] - ref).max(axis=(0, 1)) / scale
        assert (err <= 2e-5).all(), err
