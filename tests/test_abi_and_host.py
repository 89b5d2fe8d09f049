"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header declares; host logic."""
import ctypes
import os
import re
from functools import partial

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tobac_flow_b200 import build, _lib
    build.build_library()
    return _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "tobac_flow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 10
    handle = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(lib.EXPORTS) == names


def test_no_torch_or_python_linkage(lib):
    import subprocess
    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libpython" not in out and "libc10" not in out


def test_level_plan_matches_oracle(lib):
    from oracle import farneback_np as fb
    for (H, W) in [(1500, 2500), (500, 500), (3712, 3712), (5424, 5424), (100, 100), (10, 15), (130, 170), (63, 64)]:
        assert lib.level_plan(H, W) == [(p["h"], p["w"]) for p in fb.level_plan(H, W)]


def test_poly_constants_match_oracle(lib):
    from oracle import farneback_np as fb
    g, xg, xxg, ig11, ig03, ig33, ig55 = fb.prepare_gaussian(5, 1.1)
    c = lib.poly_constants()
    assert np.array_equal(c[0:6], g[5:]) and np.array_equal(c[6:12], xg[5:]) and np.array_equal(c[12:18], xxg[5:])
    assert np.allclose(c[18:], [ig11, ig03, ig33, ig55], rtol=1e-6)


def test_workspace_and_errors(lib):
    assert lib.workspace_bytes(1, 1500, 2500) > 100 * 1500 * 2500
    assert lib.workspace_bytes(4, 100, 100) >= 4 * lib.workspace_bytes(1, 100, 100) - 4096 * 8
    p = lib.default_params()
    p.win_size = 15
    assert lib.load().tf_fb_level_plan(100, 100, ctypes.byref(p), None, None) < 0
    assert b"win_size" in lib.load().tf_last_error()
    # NULL pointers are rejected before any CUDA call
    rc = lib.load().tf_sl_convolve(None, 1, 0, 0, None, None, None, 0, 4, 4, 0, 0, 1, 0, lib.structure_bytes(np.ones((3, 3, 3))), 0.0, None)
    assert rc == -1


def test_reducer_recognition(lib):
    from tobac_flow_b200 import flow, sobel
    r = flow.recognise_reducer
    assert r(None, 7, np.float32) == lib.TF_RED_NONE
    assert r(lambda x: np.nanmean(x, 0), 3, np.float32) == lib.TF_RED_NANMEAN
    assert r(lambda x: np.nanmean(x, axis=0), 5, np.float64) == lib.TF_RED_NANMEAN
    assert r(lambda x: np.nanmax(x, 0), 7, np.float32) == lib.TF_RED_NANMAX
    assert r(partial(np.any, axis=0), 3, np.int32) == lib.TF_RED_ANY
    assert r(flow.diff_func, 3, np.float32) == lib.TF_RED_DIFF
    assert r(sobel.sobel_reducer("uphill"), 27, None) == lib.TF_RED_SOBEL_UPHILL
    assert r(lambda x: np.nanmedian(x, 0), 3, np.float32) is None
    assert r(lambda x: x.sum(0), 3, np.float32) is None
    assert r(lambda x: np.nanmean(x, 0) + 1e-7, 3, np.float64) is None


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "tobac_flow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "import cv2" not in text, f


def test_no_cuda_raises_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import tobac_flow_b200 as tfb
    with pytest.raises(tfb.NativeError):
        tfb.create_flow(np.zeros((3, 40, 40), np.float32))
    with pytest.raises(tfb.NativeError):
        tfb.Flow(np.zeros((1, 4, 4, 2), np.float32), np.zeros((1, 4, 4, 2), np.float32)).diff(np.zeros((1, 4, 4), np.float32))


def test_synthetic_numpy_and_torch_agree():
    import torch
    from tobac_flow_b200 import synthetic
    a = synthetic.bt_sequence(8, 64, 80, seed=9)
    b = synthetic.bt_sequence(8, 64, 80, seed=9, device="cpu").numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) < 1e-3
    assert np.isnan(a).sum() > 0


def test_detection_abi_argument_checks(lib):
    """The detection / labelling entry points reject bad arguments with a status code before touching CUDA."""
    L = lib.load()
    assert L.tf_ccl_workspace_bytes(0, 10, 10) == 0
    n = L.tf_ccl_workspace_bytes(3, 100, 200)
    assert n >= 3 * 100 * 200 * 5                                  # int32 parents + one flag byte per pixel
    assert L.tf_flat_label(None, None, 1, 4, 4, 1, None, None, 0, None) == -1
    assert L.tf_flat_label(None, None, 0, 4, 4, 1, None, None, 0, None) == 0          # empty stack: nothing to do
    assert L.tf_binary_fill_holes(None, None, 2, 4, 4, None, 0, None) == -1
    w = (ctypes.c_double * 3)(0.25, 0.5, 0.25)
    assert L.tf_gaussian_filter_yx(None, None, None, 0, 1, 4, 4, w, 1, None) == -1
    assert L.tf_curvature_mask(None, None, 0, 1, 4, 4, 0.0, 0, None) == -1
    assert L.tf_binary_opening_cross(None, None, 1, 4, 4, None) == -1
    assert L.tf_grey_opening_cross(None, None, None, 0, 1, 4, 4, None) == -1
    assert L.tf_scale_frames(None, None, None, 1, 4, 4, None) == -1
    assert L.tf_mask_multiply(None, None, None, 0, 16, None) == -1
    assert L.tf_threshold_ge(None, 0.0, None, 0, 16, None) == -1
    assert L.tf_label_max(None, 16, None, None) == -1
    assert L.tf_relabel(None, None, None, 16, 3, None) == -1
    assert L.tf_label_stats(None, None, None, 1, 16, 3, None, None, None, None, None) == -1
    assert L.tf_label_overlap_count(None, None, None, 16, None, 3, None, None, 64, None, None) == -1
    assert b"invalid argument" in L.tf_last_error()
    # the host-side linking walk works without a GPU: two labels linked by one forward edge of 3 pixels
    keys = np.array([(1 << 31) | 2, 0xFFFFFFFFFFFFFFFF], np.uint64)
    counts = np.array([3, 0], np.int32)
    sizes = np.array([0, 5, 4, 2], np.int32)
    out = np.zeros(4, np.int32)
    assert L.tf_label_link_groups(keys.ctypes.data, counts.ctypes.data, 2, sizes.ctypes.data, 3, ctypes.c_double(0.5), 1,
                                  out.ctypes.data) == 2
    assert out.tolist() == [0, 1, 1, 2]
    assert L.tf_label_link_groups(keys.ctypes.data, counts.ctypes.data, 2, sizes.ctypes.data, 3, ctypes.c_double(0.8), 1,
                                  out.ctypes.data) == 3                               # 3 < 0.8 * min(5, 4): not linked
    assert out.tolist() == [0, 1, 2, 3]


def test_detection_host_helpers():
    """Pure host logic of the detection mirror: scipy's Gaussian weights and the centred time differences."""
    from scipy.ndimage import _filters
    from tobac_flow_b200.detection import gaussian_kernel1d, time_diff_minutes
    import pandas as pd
    for sigma in (0.5, 1.0, 2.0, 3.3):
        w, r = gaussian_kernel1d(sigma)
        assert r == int(4.0 * sigma + 0.5)
        assert np.array_equal(w, _filters._gaussian_kernel1d(sigma, 0, r))
    t = pd.to_datetime(["2020-01-01 00:00", "2020-01-01 00:05", "2020-01-01 00:15", "2020-01-01 00:16"])
    assert np.array_equal(time_diff_minutes(t), [5.0, 7.5, 5.5, 1.0])


def test_host_batch_plan():
    """Pair batches of the host path: small first (operators can start while the upload runs), doubling up to the device
    path's batch size; every pair exactly once."""
    from tobac_flow_b200 import flow as tflow
    for n_pairs, big in ((287, 96), (287, 143), (44, 44), (7, 96), (1, 1), (100, 5), (16, 8)):
        plan = tflow._host_batch_plan(n_pairs, big)
        assert sum(plan) == n_pairs and all(0 < n <= big for n in plan)
        assert plan[0] == min(tflow._HOST_PAIR_BATCH, n_pairs, big)
        assert all(b <= 2 * a for a, b in zip(plan, plan[1:-1]))        # at most doubling (the last batch is the remainder)
    assert tflow._host_batch_plan(287, 96) == [8, 16, 32, 64, 96, 71]
    assert tflow._host_batch_plan(0, 96) == []
