"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_golden.py

It imports ``/root/reference/tobac_flow`` under the dependency stubs of ``refshim.py`` and writes small
``.npz`` fixtures next to this file.  Inputs are regenerated from seeds by
``tobac_flow_b200.synthetic`` and are NOT stored.  The fixtures record the versions they came from.
"""
import hashlib
import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402
from tobac_flow_b200 import synthetic  # noqa: E402


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def small_bt():
    """6 x 64 x 96 BT-like stack with NaN pixels, a NaN stripe and one all-NaN frame."""
    bt = synthetic.bt_sequence(6, 64, 96, seed=42, nans=False)
    rng = np.random.default_rng(7)
    idx = rng.integers(0, bt.size, 40)
    bt.reshape(-1)[idx] = np.nan
    bt[2, 20:24] = np.nan
    bt[4] = np.nan
    return bt


def three_level():
    return synthetic.bt_sequence(3, 130, 170, seed=5, nans=False)


def growth_case():
    """12 x 120 x 160 wvd-like field with one fast-growing core (drives detect_growth_markers)."""
    T, H, W = 12, 120, 160
    base = synthetic.base_field(H, W, 99, sigma_px=10.0)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    wvd = np.empty((T, H, W), np.float32)
    for t in range(T):
        bg = -25.0 + 6.0 * np.roll(base, (t, 2 * t), (0, 1))
        g = min(t, 8) / 8.0
        cy, cx = 50 + t, 60 + 2 * t
        core = (22.0 * g) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * (5.0 + 5.0 * g) ** 2))
        wvd[t] = bg + core
    return wvd.astype(np.float32)


def growth_multi_case():
    """14 x 120 x 160 wvd-like field with several cores: a long strong one, a short-lived one (dropped by the length
    filter), a weak one (dropped by the >= 0.5 mask), a cold one (dropped by the wvd >= -5 mask) and two that merge."""
    T, H, W = 14, 120, 160
    base = synthetic.base_field(H, W, 123, sigma_px=10.0)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    #        y0  x0  t_start t_len depth  offset
    cores = [(30, 30, 1, 9, 24.0, 0.0),      # strong, long
             (30, 110, 4, 1, 6.0, 0.0),      # one-step jump: two frames above 0.25 -> dropped by the length filter
             (85, 40, 2, 9, 16.0, 0.0),      # weak growth: never reaches 0.5
             (90, 120, 1, 9, 30.0, -16.0),   # grows fast inside a cold patch: stays below -5
             (60, 70, 2, 8, 22.0, 0.0),      # merges with the next one
             (60, 84, 4, 8, 22.0, 0.0)]
    wvd = np.empty((T, H, W), np.float32)
    for t in range(T):
        f = -25.0 + 5.0 * np.roll(base, (t, 2 * t), (0, 1))
        for (y0, x0, ts, tl, depth, off) in cores:
            g = float(np.clip((t - ts) / tl, 0.0, 1.0))
            r2 = (yy - (y0 + t)) ** 2 + (xx - (x0 + 2 * t)) ** 2
            f = f + depth * g * np.exp(-r2 / (2 * (4.0 + 5.0 * g) ** 2)) + off * np.exp(-r2 / (2 * 12.0 ** 2))
        wvd[t] = f
    return wvd.astype(np.float32)


def make_growth_multi(meta):
    import pandas as pd
    from scipy import ndimage as ndi
    from tobac_flow.flow import create_flow, Flow
    from tobac_flow import detection
    from tobac_flow.label import flow_label
    from tobac_flow.utils.label_utils import flat_label

    wvd = growth_multi_case()
    t = pd.date_range("2020-01-01", periods=wvd.shape[0], freq="5min")
    fg = create_flow(wvd)
    fq = np.round(fg.forward_flow * 256).astype(np.int16)
    bq = np.round(fg.backward_flow * 256).astype(np.int16)
    fg = Flow(fq.astype(np.float32) / 256, bq.astype(np.float32) / 256)
    da = refshim.DataArray(wvd, coords={"t": t}, dims=("t", "y", "x"), t=t)
    smoothed, markers = detection.detect_growth_markers(fg, da)
    markers = np.asarray(markers.data if hasattr(markers, "data") else markers)
    s2 = ndi.generate_binary_structure(2, 1)[np.newaxis, ...]
    curv = detection.get_curvature_filter(wvd)
    opened = ndi.grey_opening(smoothed, footprint=s2)
    filtered = opened * curv
    seeds = ndi.binary_opening(filtered >= 0.25, structure=s2)
    flat = flat_label(seeds != 0)
    linked = fg.label(seeds)
    linked_ov = flow_label(fg, seeds, overlap=0.5, absolute_overlap=4)
    by_len = detection.filter_labels_by_length(linked, 3)
    gauss = ndi.gaussian_filter(wvd, (0, 2, 2))
    np.savez_compressed(os.path.join(HERE, "growth_multi.npz"), meta=str(meta), fwd_q256=fq, bwd_q256=bq,
                        smoothed_even=smoothed[::2].astype(np.float32), gauss_3=gauss[3], opened_5=opened[5],
                        curv=np.packbits(curv), mask025=np.packbits(filtered >= 0.25), mask05=np.packbits(filtered >= 0.5),
                        seeds=np.packbits(seeds), flat=flat.astype(np.int16), linked=linked.astype(np.int16),
                        linked_ov=linked_ov.astype(np.int16), by_len=by_len.astype(np.int16),
                        markers=markers.astype(np.int16))
    print("growth_multi: flat", int(flat.max()), "linked", int(linked.max()), "linked_ov", int(linked_ov.max()),
          "by_len", int(by_len.max()), "markers", int(markers.max()), "px", int((markers > 0).sum()))


def main():
    import cv2
    import scipy
    import pandas as pd

    refshim.load_reference()
    from tobac_flow.flow import create_flow, Flow, smooth_flow_step
    from tobac_flow import detection
    from tobac_flow.utils import to_8bit, linear_norm
    from scipy import ndimage as ndi

    meta = dict(cv2=cv2.__version__, numpy=np.__version__, scipy=scipy.__version__)
    if "--only-growth-multi" in sys.argv:
        make_growth_multi(meta)
        return

    # ---- G1: the survey's known-answer case -------------------------------------------------
    data = synthetic.blob_stack()
    f = create_flow(data)
    d = f.diff(data)
    s = f.sobel(data)
    s_up_cubic = f.sobel(data, direction="uphill", method="cubic")
    s_near = f.sobel(data, method="nearest")
    c = f.convolve(data)
    np.savez_compressed(
        os.path.join(HERE, "blob100.npz"),
        meta=str(meta),
        fwd_0_4_9=f.forward_flow[[0, 4, 9]].astype(np.float32), bwd_4=f.backward_flow[4],
        diff_4=d[4], sobel_4=s[4], sobel_up_cubic_4=s_up_cubic[4], sobel_near_4=s_near[4],
        conv_nan_counts=np.isnan(c).sum((1, 2, 3)),
        sha=np.array([sha16(f.forward_flow), sha16(f.backward_flow), sha16(d), sha16(s), sha16(c)]),
    )

    # ---- small BT stack with NaNs: every operator, reference flow stored ---------------------
    bt = small_bt()
    f = create_flow(bt)
    q = np.stack([np.stack(to_8bit(linear_norm(bt[i:i + 2]), 0, 1)) for i in range(bt.shape[0] - 1)])
    out = dict(meta=str(meta), q=q, fwd=f.forward_flow, bwd=f.backward_flow)
    out["diff"] = f.diff(bt)
    out["diff_nearest"] = f.diff(bt, method="nearest")
    out["sobel"] = f.sobel(bt)
    out["sobel_f32"] = f.sobel(bt, dtype=np.float32)
    out["sobel_uphill_cubic"] = f.sobel(bt, direction="uphill", method="cubic")
    out["sobel_downhill"] = f.sobel(bt, direction="downhill")
    out["conv7_t13"] = f.convolve(bt)[:, [1, 3]]
    out["conv7_cubic_fill0_t13"] = f.convolve(bt, method="cubic", fill_value=0.0)[:, [1, 3]]
    t_struct = np.zeros([3, 3, 3])
    t_struct[:, 1, 1] = 1
    raw64 = out["diff"] / np.full(bt.shape[0], 5.0)[:, None, None]  # float64, as detection.py:99-101
    out["tmean_f64src"] = f.convolve(raw64, structure=t_struct, func=lambda x: np.nanmean(x, 0))
    s_struct = ndi.generate_binary_structure(3, 1)
    s_struct[0] = 0
    s_struct[2] = 0
    out["smean"] = f.convolve(out["diff"], structure=s_struct, func=lambda x: np.nanmean(x, 0))
    mask = (np.nan_to_num(bt, nan=300.0) < 250).astype(int)
    out["any_nearest"] = f.convolve(mask, structure=t_struct.astype(bool), method="nearest",
                                    fill_value=False, dtype=np.int32, func=partial(np.any, axis=0))
    labels = ndi.label(mask)[0].astype(np.int32)
    l_struct = ndi.generate_binary_structure(3, 1)
    l_struct[1] = 0
    out["labels_nearest"] = f.convolve(labels, structure=l_struct, method="nearest", dtype=np.int32,
                                       fill_value=0)
    # production flow settings minus variational refinement: one smoothing pass, cubic
    fs = create_flow(bt, smoothing_passes=1, interp_method="cubic")
    out["fwd_smooth1_cubic"] = fs.forward_flow
    out["bwd_smooth1_cubic"] = fs.backward_flow
    np.savez_compressed(os.path.join(HERE, "bt_small.npz"), **out)

    # ---- production flow settings of scripts/dcc_detect_goes.py:164-166 ------------------------------------
    fp = create_flow(bt, model="Farneback", vr_steps=1, smoothing_passes=1, interp_method="cubic")
    fv = create_flow(bt, vr_steps=1)
    np.savez_compressed(os.path.join(HERE, "bt_small_production.npz"), meta=str(meta),
                        fwd_vr1_smooth1_cubic=fp.forward_flow, bwd_vr1_smooth1_cubic=fp.backward_flow,
                        fwd_vr1=fv.forward_flow, bwd_vr1=fv.backward_flow)

    # ---- three pyramid levels with half-even level sizes --------------------------------------
    bt3 = three_level()
    f3 = create_flow(bt3)
    np.savez_compressed(os.path.join(HERE, "three_level.npz"), meta=str(meta),
                        fwd_0=f3.forward_flow[0], bwd_1=f3.backward_flow[1])

    # ---- downstream consumer: detect_growth_markers -------------------------------------------
    wvd = growth_case()
    t = pd.date_range("2020-01-01", periods=wvd.shape[0], freq="5min")
    fg = create_flow(wvd)
    # store the flow on a 1/256 px grid (int16 compresses well); the reference is then run on exactly
    # the stored, de-quantised flow so both sides of the parity test see identical flow fields
    fq = np.round(fg.forward_flow * 256).astype(np.int16)
    bq = np.round(fg.backward_flow * 256).astype(np.int16)
    fg = Flow(fq.astype(np.float32) / 256, bq.astype(np.float32) / 256)
    da = refshim.DataArray(wvd, coords={"t": t}, dims=("t", "y", "x"), t=t)
    # label.py warns when int32 would overflow; detection needs xr.DataArray isinstance to be our stub
    smoothed, markers = detection.detect_growth_markers(fg, da)
    markers = np.asarray(markers.data if hasattr(markers, "data") else markers)
    s2 = ndi.generate_binary_structure(2, 1)[np.newaxis, ...]
    filtered = ndi.grey_opening(smoothed, footprint=s2) * detection.get_curvature_filter(wvd)
    np.savez_compressed(os.path.join(HERE, "growth.npz"), meta=str(meta),
                        fwd_q256=fq, bwd_q256=bq,
                        smoothed_even=smoothed[::2].astype(np.float32),
                        mask025=np.packbits(filtered >= 0.25), mask05=np.packbits(filtered >= 0.5), markers=markers.astype(np.int32),
                        n_markers=np.array(int(markers.max())))
    make_growth_multi(meta)
    print("growth markers:", int(markers.max()), "marked px:", int((markers > 0).sum()))
    for n in ("blob100", "bt_small", "bt_small_production", "three_level", "growth"):
        print(n, os.path.getsize(os.path.join(HERE, n + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
