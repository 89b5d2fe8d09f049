"""CPU oracle for the "next" rows of the hot path: the per-frame filters of ``detect_growth_markers`` and the
semi-Lagrangian labelling.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py``; the product package never imports it.

This restates the reference's Python for those rows.  The per-pixel arithmetic of these rows lives in
``scipy.ndimage`` (third party, not vendored in the reference; ``environment.yml`` lists ``scipy`` unpinned, this
image has 1.18), which the oracle calls exactly as the reference does; the Flow operators between the filters come
from ``oracle.flow_ops``.  Pinned against the unmodified reference by ``tests/golden/growth*.npz``
(``tests/test_oracle_detection.py``).

Reference lines followed:
  flat_label                tobac_flow/utils/label_utils.py:143-180
  find_overlapping_labels   tobac_flow/utils/label_utils.py:352-376
  flow_label / link         tobac_flow/label.py:84-175, 179-246, 249-321
  filter_labels_by_length   tobac_flow/analysis.py:66-75
  filter_labels_by_mask     tobac_flow/analysis.py:78-86
  filtered_tdiff            tobac_flow/detection.py:34-60
  get_curvature_filter      tobac_flow/detection.py:64-94
  detect_growth_markers     tobac_flow/detection.py:98-125
  nan_gaussian_filter       tobac_flow/detection.py:128-146
  detect_growth_markers_multichannel   tobac_flow/detection.py:203-254
                            (filter_labels_by_length_and_multimask_legacy analysis.py:182-201)
  get_growth_rate           tobac_flow/detection.py:168-198
  get_anvil_markers         tobac_flow/detection.py:494-516 (find_object_lengths analysis.py:15-35,
                            remap_labels utils/label_utils.py:265-307)
"""
import numpy as np
from scipy import ndimage as ndi

from . import flow_ops as ops


def cross3():
    return ndi.generate_binary_structure(3, 1)


def flat_label(mask, structure=None, dtype=np.int32):
    """label_utils.py:143-180: ndi.label with the structure's time links removed."""
    s = (cross3() if structure is None else np.asarray(structure)).copy()
    s[0] = 0
    s[-1] = 0
    return ndi.label(mask, structure=s, output=dtype)[0]


def find_overlapping_labels(labels, locs, bins, overlap=0, absolute_overlap=0):
    """label_utils.py:352-376."""
    n_locs = len(locs)
    if n_locs == 0:
        return []
    overlap_labels = labels.ravel()[locs]
    overlap_bins = np.bincount(np.maximum(overlap_labels, 0))
    return [m for m in np.unique(overlap_labels)
            if m != 0 and overlap_bins[m] > absolute_overlap
            and overlap_bins[m] >= overlap * np.minimum(n_locs, bins[m] - bins[m - 1])]


def link_flat_labels(flat_labels, back_labels, forward_labels, dtype=np.int32, overlap=0.0, absolute_overlap=0):
    """The linking walk of label.py:139-170 (and :301-316): breadth-first over forward then backward overlaps."""
    bins = np.cumsum(np.bincount(flat_labels.ravel()))
    args = np.argsort(flat_labels.ravel(), kind="stable")
    processed = np.zeros(bins.size, dtype=bool)
    label_map = {}
    for label in range(1, bins.size):
        if processed[label]:
            continue
        stack = label_map[label] = [label]
        processed[label] = True
        i = 0
        while i < len(stack):
            cur = stack[i]
            if bins[cur] > bins[cur - 1]:
                locs = args[bins[cur - 1]:bins[cur]]
                for nb in (forward_labels, back_labels):
                    for m in find_overlapping_labels(nb, locs, bins, overlap, absolute_overlap):
                        if not processed[m]:
                            stack.append(m)
                            processed[m] = True
            i += 1
    new_labels = np.zeros(flat_labels.shape, dtype=dtype)
    for ik, k in enumerate(label_map):
        for i in label_map[k]:
            if bins[i] > bins[i - 1]:
                new_labels.ravel()[args[bins[i - 1]:bins[i]]] = ik + 1
    return new_labels


def label_taps(flat_labels, fwd, bwd, structure=None, dtype=np.int32, backend="numpy"):
    """label.py:131-137: the (back, forward) nearest-neighbour gathers of the flat labels."""
    structure = cross3() if structure is None else np.asarray(structure)
    label_struct = structure * np.array([1, 0, 1])[:, np.newaxis, np.newaxis]
    back, forward = ops.convolve(flat_labels, fwd, bwd, structure=label_struct, method="nearest", dtype=dtype,
                                 fill_value=0, backend=backend)
    return back, forward


def flow_label(mask, fwd, bwd, structure=None, dtype=np.int32, overlap=0.0, absolute_overlap=0, backend="numpy"):
    """label.py:84-175 with subsegment_shrink == 0."""
    structure = cross3() if structure is None else np.asarray(structure)
    flat = flat_label(np.asarray(mask) != 0, structure=structure).astype(dtype)
    back, forward = label_taps(flat, fwd, bwd, structure, dtype, backend)
    return link_flat_labels(flat, back, forward, dtype, overlap, absolute_overlap)


def flow_link_overlap(flat_labels, fwd, bwd, structure=None, dtype=np.int32, overlap=0.0, absolute_overlap=0,
                      backend="numpy"):
    """label.py:249-321."""
    back, forward = label_taps(flat_labels, fwd, bwd, structure, dtype, backend)
    return link_flat_labels(flat_labels, back, forward, dtype, overlap, absolute_overlap)


def _renumber(labels, wh):
    remap = np.zeros([np.nanmax(labels) + 1], labels.dtype)
    remap[1:] = np.cumsum(wh) * wh
    return remap[labels]


def filter_labels_by_length(labels, min_length):
    """analysis.py:66-75."""
    wh = np.array([o[0].stop - o[0].start for o in ndi.find_objects(labels)]) >= min_length
    return _renumber(labels, wh)


def filter_labels_by_mask(labels, mask):
    """analysis.py:78-86."""
    wh = ndi.labeled_comprehension(mask, labels, range(1, np.nanmax(labels) + 1), np.any, None, None)
    return _renumber(labels, wh)


def filtered_tdiff(raw_diff, fwd, bwd, backend="numpy"):
    """detection.py:34-60."""
    t_struct = np.zeros([3, 3, 3])
    t_struct[:, 1, 1] = 1
    return ops.convolve(raw_diff, fwd, bwd, structure=t_struct, func=ops.nanmean_reducer, backend=backend)


def curvature_mask(field, sigma=2, threshold=0, direction="negative"):
    """detection.py:65-70 and the comparison of :77 / :85 (before fill_holes / opening)."""
    smoothed = ndi.gaussian_filter(field, (0, sigma, sigma))
    x_diff = np.zeros(field.shape)
    x_diff[:, :, 1:-1] = np.diff(smoothed, n=2, axis=2)
    y_diff = np.zeros(field.shape)
    y_diff[:, 1:-1] = np.diff(smoothed, n=2, axis=1)
    if direction == "negative":
        return np.logical_and(x_diff < -threshold, y_diff < -threshold)
    if direction == "positive":
        return np.logical_and(x_diff > threshold, y_diff > threshold)
    raise ValueError("Direction must be either positive or negative")


def get_curvature_filter(field, sigma=2, threshold=0, direction="negative"):
    """detection.py:64-94."""
    s_struct = cross3()
    s_struct[0] = 0
    s_struct[2] = 0
    m = curvature_mask(field, sigma, threshold, direction)
    return ndi.binary_opening(ndi.binary_fill_holes(m, structure=s_struct), structure=s_struct)


def detect_growth_markers(wvd, dt_minutes, fwd, bwd, backend="numpy", intermediates=False):
    """detection.py:98-125 on a plain (T, H, W) array; ``dt_minutes`` = get_time_diff_from_coord(wvd.t)."""
    wvd = np.asarray(wvd)
    raw = ops.diff(wvd, fwd, bwd, backend=backend) / np.asarray(dt_minutes, np.float64)[:, np.newaxis, np.newaxis]
    smoothed = filtered_tdiff(raw, fwd, bwd, backend=backend)
    s_struct = ndi.generate_binary_structure(2, 1)[np.newaxis, ...]
    filtered = ndi.grey_opening(smoothed, footprint=s_struct) * get_curvature_filter(wvd)
    seeds = ndi.binary_opening(filtered >= 0.25, structure=s_struct)
    linked = flow_label(seeds, fwd, bwd, absolute_overlap=1, backend=backend)    # Flow.label defaults, flow.py:281-290
    markers = filter_labels_by_length(linked, 3) if linked.max() > 0 else linked
    if markers.max() > 0:
        markers = filter_labels_by_mask(markers, filtered >= 0.5)
    if markers.max() > 0:
        markers = filter_labels_by_mask(markers, wvd >= -5)
    if intermediates:
        return dict(raw=raw, smoothed=smoothed, filtered=filtered, seeds=seeds, linked=linked, markers=markers)
    return smoothed, markers


def nan_gaussian_filter(a, *args, propagate_nan=True, **kwargs):
    """detection.py:128-146."""
    wh_nan = np.isnan(a)
    a0 = a.copy()
    a0[wh_nan] = 0
    c = np.ones_like(a)
    c[wh_nan] = 0
    a0_gaussian = ndi.gaussian_filter(a0, *args, **kwargs)
    c_gaussian = ndi.gaussian_filter(c, *args, **kwargs)
    c_gaussian[c_gaussian == 0] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        result = a0_gaussian / c_gaussian
    if propagate_nan:
        result[wh_nan] = np.nan
    return result


def filter_labels_by_length_and_multimask_legacy(labels, masks, min_length):
    """analysis.py:182-201 (labels is renumbered in place by the reference; a copy here)."""
    if type(masks) is not type(list()):
        raise ValueError("masks input must be a list of masks to process")
    labels = labels.copy()
    bins = np.cumsum(np.bincount(labels.ravel()))
    args = np.argsort(labels.ravel())
    object_lengths = np.array([o[0].stop - o[0].start for o in ndi.find_objects(labels)])
    counter = 1
    for i in range(bins.size - 1):
        if bins[i + 1] > bins[i]:
            sel = args[bins[i]:bins[i + 1]]
            if object_lengths[i] >= min_length and np.all([np.any(m.ravel()[sel]) for m in masks]):
                labels.ravel()[sel] = counter
                counter += 1
            else:
                labels.ravel()[sel] = 0
    return labels


def detect_growth_markers_multichannel(wvd, bt, dt_wvd, dt_bt, fwd, bwd, overlap=0.5, min_length=4, lower_threshold=0.25,
                                       upper_threshold=0.5, backend="numpy"):
    """detection.py:203-254 on plain (T, H, W) arrays (subsegment_shrink = 0)."""
    wvd, bt = np.asarray(wvd), np.asarray(bt)
    wvd_s = filtered_tdiff(ops.diff(wvd, fwd, bwd, backend=backend)
                           / np.asarray(dt_wvd, np.float64)[:, np.newaxis, np.newaxis], fwd, bwd, backend=backend)
    bt_s = filtered_tdiff(ops.diff(bt, fwd, bwd, backend=backend)
                          / np.asarray(dt_bt, np.float64)[:, np.newaxis, np.newaxis], fwd, bwd, backend=backend)
    with np.errstate(invalid="ignore"):
        markers = np.logical_or((wvd_s * get_curvature_filter(wvd)) >= lower_threshold,
                                (bt_s * get_curvature_filter(bt, direction="positive")) <= -lower_threshold)
    seeds = ndi.binary_opening(markers, structure=ndi.generate_binary_structure(2, 1)[np.newaxis, ...])
    markers = flow_label(seeds, fwd, bwd, overlap=overlap, absolute_overlap=1, backend=backend)
    if np.count_nonzero(markers) > 0:
        with np.errstate(invalid="ignore"):
            markers = filter_labels_by_length_and_multimask_legacy(
                markers, [wvd_s >= upper_threshold, bt_s <= -upper_threshold, wvd > -5], min_length)
    return wvd_s, bt_s, markers


def get_growth_rate(field, dt_minutes, fwd, bwd, method="linear", backend="numpy"):
    """detection.py:168-198: time derivative, then the nanmean over the 5-point same-step cross."""
    field = np.asarray(field)
    growth = ops.diff(field, fwd, bwd, method=method, backend=backend) / np.asarray(dt_minutes, np.float64)[:, np.newaxis, np.newaxis]
    s_struct = cross3()
    s_struct[0] = 0
    s_struct[2] = 0
    return ops.convolve(growth, fwd, bwd, structure=s_struct, method=method, func=ops.nanmean_reducer, backend=backend)


def find_object_lengths(labels):
    """analysis.py:15-35."""
    return np.array([o[0].stop - o[0].start for o in ndi.find_objects(labels)])


def remap_labels(labels, locations):
    """utils/label_utils.py:265-307 for a boolean `locations` (the only form the detection path uses)."""
    remapper = np.zeros(np.nanmax(labels) + 1, labels.dtype)
    remapper[1:][locations] = np.arange(1, np.sum(locations) + 1)
    return remapper[labels]


def get_anvil_markers(field, fwd, bwd, threshold=-5, overlap=0.5, absolute_overlap=5, min_length=3, backend="numpy"):
    """detection.py:494-516 with subsegment_shrink == 0."""
    structure = cross3()
    s_struct = structure * np.array([0, 1, 0])[:, np.newaxis, np.newaxis].astype(bool)
    mask = ndi.binary_opening(np.asarray(field) >= threshold, structure=s_struct)
    labels = flow_label(mask, fwd, bwd, overlap=overlap, absolute_overlap=absolute_overlap, backend=backend)
    if labels.max() == 0:
        return labels
    return remap_labels(labels, find_object_lengths(labels) > min_length)
