"""CPU restatement of the reference's semi-Lagrangian watershed.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``watershed_raveled`` follows ``/root/reference/tobac_flow/_watershed.pyx:222-344`` for compactness 0 / no watershed
lines (the reference's only call, ``tobac_flow/watershed.py:147-160``), including its binary heap (marker pixels share age 0, so
the pop order of equal-valued markers is a property of the heap procedure, not of the (value, age) key).  ``watershed`` follows
``tobac_flow/watershed.py:17-168`` with the two scikit-image helpers restated (scikit-image is not installed in this
image: ``_validate_connectivity`` / ``_offsets_to_raveled_neighbors`` of skimage/morphology/_util.py, restated from
their documented behaviour -- that part of the oracle is pinned only through the reference's compiled flood,
``oracle/build_ref_watershed.py``, not through scikit-image itself).
"""
import numpy as np
from scipy import ndimage as ndi


def _smaller(a, b):
    """_watershed.pyx:161-164: by value, then by age."""
    if a[0] != b[0]:
        return a[0] < b[0]
    return a[1] < b[1]


def _heappush(h, e):
    """_watershed.pyx:117-151: append, sift up while the child is smaller than its parent."""
    h.append(e)
    child = len(h) - 1
    while child > 0:
        parent = (child + 1) // 2 - 1
        if _smaller(h[child], h[parent]):
            h[child], h[parent] = h[parent], h[child]
            child = parent
        else:
            break


def _heappop(h):
    """_watershed.pyx:64-108: the last element replaces the root and sifts down towards the smaller child."""
    top = h[0]
    last = h.pop()
    n = len(h)
    if n == 0:
        return top
    h[0] = last
    i = 0
    while True:
        l, r = 2 * i + 1, 2 * i + 2
        if l >= n:
            break
        smallest = i
        if _smaller(h[l], h[i]):
            smallest = l
        if r < n and _smaller(h[r], h[smallest]):
            smallest = r
        if smallest == i:
            break
        h[i], h[smallest] = h[smallest], h[i]
        i = smallest
    return top


def watershed_raveled(image, marker_locations, structure, forward_offset, backward_offset, forward_offset_locations,
                      backward_offset_locations, mask, output):
    """The flood with the reference's own binary heap: marker pixels all carry age 0, so equal-valued markers are NOT
    strictly ordered by (value, age) and their pop order is a property of the heap procedure, which is restated here."""
    heap = []
    for index in marker_locations:
        _heappush(heap, (np.float32(image[index]), 0, int(index)))
    age = 1
    nn = len(structure)
    while heap:
        _, _, index = _heappop(heap)
        for i in range(nn):
            nb = (int(structure[i]) + index + int(forward_offset_locations[i]) * int(forward_offset[index])
                  + int(backward_offset_locations[i]) * int(backward_offset[index]))
            if not mask[nb] or output[nb]:
                continue
            age += 1
            output[nb] = output[index]
            _heappush(heap, (np.float32(image[nb]), age, nb))
    return output


def offsets_to_raveled_neighbors(image_shape, footprint, center):
    offsets = np.stack([idx - c for idx, c in zip(np.nonzero(footprint), center)], axis=-1)
    ravel_factors = np.cumprod((tuple(image_shape[1:]) + (1,))[::-1])[::-1]
    raveled = (offsets * ravel_factors).sum(axis=1)
    distances = np.sqrt((offsets.astype(float) ** 2).sum(axis=1))
    return raveled[np.argsort(distances, kind="stable")][1:]


def watershed(forward_flow, backward_flow, field, markers, mask=None, connectivity=1, flood=watershed_raveled):
    field = np.asarray(field, np.float32)
    markers = np.asarray(markers, np.int32)
    mask = np.ones(field.shape, np.int8) if mask is None else np.asarray(mask, np.int8)
    footprint = ndi.generate_binary_structure(field.ndim, connectivity) if np.isscalar(connectivity) else np.asarray(connectivity, bool)
    offset = np.array(footprint.shape) // 2
    pad_offset = offset.copy()
    pad_offset[1] += int(max(np.max(np.round(np.abs(forward_flow[..., 1]))), np.max(np.round(np.abs(backward_flow[..., 1])))))
    pad_offset[2] += int(max(np.max(np.round(np.abs(forward_flow[..., 0]))), np.max(np.round(np.abs(backward_flow[..., 0])))))
    pad_width = [(p, p) for p in pad_offset]
    field_p = np.pad(field, pad_width, mode="constant")
    mask_p = np.pad(mask, pad_width, mode="constant").ravel()
    output = np.pad(markers, pad_width, mode="constant")
    flat = offsets_to_raveled_neighbors(field_p.shape, footprint, offset)
    marker_locations = np.flatnonzero(output)
    strides = np.array(field_p.strides, dtype=np.int32) // field_p.itemsize

    def rav(flow):
        return (np.pad(np.round(flow[..., 0]).astype(np.int32), pad_width, mode="constant").ravel() * strides[2]
                + np.pad(np.round(flow[..., 1]).astype(np.int32), pad_width, mode="constant").ravel() * strides[1])

    fo, bo = rav(forward_flow), rav(backward_flow)
    fl = (np.round(flat / strides[0]) == 1).astype(np.int32)
    bl = (np.round(flat / strides[0]) == -1).astype(np.int32)
    out_flat = output.ravel()
    flood(field_p.ravel(), marker_locations.astype(np.intp), flat.astype(np.intp), fo.astype(np.int32), bo.astype(np.int32),
          fl, bl, mask_p, out_flat)
    sl = tuple(slice(a, s - b) for (a, b), s in zip(pad_width, output.shape))
    return output[sl].copy()
