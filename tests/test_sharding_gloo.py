"""world_size-2 (and 3) gloo tests of the time-sharding logic on CPU.

The compute back-end is the oracle (allowed in tests); what is under test is the host logic of
``tobac_flow_b200.distributed``: shard bounds, halo frames, the backward-flow hand-over, end rules per rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flow_ops as ops
from tobac_flow_b200 import distributed as D
from tobac_flow_b200 import synthetic

T, H, W = 7, 40, 48


class OracleOps:
    def calculate_flow(self, frames, fwd, bwd, smoothing_passes, interp_method, max_value):
        fr = frames.numpy()
        for i in range(fr.shape[0] - 1):
            q0, q1 = ops.pair_to_u8(fr[i], fr[i + 1])
            f, b = ops.farneback_pair(q0, q1, "numpy")
            if max_value is not None and smoothing_passes == 0:
                f, b = np.clip(f, -max_value, max_value), np.clip(b, -max_value, max_value)
            fwd[i] = torch.from_numpy(f)
            bwd[i + 1] = torch.from_numpy(b)

    def finalise(self, fwd, bwd, max_value, clamp_all, mirror_first, mirror_last):
        if mirror_last:
            fwd[-1] = -bwd[-1]
        if mirror_first:
            bwd[0] = -fwd[0]

    def convolve(self, data, fwd, bwd, structure, method, fill_value, dtype, reducer, has_prev, has_next, out=None):
        d = data.numpy()
        n = d.shape[0] - int(has_prev) - int(has_next)
        res = []
        for i in range(n):
            j = i + int(has_prev)
            blank = np.full(d[j].shape, fill_value, dtype=dtype)
            prev = d[j - 1] if j > 0 else blank
            nxt = d[j + 1] if j < d.shape[0] - 1 else blank
            stack = ops.tap_stack(prev, d[j], nxt, fwd[i].numpy(), bwd[i].numpy(), structure, method, dtype, fill_value)
            r = (ops.nanmean_reducer(stack) if reducer == 2 else ops.diff_reducer(stack)).astype(dtype)
            r[np.isnan(d[j])] = fill_value
            res.append(r)
        return torch.from_numpy(np.stack(res))


    # -- labelling back-end: scipy / numpy / the oracle's nearest-neighbour gather -----------------------------------
    def flat_label(self, mask_u8, connectivity):
        from oracle import detection_ops as det
        from scipy import ndimage as ndi
        s = ndi.generate_binary_structure(3, 1) if connectivity == 1 else np.ones((3, 3, 3), bool)
        lab = det.flat_label(mask_u8.numpy() != 0, s)
        return torch.from_numpy(lab.astype(np.int32)), int(lab.max())

    def overlap_table(self, flat_view, fwd, bwd, label_struct, has_prev, has_next, n_labels):
        d = flat_view.numpy()
        n = d.shape[0] - int(has_prev) - int(has_next)
        keys, counts = [], []
        local = d[int(has_prev):int(has_prev) + n]
        back = np.zeros_like(local)
        forw = np.zeros_like(local)
        for i in range(n):
            j = i + int(has_prev)
            blank = np.zeros(d[j].shape, np.int32)
            prev = d[j - 1] if j > 0 else blank
            nxt = d[j + 1] if j < d.shape[0] - 1 else blank
            back[i] = ops.warp_image(prev, bwd[i].numpy(), "nearest", 0, 0, 0, "numpy")
            forw[i] = ops.warp_image(nxt, fwd[i].numpy(), "nearest", 0, 0, 0, "numpy")
        for direction, nb in ((0, forw), (1, back)):
            sel = (local > 0) & (nb > 0)
            k = (np.uint64(direction) << np.uint64(62)) | (local[sel].astype(np.uint64) << np.uint64(31)) | nb[sel].astype(np.uint64)
            u, c = np.unique(k, return_counts=True)
            keys.append(u)
            counts.append(c.astype(np.int32))
        sizes = np.bincount(local.ravel(), minlength=n_labels + 1).astype(np.int32)
        sizes[0] = 0
        return np.concatenate(keys), np.concatenate(counts), sizes

    def relabel(self, flat, mapping):
        return torch.from_numpy(mapping[flat.numpy()].astype(np.int32))

    # -- detection back-end: scipy.ndimage, as the reference calls it --------------------------------------------------
    def scale_frames(self, raw32, dt_minutes):
        return torch.from_numpy(raw32.numpy() / np.asarray(dt_minutes, np.float64)[:, None, None])

    def growth_seed_masks(self, smoothed, wvd):
        from oracle import detection_ops as det
        from scipy import ndimage as ndi
        s2 = ndi.generate_binary_structure(2, 1)[np.newaxis, ...]
        filtered = ndi.grey_opening(smoothed.numpy(), footprint=s2) * det.get_curvature_filter(wvd.numpy())
        seeds = ndi.binary_opening(filtered >= 0.25, structure=s2)
        as_u8 = lambda a: torch.from_numpy(a.astype(np.uint8))
        return as_u8(filtered >= 0.5), as_u8(wvd.numpy() >= -5), as_u8(seeds)

    def label_stats(self, labels, mask_a, mask_b, n_labels):
        lab, a, b = labels.numpy(), mask_a.numpy() != 0, mask_b.numpy() != 0
        tmin = np.full(n_labels + 1, 0x7f7f7f7f, np.int32)
        tmax = np.full(n_labels + 1, -1, np.int32)
        any_a = np.zeros(n_labels + 1, np.int32)
        any_b = np.zeros(n_labels + 1, np.int32)
        for t in range(lab.shape[0]):
            for l in np.unique(lab[t]):
                if l > 0:
                    tmin[l] = min(tmin[l], t)
                    tmax[l] = max(tmax[l], t)
        any_a[np.unique(lab[a])] = 1
        any_b[np.unique(lab[b])] = 1
        any_a[0] = any_b[0] = 0
        return tmin, tmax, any_a, any_b


def _label_case():
    """Masks of a drifting blob field plus the flow that advects it (so labels link across frames and shards)."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(31)
    base = ndi.gaussian_filter(rng.standard_normal((H + 30, W + 30)), 2.0)
    mask = np.stack([np.roll(base, (t, 2 * t), (0, 1))[15:15 + H, 15:15 + W] > 0.03 for t in range(T)])
    mask[3, 10:20] = False
    fwd = (rng.standard_normal((T, H, W, 2)) * 0.2 + np.array([2.0, 1.0])).astype(np.float32)
    bwd = (rng.standard_normal((T, H, W, 2)) * 0.2 - np.array([2.0, 1.0])).astype(np.float32)
    return mask, fwd, bwd


def _label_worker(rank, world, port, q, overlap, absolute):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mask, fwd, bwd = _label_case()
        t0, t1 = D.shard_bounds(T, world, rank)
        fl = D.ShardedFlow(torch.from_numpy(fwd[t0:t1].copy()), torch.from_numpy(bwd[t0:t1].copy()), rank, world,
                           ops=OracleOps())
        lab = fl.label(torch.from_numpy(mask[t0:t1].copy()), overlap=overlap, absolute_overlap=absolute)
        q.put((rank, t0, t1, lab.numpy().copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap,absolute", [(2, 0.0, 1), (3, 0.0, 1), (3, 0.4, 3)])
def test_sharded_label_equals_unsharded(world, overlap, absolute):
    from oracle import detection_ops as det
    mask, fwd, bwd = _label_case()
    want = det.flow_label(mask, fwd, bwd, overlap=overlap, absolute_overlap=absolute)
    assert want.max() >= 3 and len(np.unique(want[0])) > 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_label_worker, args=(r, world, port, q, overlap, absolute)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t0, t1, lab in got:
        assert np.array_equal(lab, want[t0:t1]), rank


def _growth_case():
    import make_golden as mg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "growth_multi.npz"))
    wvd = mg.growth_multi_case()[:, 10:110, 10:150].copy()
    fwd = (g["fwd_q256"].astype(np.float32) / 256)[:, 10:110, 10:150].copy()
    bwd = (g["bwd_q256"].astype(np.float32) / 256)[:, 10:110, 10:150].copy()
    dt = np.full(wvd.shape[0], 5.0)
    dt[5] = 5.5
    return wvd, fwd, bwd, dt


def _growth_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wvd, fwd, bwd, dt = _growth_case()
        t0, t1 = D.shard_bounds(wvd.shape[0], world, rank)
        fl = D.ShardedFlow(torch.from_numpy(fwd[t0:t1].copy()), torch.from_numpy(bwd[t0:t1].copy()), rank, world,
                           ops=OracleOps())
        shard = D.make_shard(torch.from_numpy(wvd[t0:t1].copy()), rank, world)
        smoothed, markers = fl.detect_growth_markers(shard, dt[t0:t1], t0)
        q.put((rank, t0, t1, smoothed.numpy().copy(), markers.numpy().copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_growth_markers_equal_unsharded(world):
    from oracle import detection_ops as det
    wvd, fwd, bwd, dt = _growth_case()
    want = det.detect_growth_markers(wvd, dt, fwd, bwd, intermediates=True)
    assert want["markers"].max() >= 2 and want["linked"].max() > want["markers"].max()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_growth_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t0, t1, smoothed, markers in got:
        assert np.array_equal(smoothed, want["smoothed"][t0:t1], equal_nan=True), rank
        assert np.array_equal(markers, want["markers"][t0:t1]), rank


def _data():
    bt = synthetic.bt_sequence(T, H, W, seed=77, nans=False)
    bt[3, 5:8, 10:20] = np.nan
    return bt


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bt = _data()
        t0, t1 = D.shard_bounds(T, world, rank)
        shard = D.make_shard(torch.from_numpy(bt[t0:t1].copy()), rank, world)
        fl = D.create_flow_sharded(shard, ops=OracleOps())
        s = np.zeros((3, 3, 3))
        s[:, 1, 1] = 1
        d = fl.convolve(shard, s, reducer=1)
        q.put((rank, t0, t1, fl.fwd.numpy().copy(), fl.bwd.numpy().copy(), d.numpy().copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds():
    for Tn, world in [(288, 8), (7, 2), (7, 3), (10, 4), (3, 3)]:
        b = [D.shard_bounds(Tn, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == Tn
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(world):
    bt = _data()
    ref_f, ref_b = ops.create_flow(bt, backend="numpy")
    ref_d = ops.diff(bt, ref_f, ref_b)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t0, t1, f, b, d in got:
        assert np.array_equal(f, ref_f[t0:t1]), rank
        assert np.array_equal(b, ref_b[t0:t1]), rank
        assert np.array_equal(d, ref_d[t0:t1], equal_nan=True), rank
