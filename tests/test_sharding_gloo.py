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
            r = ops.diff_reducer(stack).astype(dtype)
            r[np.isnan(d[j])] = fill_value
            res.append(r)
        return torch.from_numpy(np.stack(res))


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
