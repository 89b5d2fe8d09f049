"""Multi-rank NCCL checks (run under torchrun by tests/test_gpu_multi.py; one process per GPU).

Each check computes the unsharded result on every rank (the inputs are small and synthetic) and asserts that the
time-sharded operators, with their NCCL point-to-point halos / gathered tables, reproduce this rank's slice bit for bit:

  flow      create_flow_sharded (+ diff / sobel / convolve on the sharded operand) == create_flow on the whole series
  t_equals  the degenerate split T == world (one frame per rank: end rules applied on one-frame shards)
  label     ShardedFlow.label == Flow.label (label numbers included)
  detect    ShardedFlow.detect_growth_markers == the single-GPU growth-marker pipeline
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import tobac_flow_b200 as tfb  # noqa: E402
from tobac_flow_b200 import _lib, distributed as D, synthetic  # noqa: E402


def same(a, b):
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all().item())


def check_flow(rank, world, dev, T, H, W):
    bt = synthetic.bt_sequence(T, H, W, seed=1239, nans=True, device=dev)
    ref = tfb.create_flow(bt)
    d_ref, s_ref, c_ref = ref.diff(bt), ref.sobel(bt), ref.convolve(bt)
    t0, t1 = D.shard_bounds(T, world, rank)
    shard = D.make_shard(bt[t0:t1].contiguous(), rank, world)
    fl = D.create_flow_sharded(shard, max_value=20)
    s_diff = np.zeros((3, 3, 3)); s_diff[:, 1, 1] = 1
    s_cross = np.zeros((3, 3, 3)); s_cross[1, 1, :] = s_cross[1, :, 1] = s_cross[:, 1, 1] = 1
    d = fl.convolve(shard, s_diff, reducer=_lib.TF_RED_DIFF, exchange=False)
    s = fl.convolve(shard, np.ones((3, 3, 3)), dtype=None, reducer=_lib.TF_RED_SOBEL, exchange=False)
    c = fl.convolve(shard, s_cross, reducer=_lib.TF_RED_NONE, exchange=False)
    ok = (same(fl.fwd, ref.forward_flow_device[t0:t1]) and same(fl.bwd, ref.backward_flow_device[t0:t1])
          and same(d, d_ref[t0:t1]) and same(s, s_ref[t0:t1]) and same(c, c_ref[:, t0:t1]))
    assert ok, f"rank {rank}: sharded flow / stencils differ from the unsharded run (T={T})"


def check_label(rank, world, dev):
    T, H, W = 12, 300, 500
    bt = synthetic.bt_sequence(T, H, W, seed=1237, nans=True)
    mask = np.nan_to_num(bt, nan=300.0) < 262.0
    flow = tfb.create_flow(bt)
    want = flow.label(torch.from_numpy(mask).to(dev), overlap=0.3, absolute_overlap=2)
    t0, t1 = D.shard_bounds(T, world, rank)
    fl = D.ShardedFlow(flow.forward_flow_device[t0:t1].contiguous(), flow.backward_flow_device[t0:t1].contiguous(), rank, world)
    got = fl.label(torch.from_numpy(mask[t0:t1]).to(dev), overlap=0.3, absolute_overlap=2)
    assert torch.equal(got, want[t0:t1]), f"rank {rank}: sharded labels differ"


def check_detect(rank, world, dev):
    import make_golden as mg
    from tobac_flow_b200.detection import growth_markers_device
    wvd = np.tile(mg.growth_multi_case(), (1, 2, 2))
    T = wvd.shape[0]
    dt = np.full(T, 5.0)
    flow = tfb.create_flow(wvd)
    ref = growth_markers_device(flow, torch.from_numpy(wvd).to(dev), dt)
    t0, t1 = D.shard_bounds(T, world, rank)
    fl = D.ShardedFlow(flow.forward_flow_device[t0:t1].contiguous(), flow.backward_flow_device[t0:t1].contiguous(), rank, world)
    shard = D.make_shard(torch.from_numpy(wvd[t0:t1]).to(dev), rank, world)
    smoothed, markers = fl.detect_growth_markers(shard, dt[t0:t1], t0)
    assert same(smoothed, ref["smoothed"][t0:t1]) and torch.equal(markers, ref["markers"][t0:t1]), \
        f"rank {rank}: sharded growth markers differ"


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    check_flow(rank, world, dev, 12, 300, 500)
    check_flow(rank, world, dev, world, 120, 160)       # T == world: one frame per rank
    check_label(rank, world, dev)
    check_detect(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print("sharded checks passed on", world, "ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
