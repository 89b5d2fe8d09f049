"""torchrun check: ShardedFlow.detect_growth_markers over N GPUs (NCCL) equals the single-GPU pipeline (scratch tool).
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scratch/detect_sharded_check.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import tobac_flow_b200 as tfb
from tobac_flow_b200 import distributed as D
from tobac_flow_b200.detection import growth_markers_device
import make_golden as mg

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
wvd = np.tile(mg.growth_multi_case(), (1, 6, 8))            # 14 x 720 x 1280, 48 copies of the multi-core case
T = wvd.shape[0]
dt = np.full(T, 5.0)
flow = tfb.create_flow(wvd)                                  # every rank computes the full flow (reference for the check)
ref = growth_markers_device(flow, torch.from_numpy(wvd).to(dev), dt)
t0, t1 = D.shard_bounds(T, world, rank)
fl = D.ShardedFlow(flow.forward_flow_device[t0:t1].contiguous(), flow.backward_flow_device[t0:t1].contiguous(), rank, world)
shard = D.make_shard(torch.from_numpy(wvd[t0:t1]).to(dev), rank, world)
fl.detect_growth_markers(shard, dt[t0:t1], t0)
torch.cuda.synchronize(); dist.barrier()
s = time.perf_counter()
smoothed, markers = fl.detect_growth_markers(shard, dt[t0:t1], t0)
torch.cuda.synchronize(); dist.barrier()
ms = 1e3 * (time.perf_counter() - s)
ok_s = torch.equal(torch.nan_to_num(smoothed, nan=-777.0), torch.nan_to_num(ref["smoothed"][t0:t1], nan=-777.0))
ok_m = torch.equal(markers, ref["markers"][t0:t1])
print(f"rank {rank}: frames [{t0},{t1}) smoothed identical={ok_s} markers identical={ok_m} "
      f"(linked {int(ref['linked'].max())}, markers {int(ref['markers'].max())}) sharded call {ms:.1f} ms", flush=True)
assert ok_s and ok_m
dist.destroy_process_group()
