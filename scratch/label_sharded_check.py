"""torchrun check: ShardedFlow.label over N GPUs (NCCL) equals the single-GPU Flow.label (scratch tool).
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 scratch/label_sharded_check.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import tobac_flow_b200 as tfb
from tobac_flow_b200 import distributed as D, synthetic

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
T, H, W = 16, 1500, 2500
bt = synthetic.bt_sequence(T, H, W, seed=1237, nans=True)
mask = np.nan_to_num(bt, nan=300.0) < 262.0
flow = tfb.create_flow(bt)                      # every rank computes the same full flow (reference for the check)
want = flow.label(torch.from_numpy(mask).to(dev), overlap=0.3, absolute_overlap=2).cpu().numpy()
t0, t1 = D.shard_bounds(T, world, rank)
fl = D.ShardedFlow(flow.forward_flow_device[t0:t1].contiguous(), flow.backward_flow_device[t0:t1].contiguous(), rank, world)
m = torch.from_numpy(mask[t0:t1]).to(dev)
fl.label(m, overlap=0.3, absolute_overlap=2)
torch.cuda.synchronize(); dist.barrier()
s = time.perf_counter()
got = fl.label(m, overlap=0.3, absolute_overlap=2)
torch.cuda.synchronize(); dist.barrier()
ms = 1e3 * (time.perf_counter() - s)
ok = np.array_equal(got.cpu().numpy(), want[t0:t1])
print(f"rank {rank}: frames [{t0},{t1}) identical={ok} labels={int(want.max())} sharded label {ms:.1f} ms", flush=True)
assert ok
dist.destroy_process_group()
