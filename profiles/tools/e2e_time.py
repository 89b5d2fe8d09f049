"""Per-call timing of the numpy (host-buffer) API at CONUS size (scratch tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import tobac_flow_b200 as tfb
from tobac_flow_b200 import synthetic
T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bt = synthetic.bt_sequence(T, 1500, 2500, seed=1234, nans=True, device="cuda")
host = torch.empty(bt.shape, dtype=torch.float32, pin_memory=True); host.copy_(bt); torch.cuda.synchronize()
a = host.numpy()
for rep in range(3):
    t0 = time.perf_counter(); fl = tfb.create_flow(a); t1 = time.perf_counter()
    torch.cuda.synchronize(); t1s = time.perf_counter()
    d = fl.diff(a); t2 = time.perf_counter()
    s = fl.sobel(a); t3 = time.perf_counter()
    c = fl.convolve(a); t4 = time.perf_counter()
    print(f"rep {rep}: create_flow call {1e3*(t1-t0):.1f} ms (+sync {1e3*(t1s-t1):.1f}), diff {1e3*(t2-t1s):.1f}, sobel {1e3*(t3-t2):.1f}, convolve {1e3*(t4-t3):.1f}, total {1e3*(t4-t0):.1f} ms -> {T/(t4-t0):.1f} frames/s")
    print(f"   D2H rates: diff {d.nbytes/1e9/(t2-t1s):.1f} GB/s, sobel {s.nbytes/1e9/(t3-t2):.1f} GB/s, convolve {c.nbytes/1e9/(t4-t3):.1f} GB/s")
    del d, s, c
# raw pinned copy rates
x = torch.empty((7, T, 1500, 2500), dtype=torch.float32, device="cuda")
h = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"raw D2H {x.numel()*4/1e9/(t1-t0):.1f} GB/s")
    t0 = time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"raw H2D {x.numel()*4/1e9/(t1-t0):.1f} GB/s")
t0 = time.perf_counter(); h2 = torch.empty(x.shape, dtype=torch.float32, pin_memory=True); t1 = time.perf_counter()
print(f"fresh pinned alloc of {h2.numel()*4/1e9:.2f} GB: {1e3*(t1-t0):.1f} ms")
