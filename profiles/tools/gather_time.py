import sys, time, os
import numpy as np, torch
sys.path.insert(0, ".")
import tobac_flow_b200 as tfb
from tobac_flow_b200 import synthetic
T = 24
for nans in (False, True):
    bt = synthetic.bt_sequence(T, 1500, 2500, seed=1234, nans=nans, device="cuda")
    fl = tfb.create_flow(bt)
    for name, fn in (("diff", lambda: fl.diff(bt)), ("sobel", lambda: fl.sobel(bt)), ("conv7", lambda: fl.convolve(bt))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"nans={nans} {name:6s} {e0.elapsed_time(e1)/5:.3f} ms per {T} frames")
