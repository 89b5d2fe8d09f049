"""Timing of the production-settings calls (scripts/dcc_detect_goes.py) on resident CONUS frames (scratch tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import tobac_flow_b200 as tfb
from tobac_flow_b200 import synthetic
from tobac_flow_b200.detection import growth_rate_device
T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bt = synthetic.bt_sequence(T, 1500, 2500, seed=1234, nans=True, device="cuda")
def timed(name, f, reps=2):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:58s} {ms:9.2f} ms  {ms/T:7.3f} ms/frame  {T/ms*1e3:8.1f} frames/s")
    return r
fl0 = timed("create_flow (defaults)", lambda: tfb.create_flow(bt))
fl = timed("create_flow(vr_steps=1, smoothing_passes=1, cubic)", lambda: tfb.create_flow(bt, vr_steps=1, smoothing_passes=1, interp_method="cubic"))
timed("sobel(linear)", lambda: fl.sobel(bt))
timed("sobel(uphill, linear)", lambda: fl.sobel(bt, direction="uphill", method="linear"))
timed("sobel(uphill, cubic)", lambda: fl.sobel(bt, direction="uphill", method="cubic"))
timed("sobel(nearest)", lambda: fl.sobel(bt, method="nearest"))
timed("diff(cubic)", lambda: fl.diff(bt, method="cubic"))
timed("get_growth_rate(cubic)", lambda: growth_rate_device(fl, bt, np.full(T, 5.0), "cubic"))
timed("get_growth_rate(linear)", lambda: growth_rate_device(fl, bt, np.full(T, 5.0), "linear"))
