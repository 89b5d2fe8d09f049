"""Time the device-resident detect_growth_markers pipeline at CONUS size (scratch tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import tobac_flow_b200 as tfb
from tobac_flow_b200 import synthetic, _lib
from tobac_flow_b200.detection import growth_markers_device

T = int(sys.argv[1]) if len(sys.argv) > 1 else 48
H, W = 1500, 2500
bt = synthetic.bt_sequence(T, H, W, seed=1235, nans=True, device="cuda")
wvd = synthetic.wvd_from_bt(bt).float().contiguous()
flow = tfb.create_flow(bt)
dt = np.full(T, 5.0)
torch.cuda.synchronize()
for rep in range(3):
    _lib.profile_reset(); _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    r = growth_markers_device(flow, wvd, dt)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    print(f"rep {rep}: wall {1e3*(t1-t0):.1f} ms, device {e0.elapsed_time(e1):.1f} ms, {T/(t1-t0):.1f} frames/s; "
          f"flat labels {int(r['flat'].max())}, linked {int(r['linked'].max())}, markers {int(r['markers'].max())}")
    for k, v in prof.items():
        if v["launches"]:
            print(f"   {k:24s} {v['ms']:8.2f} ms {v['bytes']/max(v['ms'],1e-9)/1e6:8.0f} GB/s  {v['launches']} launches")
