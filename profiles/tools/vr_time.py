import torch, time, numpy as np, sys
sys.path.insert(0,'.')
import tobac_flow_b200 as tfb
from tobac_flow_b200 import synthetic, _lib
dev=torch.device('cuda')
T=27
bt=synthetic.bt_sequence(T,1500,2500,seed=1234,nans=True,device=dev)
for kw in [dict(), dict(vr_steps=1), dict(vr_steps=1,smoothing_passes=1,interp_method='cubic')]:
    f=tfb.create_flow(bt,**kw); torch.cuda.synchronize()
    _lib.profile_reset(); _lib.profile_enable(True)
    t0=time.perf_counter(); f=tfb.create_flow(bt,**kw); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    p=_lib.profile_read(); _lib.profile_enable(False)
    print(kw, 'ms per pair', dt/(T-1)*1e3, {k:round(v['ms']/(T-1),3) for k,v in p.items() if v['launches']})
