import torch, sys
sys.path.insert(0,'.')
import tobac_flow_b200 as tfb
from tobac_flow_b200 import synthetic
bt=synthetic.bt_sequence(5,1500,2500,seed=1234,nans=True,device=torch.device('cuda'))
f=tfb.create_flow(bt,vr_steps=1); torch.cuda.synchronize()
print(float(f.forward_flow_device.abs().mean()))
