import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, cases
from detect_to_track_b200 import ps_roipool as ps
dev=torch.device("cuda:0"); g=torch.Generator(device="cpu").manual_seed(1)
H,W,K,R,nT,NF=38,63,7,300,31,16
brois=torch.stack([torch.from_numpy(cases.rois_random(R,1237+f)) for f in range(NF)]).to(dev)
sgo=torch.randn(NF,R,nT,K,K,generator=g).to(dev)
for _ in range(2):
    ps.ps_roipool_backward_batched(sgo,brois,H,W)
    ps.ps_roipool_backward(sgo[0],brois[0],H,W)
torch.cuda.synchronize(); print("ok")
