import time, torch, sys
sys.path.insert(0,'/root/repo')
from esvit_b200 import engine
from oracle import losses as L, step as ST, swin as S
K,ncrops,B=65536,10,2
torch.manual_seed(0)
net=engine.build_network(dict(engine.SWIN_SPECS['swin_tiny_w7']),K,True)
g=torch.Generator().manual_seed(11)
with torch.no_grad():
    for n,p in net.named_parameters():
        if n.endswith(".bias") or (p.dim()==1 and "norm" in n) or "relative_position_bias_table" in n:
            p.add_(torch.randn(p.shape,generator=g)*0.1)
crops=ST.synthetic_crops(B,ncrops-2,seed=1234)
ospec=S.SwinSpec(img_size=224,use_dense_prediction=True,**S.SWIN_T_W7)
def run(autocast):
    sd={k:v.detach().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith('weight_g')) for k,v in net.state_dict().items()}
    with torch.autocast("cpu",dtype=torch.bfloat16,enabled=autocast):
        with torch.no_grad():
            t_ref=S.multicrop_forward(crops[:2],{k:v.detach() for k,v in sd.items()},ospec)
        s_ref=S.multicrop_forward(crops,sd,ospec)
        l=L.ddino_loss(s_ref,t_ref,torch.zeros(1,K),torch.zeros(1,K),ncrops,0.04)
    l.backward()
    return float(l),{k:v.grad for k,v in sd.items() if v.grad is not None}
l0,g0=run(False); l1,g1=run(True)
print('loss fp32',l0,'bf16 autocast',l1)
rows=[]
for k in g0:
    if float(g0[k].norm())>1e-7:
        rows.append((float((g1[k].double()-g0[k].double()).norm()/g0[k].double().norm()),k))
rows.sort(reverse=True)
for r in rows[:25]: print('%.4f %s'%r)
import statistics
print('median',statistics.median([r[0] for r in rows]))
