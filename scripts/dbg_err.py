import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import gad_testutil as util
from g_adaptivity_b200 import GNN, synth
from oracle import gnn_oracle
md=(30,30); B=int(sys.argv[1]) if len(sys.argv)>1 else 64
opt=synth.default_opt(md); ds=synth.SyntheticDataset(2,md); data=synth.make_batch(md,B,seed=100)
torch.manual_seed(42)
ref=gnn_oracle.GNNRef(ds,copy.deepcopy(opt))
ro=ref(data); gnn_oracle.mesh_loss(ro,data.x_phys).backward()
rg={n:p.grad for n,p in ref.named_parameters() if p.grad is not None}
g64,floor,scale=util.fp64_grads_and_noise_floor(ds,opt,data,ref,rg)
print("floor",{k:f"{v:.2e}" for k,v in floor.items()})
for no_ell in (False,True):
    o=copy.deepcopy(opt); o["device"]="cuda"; o["gad_no_ell"]=no_ell
    m=GNN(ds,o).to("cuda"); m.load_state_dict(ref.state_dict()); m.train()
    out=m(data); F.l1_loss(out,data.x_phys.cuda()).backward()
    res={}
    for n,g in g64.items():
        if "lin_key.bias" in n: continue
        p=dict(m.named_parameters())[n].grad.double().cpu()
        den=max(g.abs().max().item(),1e-3*scale)
        res[n]=f"{(p-g).abs().max().item()/den:.2e}"
    print("no_ell",no_ell,"fwd vs fp32 oracle",util.rel_err(out,ro),res)
