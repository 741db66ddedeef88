import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from g_adaptivity_b200 import GNN, synth
from g_adaptivity_b200.trainer import DeformerTrainer
from oracle import gnn_oracle
md=(30,30)
opt=synth.default_opt(md); ds=synth.SyntheticDataset(2,md); data=synth.make_batch(md,32,seed=8)
torch.manual_seed(42)
ref=gnn_oracle.GNNRef(ds,copy.deepcopy(opt))
for graph in (False, True):
  for nofused in (False, True):
    for rep in range(2):
        o=copy.deepcopy(opt); o["device"]="cuda"; o["gad_no_fused_train"]=nofused
        model=GNN(ds,o).to("cuda"); model.load_state_dict(ref.state_dict())
        tr=DeformerTrainer(model, lr=1e-2, use_cuda_graph=graph)
        sid=tr.add_batch(data)
        losses=[]
        for _ in range(6):
            loss=tr.step(sid)
            with torch.cuda.stream(tr.stream):
                losses.append(loss.clone())
        tr.synchronize()
        print("graph",graph,"nofused",nofused,[round(float(x.item()),9) for x in losses], float(tr.flat.sum()))
