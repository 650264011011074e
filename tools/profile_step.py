"""One MolCLR pre-training step at the bench configuration, bracketed by cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures exactly the kernels of one steady-state step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import Batch, GCN, GINet, NTXentLoss, pretrain_loss
from molclr_b200.synth import make_pair_batch

B = int(os.environ.get("BATCH", 4096))
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = (GCN if os.environ.get("MODEL") == "gcn" else GINet)(5, 300, 512).to(dev)
model.precision = os.environ.get("PREC", model.precision)
crit = NTXentLoss(dev, B, 0.1, True)
opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-5, fused=True)
bi, bj = (b.to(dev) for b in make_pair_batch(B, seed=0))


def step():
    opt.zero_grad(set_to_none=True)
    f = lambda b: Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.num_graphs)
    loss = pretrain_loss(model, crit, f(bi), f(bj))
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
