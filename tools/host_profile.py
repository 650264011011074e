"""cProfile of the host side of the pre-training step at a small batch (where the host, not the GPU, sets the step time)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import Batch, GINet, NTXentLoss, pretrain_loss
from molclr_b200.synth import make_pair_batch

dev = torch.device("cuda:0")
B = int(os.environ.get("BATCH", 512))
torch.manual_seed(0)
model = GINet(5, 300, 512).to(dev)
crit = NTXentLoss(dev, B, 0.1, True)
opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-5, fused=True)
bi, bj = (b.to(dev) for b in make_pair_batch(B, seed=0))
f = lambda b: Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.num_graphs)


def step():
    opt.zero_grad(set_to_none=True)
    loss = pretrain_loss(model, crit, f(bi), f(bj))
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
