"""Host-side cost of one step: wall time of enqueueing the step (no device sync inside) vs the device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import Batch, GCN, GINet, NTXentLoss, pretrain_loss
from molclr_b200.synth import make_pair_batch

dev = torch.device("cuda:0")
B = int(os.environ.get("BATCH", 4096))
for name, cls in (("gin", GINet), ("gcn", GCN)):
    torch.manual_seed(0)
    model = cls(5, 300, 512).to(dev)
    crit = NTXentLoss(dev, B, 0.1, True)
    opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-5, fused=True)
    bi, bj = (b.to(dev) for b in make_pair_batch(B, seed=0))
    f = lambda b: Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.num_graphs)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = pretrain_loss(model, crit, f(bi), f(bj))
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    cpu = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        cpu.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    print(f"{name}: host enqueue {1e3 * min(cpu):.2f} ms/step (min of 5), steady-state {1e2 * (time.perf_counter() - t0):.2f} ms/step")
