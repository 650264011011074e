"""Launches the GINE aggregation kernels a few times at the bench shape (target of `ncu --kernel-name regex:aggregate`).
    python tools/run_aggregate.py [fwd|fwd0|bwd|row] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import ops
from molclr_b200.graph import GraphPlan
from molclr_b200.synth import make_pair_batch

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = "cuda:0"
bi, _ = make_pair_batch(int(os.environ.get("BATCH", 4096)), seed=0)
plan = GraphPlan(bi.to(dev))
N, D = plan.N, 300
g = torch.Generator().manual_seed(0)
srcs = [torch.randn(N, D, generator=g).to(dev) for _ in range(3)]
coef = torch.stack([torch.rand(D) + 0.5, torch.randn(D), torch.randn(D), torch.rand(D) + 0.5]).to(dev)
B1, B2 = torch.randn(5, D).to(dev), torch.randn(3, D).to(dev)
for k in range(iters):
    if which == "fwd":
        ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef, round_out=False)
    elif which == "fwd0":
        ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, round_out=False)
    elif which == "row":
        ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef, round_out=False, use_nbr=False)
    else:
        ops.gine_aggregate_bwd(plan, srcs[k % 3], z_prev=srcs[(k + 1) % 3], bn_coef=coef)
torch.cuda.synchronize()
print("done", which, iters)
