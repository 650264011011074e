"""Times the HBM-bound row-wise kernels at the bench shape (CUDA events, 20 iterations, inputs >> L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import ops
from molclr_b200.graph import GraphPlan
from molclr_b200.synth import make_pair_batch

dev = "cuda:0"
B = int(os.environ.get("BATCH", 4096))
bi, _ = make_pair_batch(B, seed=0)
plan = GraphPlan(bi.to(dev))
N, E, D = plan.N, plan.E, 300
g = torch.Generator().manual_seed(0)
srcs = [torch.randn(N, D, generator=g).to(dev) for _ in range(3)]          # rotate inputs: 3 x 123 MB > L2
coef = torch.stack([torch.rand(D) + 0.5, torch.randn(D), torch.randn(D), torch.rand(D) + 0.5]).to(dev)
B1, B2 = torch.randn(5, D).to(dev), torch.randn(3, D).to(dev)


def timeit(fn, iters=21):
    for k in range(3):
        fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(iters):
        fn(k)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def report(name, us, nbytes):
    print(f"{name:44s} {us:8.1f} us  {nbytes / us / 1e3:8.1f} GB/s  ({nbytes / 1e6:.0f} MB algorithmic)")


idx = 4 * (N + 1) + 5 * E
if os.environ.get("ONLY_TABLES"):
    report("edge_table_grad (fp32, fixed order)", timeit(lambda k: ops.edge_table_grad_raw(plan, srcs[k % 3])), 4 * D * N + 32 * N)
    report("embed_nodes_bwd (fp32, fixed order)", timeit(lambda k: ops.embed_nodes_bwd(plan, srcs[k % 3])), 4 * D * N + 4 * N)
    sys.exit(0)
if os.environ.get("ONLY_AGG"):
    report("aggregate fwd (BN+ReLU, fp32 out, tile kernel)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef, round_out=False)),
           4 * D * N * 2 + idx)
    report("aggregate fwd (layer 0, fp32 out, tile kernel)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, round_out=False)), 4 * D * N * 2 + idx)
    sys.exit(0)
report("aggregate fwd (BN+ReLU fused, hi+lo out)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef, want_lo=True)),
       4 * D * N * 3 + idx)
report("aggregate fwd (BN+ReLU fused, hi out)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef)), 4 * D * N * 2 + idx)
report("aggregate fwd (BN+ReLU, fp32 out, tile kernel)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef, round_out=False)),
       4 * D * N * 2 + idx)
report("aggregate fwd (BN+ReLU, fp32 out, row kernel)",
       timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, bn_coef=coef, round_out=False, use_nbr=False)), 4 * D * N * 2 + idx)
report("aggregate fwd (layer 0, fp32 out, tile kernel)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, round_out=False)), 4 * D * N * 2 + idx)
report("aggregate fwd (layer 0, fp32 out, row kernel)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, round_out=False, use_nbr=False)),
       4 * D * N * 2 + idx)
report("aggregate fwd (layer 0, hi+lo out)", timeit(lambda k: ops.gine_aggregate_fwd(plan, srcs[k % 3], B1, B2, want_lo=True)), 4 * D * N * 3 + idx)
report("aggregate bwd (ReLU/BN stats fused)", timeit(lambda k: ops.gine_aggregate_bwd(plan, srcs[k % 3], z_prev=srcs[(k + 1) % 3], bn_coef=coef)),
       4 * D * N * 3 + 4 * (N + 1) + 4 * E)
report("aggregate bwd (ReLU/BN stats fused, row kernel)", timeit(lambda k: ops.gine_aggregate_bwd(plan, srcs[k % 3], z_prev=srcs[(k + 1) % 3], bn_coef=coef, use_nbr=False)),
       4 * D * N * 3 + 4 * (N + 1) + 4 * E)
report("aggregate bwd (plain, row kernel)", timeit(lambda k: ops.gine_aggregate_bwd(plan, srcs[k % 3], use_nbr=False)), 4 * D * N * 2 + 4 * (N + 1) + 4 * E)
report("aggregate bwd (plain)", timeit(lambda k: ops.gine_aggregate_bwd(plan, srcs[k % 3])), 4 * D * N * 2 + 4 * (N + 1) + 4 * E)
bc = torch.randn(3, D).to(dev)
report("bn_bwd_apply", timeit(lambda k: ops.bn_bwd_apply(srcs[k % 3], bc, gy=srcs[(k + 1) % 3])), 4 * D * N * 3)
report("edge_table_grad (fp32, fixed order)", timeit(lambda k: ops.edge_table_grad_raw(plan, srcs[k % 3])), 4 * D * N + 32 * N)
report("embed_nodes_bwd (fp32, fixed order)", timeit(lambda k: ops.embed_nodes_bwd(plan, srcs[k % 3])), 4 * D * N + 4 * N)
report("pool_fwd", timeit(lambda k: ops.pool_fwd(plan, srcs[k % 3], coef, 0)), 4 * D * (N + plan.G))
