"""Verbose parity report (GPU): per-parameter gradient errors of GINet vs the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import rel_err, max_rel, sync_oracle_from
from molclr_b200 import GINet, NTXentLoss, pretrain_loss
from molclr_b200.synth import make_pair_batch
from oracle import gnn as ognn
from oracle.nt_xent import NTXentRestated
from oracle.step import pretrain_loss as opl

DEV = "cuda:0"
torch.manual_seed(0)
bs = int(os.environ.get("BS", 64))
m = GINet(5, 300, 512).to(DEV)
m.precision = os.environ.get("PREC", "tf32x3")
with torch.no_grad():
    for bn in m.batch_norms:
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
o = sync_oracle_from(m, ognn.GINet(5, 300, 512))
o64 = ognn.GINet(5, 300, 512).double()
o64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in o.state_dict().items()})
bi, bj = make_pair_batch(bs, seed=12)
wh, wo = torch.randn(bs, 512), torch.randn(bs, 256)
h, out = m(bi.to(DEV)); ((h * wh.to(DEV)).sum() + (out * wo.to(DEV)).sum()).backward()
ho, oo = o(bi); ((ho * wh).sum() + (oo * wo).sum()).backward()
h6, o6 = o64(bi); ((h6 * wh.double()).sum() + (o6 * wo.double()).sum()).backward()
print("impl", os.environ.get("MOLCLR_GEMM_IMPL", "tc"), m.precision, "fwd h", max_rel(h, ho), "out", max_rel(out, oo), "| oracle fp32 vs fp64 h", max_rel(ho, h6))
for (k, p), (_, q), (_, r) in zip(m.named_parameters(), o.named_parameters(), o64.named_parameters()):
    print(f"{k:36s} cuda-vs-fp32 {rel_err(p.grad, q.grad):.3e}  cuda-vs-fp64 {rel_err(p.grad, r.grad):.3e}  fp32-vs-fp64 {rel_err(q.grad, r.grad):.3e}")
