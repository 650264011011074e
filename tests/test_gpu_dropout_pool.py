"""Dropout (ginet_molclr.py:108-111) and max / add pooling (ginet_molclr.py:83-88) on the fused kernels.

Dropout cannot be bit-compared with torch's RNG: the kernels use a counter-based mask that is a pure function
of (seed, node, feature).  The tests (1) check the mask statistically, and (2) feed the SAME mask (molclr_dropout_mask) to the
oracle explicitly and compare outputs and every parameter gradient value by value."""
import pytest
import torch

from tests.util import rel_err, max_rel, sync_oracle_from, tol, SMALL_BATCH_RTOL_GRAD

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import GCN, GINet, ginet_finetune, ops
    from molclr_b200.synth import make_pair_batch
    from oracle import gnn as ognn

DEV = "cuda:0"
RTOL_OUT, RTOL_GRAD = 5e-5, tol("RTOL_GRAD", SMALL_BATCH_RTOL_GRAD)


def test_dropout_mask_statistics():
    p, N, D = 0.3, 4000, 300
    m = ops.dropout_mask(1234, p, N, D, DEV)
    vals = torch.unique(m)
    assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1 / (1 - p)) < 1e-6
    keep = (m > 0).float()
    assert abs(float(keep.mean()) - (1 - p)) < 3e-3                       # 1.2 M draws: sigma = 4e-4
    assert float((keep.mean(0) - (1 - p)).abs().max()) < 0.04 and float((keep.mean(1) - (1 - p)).abs().max()) < 0.12
    m2 = ops.dropout_mask(1235, p, N, D, DEV)
    agree = ((m > 0) == (m2 > 0)).float().mean()
    assert abs(float(agree) - (p * p + (1 - p) ** 2)) < 5e-3              # different seeds: independent masks
    assert torch.equal(ops.dropout_mask(1234, p, N, D, DEV), m)            # same seed: same mask
    assert torch.equal(ops.dropout_mask(7, 0.0, 10, D, DEV), torch.ones(10, D, device=DEV))


def _run_pair(m, o, b, wh, wo):
    h, out = m(b.to(DEV))
    ((h * wh.to(DEV)).sum() + (out * wo.to(DEV)).sum()).backward()
    ho, oo = o(b)
    ((ho * wh).sum() + (oo * wo).sum()).backward()
    return h, out, ho, oo


@pytest.mark.parametrize("kind,pool", [("gin", "mean"), ("gin", "max"), ("gcn", "add"), ("finetune", "mean")])
def test_dropout_training_matches_oracle_with_the_same_masks(kind, pool):
    torch.manual_seed(0)
    p = 0.3
    if kind == "gin":
        m, o = GINet(5, 300, 512, p, pool).to(DEV), ognn.GINet(5, 300, 512, p, pool)
    elif kind == "gcn":
        m, o = GCN(5, 300, 512, p, pool).to(DEV), ognn.GCN(5, 300, 512, p, pool)
    else:
        m, o = ginet_finetune.GINet("regression", 5, 300, 512, p, pool).to(DEV), ognn.GINetFinetune("regression", 5, 300, 512, p, pool)
    with torch.no_grad():
        for bn in m.batch_norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    sync_oracle_from(m, o)
    b, _ = make_pair_batch(48, seed=13)
    N = b.x.size(0)
    torch.manual_seed(77)
    seeds = m._dropout_seeds()                       # what the next forward will draw ...
    torch.manual_seed(77)                            # ... so rewind the generator for it
    assert all(pp == p for _, pp in seeds) and len({s for s, _ in seeds}) == 5
    o.dropout_masks = [ops.dropout_mask(s, pp, N, 300, DEV).cpu() for s, pp in seeds]
    torch.manual_seed(5) if False else None
    g = torch.Generator().manual_seed(5)
    out_dim = 1 if kind == "finetune" else 256
    wh, wo = torch.randn(48, 512, generator=g), torch.randn(48, out_dim, generator=g)
    h, out, ho, oo = _run_pair(m, o, b, wh, wo)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT, (max_rel(h, ho), max_rel(out, oo))
    bad = []
    for (k, pm), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        if k.endswith("mlp.2.bias") or (kind == "gcn" and k.startswith("gnns") and k.endswith(".bias")):
            continue                                  # in front of a BatchNorm: true gradient 0
        e = rel_err(pm.grad, q.grad)
        if not e < RTOL_GRAD:
            bad.append((k, e))
    assert not bad, bad
    # eval mode: dropout is the identity
    m.eval(); o.eval()
    with torch.no_grad():
        h, out = m(b.to(DEV))
        ho, oo = o(b)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT


@pytest.mark.parametrize("pool", ["max", "add"])
def test_pooling_variants_without_dropout(pool):
    torch.manual_seed(1)
    m, o = GINet(3, 300, 512, 0, pool).to(DEV), ognn.GINet(3, 300, 512, 0, pool)
    sync_oracle_from(m, o)
    b, _ = make_pair_batch(40, seed=3)
    g = torch.Generator().manual_seed(2)
    wh, wo = torch.randn(40, 512, generator=g), torch.randn(40, 256, generator=g)
    h, out, ho, oo = _run_pair(m, o, b, wh, wo)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT
    for (k, pm), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        if not k.endswith("mlp.2.bias"):
            assert rel_err(pm.grad, q.grad) < RTOL_GRAD, k
