"""GPU parity tests of the GCN drop-in (gcn_molclr.py) and its kernels against the CPU oracle."""
import json
import os

import pytest
import torch

from tests.util import rel_err, max_rel, sync_oracle_from, tol, SMALL_BATCH_RTOL_GRAD

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import GCN, NTXentLoss, ops, pretrain_loss
    from molclr_b200.graph import GraphPlan
    from molclr_b200.synth import make_pair_batch
    from oracle import gnn as ognn
    from oracle.nt_xent import NTXentRestated
    from oracle.step import pretrain_loss as oracle_pretrain_loss

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL_OUT, RTOL_LOSS, RTOL_GRAD = 2e-5, 1e-4, tol("RTOL_GRAD", SMALL_BATCH_RTOL_GRAD)      # same policy as tests/test_gpu_model.py (tf32x3 forward)


def _models(seed=0, emb=300, feat=512, precision=None):
    torch.manual_seed(seed)
    m = GCN(5, emb, feat, 0, "mean").to(DEV)
    if precision is not None:              # (None: the class default, "fp16x3")
        m.precision = precision
    with torch.no_grad():
        for bn in m.batch_norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
        for g in m.gnns:
            g.bias.uniform_(-0.2, 0.2)
    return m, sync_oracle_from(m, ognn.GCN(5, emb, feat, 0, "mean"))


def test_gcn_aggregate_bit_exact_and_helpers():
    """(A+I) y + s + b in the reference's fp32 order; tile statistics; row sums; the stand-alone ReLU/BN backward stage."""
    bi, _ = make_pair_batch(40, seed=3)
    plan = GraphPlan(bi.to(DEV))
    N, D = plan.N, 300
    g = torch.Generator().manual_seed(1)
    y = torch.randn(N, D, generator=g)
    conv = ognn.GCNConv(D)
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    ei = ognn.add_self_loops(bi.edge_index, N)
    ea = ognn._self_loop_attr(bi.edge_attr, N)
    want = ognn.propagate_add(ei, y, conv.edge_embedding1(ea[:, 0]) + conv.edge_embedding2(ea[:, 1])) + conv.bias
    got = ops.gcn_aggregate_fwd(plan, y.to(DEV), conv.edge_embedding1.weight.detach().to(DEV), conv.edge_embedding2.weight.detach().to(DEV),
                                conv.bias.detach().to(DEV))
    assert torch.equal(got.cpu(), want.detach())
    # tile statistics feed the same finalize as the GEMM epilogue's
    st, T = ops.bn_tile_stats(got)
    for t in (0, 1, T - 1 if (T - 1) * 32 < N else (N - 1) // 32):
        blk = want.detach()[t * 32:(t + 1) * 32].double()
        if blk.shape[0] == 0:
            continue
        assert rel_err(st[t, 0], blk.mean(0)) < 1e-5
        assert rel_err(st[t, 1], ((blk - blk.mean(0)) ** 2).sum(0)) < 1e-4
    x = torch.randn(8, D, generator=g)
    assert rel_err(ops.row_sum(x.to(DEV)), x.double().sum(1)) < 1e-6
    # hi/lo operand materialisation
    coef = torch.stack([torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g), torch.zeros(D), torch.ones(D)]).to(DEV)
    hi, lo = ops.bn_apply_fwd(got, coef, True, True)
    xr = torch.relu(want.detach() * coef[0].cpu() + coef[1].cpu())
    assert max_rel(hi + lo, xr) < 1e-6
    # ReLU backward + BatchNorm partial sums without a gather
    gg = torch.randn(N, D, generator=g).to(DEV)
    coef4 = torch.stack([coef[0], coef[1], torch.randn(D, generator=g).to(DEV), (torch.rand(D, generator=g) + 0.5).to(DEV)])
    gy, partials, P = ops.relu_bn_bwd_stats(gg, got, coef4, relu=True)
    mask = (got * coef4[0] + coef4[1]) > 0
    assert torch.equal(gy, gg * mask)
    s = partials[:P].double().sum(0)
    xhat = (got.double() - coef4[2].double()) * coef4[3].double()
    assert rel_err(s[0], (gg * mask).double().sum(0)) < 1e-5 and rel_err(s[1], ((gg * mask).double() * xhat).sum(0)) < 1e-5


def test_gcn_state_dict_loads_reference_checkpoint_layout():
    man = json.load(open(os.path.join(GOLDEN, "gcn_ckpt_manifest.json")))["entries"]
    sd = GCN(5, 300, 512, 0, "mean").state_dict()
    assert set(sd.keys()) == set(man.keys())
    for k, v in sd.items():
        assert list(v.shape) == man[k]["shape"] and str(v.dtype).replace("torch.", "") == man[k]["dtype"], k


@pytest.mark.parametrize("bs,precision", [(3, None), (64, None), (64, "tf32x3")])
def test_gcn_forward_train_and_eval(bs, precision):
    m, o = _models(precision=precision)
    bi, _ = make_pair_batch(bs, seed=11)
    h, out = m(bi.to(DEV))
    ho, oo = o(bi)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT, (max_rel(h, ho), max_rel(out, oo))
    for l in range(5):
        assert int(m.batch_norms[l].num_batches_tracked) == 1
        assert max_rel(m.batch_norms[l].running_var, o.batch_norms[l].running_var) < 10 * RTOL_OUT
    m.eval(); o.eval()
    with torch.no_grad():
        h, out = m(bi.to(DEV))
        ho, oo = o(bi)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT


@pytest.mark.parametrize("precision", [None, "tf32x3"])
def test_gcn_backward_all_parameter_gradients(precision):
    m, o = _models(precision=precision)
    bi, _ = make_pair_batch(64, seed=12)
    torch.manual_seed(5)
    wh, wo = torch.randn(64, 512), torch.randn(64, 256)
    h, out = m(bi.to(DEV))
    ((h * wh.to(DEV)).sum() + (out * wo.to(DEV)).sum()).backward()
    ho, oo = o(bi)
    ((ho * wh).sum() + (oo * wo).sum()).backward()
    bad = []
    for (k, p), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, k
        if k.startswith("gnns") and k.endswith(".bias"):
            # a per-feature constant in front of a BatchNorm: the true gradient is 0 and both sides hold rounding noise
            assert float(p.grad.abs().max()) < 1e-3 * float(m.gnns[0].weight.grad.abs().max() + 1e-6), k
            continue
        e = rel_err(p.grad, q.grad)
        if not e < RTOL_GRAD:
            bad.append((k, e))
    assert not bad, bad


def test_gcn_pretrain_step_matches_oracle():
    bs = 96
    m, o = _models(seed=3)
    bi, bj = make_pair_batch(bs, seed=21)
    loss = pretrain_loss(m, NTXentLoss(DEV, bs, 0.1, True), bi.to(DEV), bj.to(DEV))
    loss.backward()
    lo = oracle_pretrain_loss(o, NTXentRestated("cpu", bs, 0.1, True), bi, bj)
    lo.backward()
    assert abs(loss.item() - lo.item()) < RTOL_LOSS * abs(lo.item()), (loss.item(), lo.item())
    for k in ("gnns.0.weight", "gnns.4.weight", "x_embedding1.weight", "batch_norms.2.weight", "feat_lin.weight"):
        assert rel_err(m.get_parameter(k).grad, o.get_parameter(k).grad) < RTOL_GRAD, k


def test_gcn_constructor_errors():
    with pytest.raises(ValueError):
        GCN(1, 32, 16)
    with pytest.raises(ValueError):
        GCN(3, 32, 16, 0, "bogus")
