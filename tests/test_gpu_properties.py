"""Randomised and full-size property tests: irregular graphs the synthetic molecule generator never
produces (hubs with > 32 in-edges, multi-edges, self edges, isolated nodes, one-node graphs) against the oracle, and
size-independent invariants at the BASELINE batch size (4096 pairs) where the CPU oracle would take minutes."""
import numpy as np
import pytest
import torch

from tests.util import rel_err, max_rel, sync_oracle_from

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import Batch, GINet, ops
    from molclr_b200.graph import GraphPlan
    from molclr_b200.synth import make_pair_batch
    from oracle import gnn as ognn
    from oracle.csr import build_csr

DEV = "cuda:0"


def _random_batch(rng, n_graphs, hub_degree):
    xs, eis, eas, bs, off = [], [], [], [], 0
    for g in range(n_graphs):
        n = int(rng.integers(1, 40)) if g else hub_degree + 1          # graph 0: a star whose centre has `hub_degree` in-edges
        xs.append(np.stack([rng.integers(0, 119, n), rng.integers(0, 3, n)], 1))
        if g == 0:
            src = np.arange(1, n); dst = np.zeros(n - 1, dtype=np.int64)
            e = np.stack([np.concatenate([src, dst]), np.concatenate([dst, src])])
        else:
            m = int(rng.integers(0, 3 * n))                               # random multigraph: duplicates and self edges allowed
            e = rng.integers(0, n, (2, m))
        eis.append(e + off)
        eas.append(np.stack([rng.integers(0, 4, e.shape[1]), rng.integers(0, 3, e.shape[1])], 1))
        bs.append(np.full(n, g))
        off += n
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
    return Batch(t(np.concatenate(xs)), t(np.concatenate(eis, 1)), t(np.concatenate(eas)), t(np.concatenate(bs)), n_graphs)


@pytest.mark.parametrize("seed,hub", [(0, 33), (1, 100), (2, 2500)])
def test_irregular_graphs_match_oracle(seed, hub):
    rng = np.random.default_rng(seed)
    b = _random_batch(rng, 24, hub)
    plan = GraphPlan(b.to(DEV))
    want = build_csr(b.edge_index.numpy(), b.edge_attr.numpy(), b.num_nodes)
    assert np.array_equal(plan.rowptr[:plan.N + 1].cpu().numpy(), want["rowptr"])
    assert np.array_equal(plan.col[:plan.E].cpu().numpy(), want["col"])
    assert np.array_equal(plan.col_t[:plan.E].cpu().numpy(), want["col_t"])
    # aggregation forward bit-exact, backward within fp32 rounding (hub rows exercise the > 32-entry fallback of the prefetch)
    D = 300
    torch.manual_seed(seed)
    conv = ognn.GINEConv(D)
    h = torch.randn(b.num_nodes, D)
    ref = conv.aggregate(h, b.edge_index, b.edge_attr)
    got = ops.gine_aggregate_fwd(plan, h.to(DEV), conv.edge_embedding1.weight.detach().to(DEV), conv.edge_embedding2.weight.detach().to(DEV),
                                 round_out=False)
    assert torch.equal(got.cpu(), ref.detach())
    h64 = h.double().requires_grad_(True)
    ga = torch.randn(b.num_nodes, D)
    conv.double().aggregate(h64, b.edge_index, b.edge_attr).backward(ga.double())
    gy, _, _ = ops.gine_aggregate_bwd(plan, ga.to(DEV))
    assert rel_err(gy, h64.grad) < 1e-6
    # whole model, forward and every gradient
    torch.manual_seed(seed)
    m = GINet(3, 64, 64, 0, "mean").to(DEV)
    o = sync_oracle_from(m, ognn.GINet(3, 64, 64, 0, "mean"))
    hh, out = m(b.to(DEV))
    out.square().sum().backward()
    ho, oo = o(b)
    oo.square().sum().backward()
    assert max_rel(out, oo) < 5e-5
    for (k, p), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        if p.grad is not None and q.grad is not None and not k.endswith("mlp.2.bias"):
            assert rel_err(p.grad, q.grad) < 2e-2, k


def test_full_size_invariants():
    """BASELINE config 2 size (4096 pairs, ~100k nodes per view): properties that need no oracle run."""
    bi, _ = make_pair_batch(4096, seed=123)
    plan = GraphPlan(bi.to(DEV))
    N, E, D = plan.N, plan.E, 300
    rowptr, col = plan.rowptr[:N + 1].long(), plan.col[:E].long()
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == E and bool((rowptr[1:] >= rowptr[:-1]).all())
    deg = torch.bincount(bi.edge_index[1].to(DEV), minlength=N)
    assert torch.equal(rowptr[1:] - rowptr[:-1], deg)                                    # in-degrees
    assert torch.equal(torch.sort(col).values, torch.sort(bi.edge_index[0].to(DEV)).values)   # same multiset of sources
    cnt = plan.cnt[:8 * N].view(N, 8)
    assert torch.equal(cnt[:, :5].sum(1), (deg + 1).float()) and torch.equal(cnt[:, 5:].sum(1), (deg + 1).float())
    gptr = plan.gptr[:plan.G + 1].long()
    assert int(gptr[-1]) == N and torch.equal(gptr[1:] - gptr[:-1], torch.bincount(bi.batch.to(DEV), minlength=plan.G))
    # linearity of the aggregation: agg(x + y) - agg(y) == (A + I) x   (table terms cancel)
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(N, D, generator=g).to(DEV), torch.randn(N, D, generator=g).to(DEV)
    B1, B2 = torch.randn(5, D, generator=g).to(DEV), torch.randn(3, D, generator=g).to(DEV)
    zero1, zero2 = torch.zeros(5, D, device=DEV), torch.zeros(3, D, device=DEV)
    lin = ops.gine_aggregate_fwd(plan, x, zero1, zero2, round_out=False)
    diff = ops.gine_aggregate_fwd(plan, x + y, B1, B2, round_out=False) - ops.gine_aggregate_fwd(plan, y, B1, B2, round_out=False)
    assert max_rel(diff, lin) < 1e-5
    # transpose identity: <A x, w> == <x, A^T w>
    w = torch.randn(N, D, generator=g).to(DEV)
    aw, _, _ = ops.gine_aggregate_bwd(plan, w)
    lhs, rhs = (lin.double() * w.double()).sum(), (x.double() * aw.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-6 * abs(float(lhs))
    # NT-Xent at full size: loss of identical views with orthogonal-ish rows is finite and gradients of zis / zjs are consistent
    z = torch.nn.functional.normalize(torch.randn(4096, 256, generator=g), dim=1).to(DEV)
    from molclr_b200 import NTXentLoss
    a, bb = z.clone().requires_grad_(True), z.clone().requires_grad_(True)
    loss = NTXentLoss(DEV, 4096, 0.1, True)(a, bb)
    loss.backward()
    assert torch.isfinite(loss) and float(loss) < float(np.log(2 * 4096 - 1))            # positives have similarity 1
    assert max_rel(a.grad, bb.grad) < 1e-3                                                # symmetric inputs -> symmetric gradients
