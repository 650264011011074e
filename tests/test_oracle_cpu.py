"""CPU tests: the oracle against the reference's golden vectors and against itself."""
import glob
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import gnn as ognn
from oracle.csr import build_csr, build_graph_segments
from oracle.nt_xent import NTXentRestated, ntxent_closed_form
from oracle.step import pretrain_loss
from molclr_b200.synth import make_pair_batch, make_plain_batch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NTX = sorted(glob.glob(os.path.join(GOLDEN, "ntxent_*.npz")))


@pytest.mark.parametrize("path", NTX, ids=[os.path.basename(p)[:-4] for p in NTX])
def test_ntxent_restated_matches_reference_golden(path):
    g = np.load(path)
    zis = torch.tensor(g["zis"], requires_grad=True)
    zjs = torch.tensor(g["zjs"], requires_grad=True)
    crit = NTXentRestated("cpu", int(g["batch_size"]), float(g["temperature"]), bool(g["cosine"]))
    loss = crit(zis, zjs)
    loss.backward()
    # same ATen ops in the same order as the reference class: tight tolerance (fp32)
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-6)
    np.testing.assert_allclose(zis.grad.numpy(), g["dzis"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(zjs.grad.numpy(), g["dzjs"], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("path", NTX, ids=[os.path.basename(p)[:-4] for p in NTX])
def test_ntxent_closed_form_matches_reference_golden(path):
    g = np.load(path)
    zis = torch.tensor(g["zis"], dtype=torch.float64, requires_grad=True)
    zjs = torch.tensor(g["zjs"], dtype=torch.float64, requires_grad=True)
    loss = ntxent_closed_form(zis, zjs, float(g["temperature"]), bool(g["cosine"]), chunk=37)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-6)
    np.testing.assert_allclose(zis.grad.numpy(), g["dzis"], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(zjs.grad.numpy(), g["dzjs"], rtol=2e-4, atol=2e-7)


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils"), reason="reference not mounted")
def test_ntxent_restated_matches_live_reference():
    sys.path.insert(0, "/root/reference")
    from utils.nt_xent import NTXentLoss
    torch.manual_seed(3)
    for n, c, tau, cos in [(16, 32, 0.1, True), (24, 8, 0.7, False)]:
        a, b = torch.randn(n, c), torch.randn(n, c)
        ref = NTXentLoss("cpu", n, tau, cos)(a, b)
        mine = NTXentRestated("cpu", n, tau, cos)(a, b)
        assert torch.equal(ref, mine)


def test_gcn_state_dict_layout_matches_shipped_checkpoint():
    man = json.load(open(os.path.join(GOLDEN, "gcn_ckpt_manifest.json")))["entries"]
    sd = ognn.GCN(5, 300, 512, 0, "mean").state_dict()
    assert set(sd.keys()) == set(man.keys())
    for k, v in sd.items():
        assert list(v.shape) == man[k]["shape"], k
        assert str(v.dtype).replace("torch.", "") == man[k]["dtype"], k


def test_gin_state_dict_layout():
    sd = ognn.GINet(5, 300, 512, 0, "mean").state_dict()
    assert sd["gnns.0.mlp.0.weight"].shape == (600, 300)
    assert sd["gnns.4.mlp.2.weight"].shape == (300, 600)
    assert sd["gnns.2.edge_embedding1.weight"].shape == (5, 300)
    assert sd["x_embedding1.weight"].shape == (119, 300)
    assert sd["out_lin.2.weight"].shape == (256, 512)
    assert sd["batch_norms.3.num_batches_tracked"].dtype == torch.int64
    assert sum(p.numel() for p in ognn.GINet(5, 300, 512).parameters()) == 2404196


def test_gin_aggregate_equals_sequential_csr_sum_bitwise():
    """The ordering contract: scatter_add_ == sequential fp32 sum over the stable dst-sorted
    edge list, message formed as h[src] + (B1[t] + B2[d]), self-loop last."""
    torch.manual_seed(0)
    bi, _ = make_pair_batch(24, seed=5)
    conv = ognn.GINEConv(12)
    h = torch.randn(bi.num_nodes, 12)
    with torch.no_grad():
        ref = conv.aggregate(h, bi.edge_index, bi.edge_attr).numpy()
        csr = build_csr(bi.edge_index.numpy(), bi.edge_attr.numpy(), bi.num_nodes)
        tab = (conv.edge_embedding1.weight[:, None, :] + conv.edge_embedding2.weight[None, :, :]).reshape(15, 12).numpy()
        hn = h.numpy()
        out = np.zeros_like(ref)
        for i in range(bi.num_nodes):
            acc = np.zeros(12, dtype=np.float32)
            for e in range(csr["rowptr"][i], csr["rowptr"][i + 1]):
                acc = acc + (hn[csr["col"][e]] + tab[csr["eattr"][e]])
            acc = acc + (hn[i] + tab[12])          # self loop: type 4, dir 0 -> row 4*3+0
            out[i] = acc
    assert np.array_equal(out, ref)


def test_csr_oracle_roundtrip_and_counts():
    bi, _ = make_pair_batch(16, seed=2)
    ei, ea, n = bi.edge_index.numpy(), bi.edge_attr.numpy(), bi.num_nodes
    c = build_csr(ei, ea, n)
    assert c["rowptr"][-1] == ei.shape[1] and c["rowptr_t"][-1] == ei.shape[1]
    # every (src,dst,attr) triple is present exactly once in the dst-sorted structure
    trip = sorted(zip(ei[0], ei[1], ea[:, 0] * 3 + ea[:, 1]))
    got = sorted((c["col"][e], i, c["eattr"][e]) for i in range(n) for e in range(c["rowptr"][i], c["rowptr"][i + 1]))
    assert trip == [tuple(int(v) for v in t) for t in got]
    assert (c["cnt"][:, :5].sum(1) == np.diff(c["rowptr"]) + 1).all()
    gptr, perm = build_graph_segments(bi.batch.numpy(), bi.num_graphs)
    assert gptr[-1] == n and (perm == np.arange(n)).all()


def test_empty_edge_set_and_single_graph():
    x = torch.tensor([[5, 0], [118, 0], [7, 1]])
    from molclr_b200.batch import Batch
    b = Batch(x, torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 2, dtype=torch.long), torch.zeros(3, dtype=torch.long))
    m = ognn.GINet(2, 16, 8)
    h, out = m(b)
    assert h.shape == (1, 8) and out.shape == (1, 4) and torch.isfinite(out).all()


def test_fp32_vs_fp64_oracle_and_step():
    torch.manual_seed(1)
    bi, bj = make_pair_batch(12, seed=7)
    m32 = ognn.GINet(3, 20, 16)
    m64 = ognn.GINet(3, 20, 16).double()
    m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in m32.state_dict().items()})
    l32 = pretrain_loss(m32, NTXentRestated("cpu", 12, 0.1, True), bi, bj)
    l64 = pretrain_loss(m64, lambda a, b: ntxent_closed_form(a, b, 0.1, True), bi, bj)
    assert abs(l32.item() - l64.item()) < 1e-4 * abs(l64.item())
    # BN updated twice per step (molclr.py:57,60)
    assert int(m32.batch_norms[0].num_batches_tracked) == 2


def test_gcn_oracle_is_unnormalised_sum():
    torch.manual_seed(2)
    b = make_plain_batch(5, seed=3)
    conv = ognn.GCNConv(10)
    conv.bias.data.normal_()
    h = torch.randn(b.num_nodes, 10)
    out = conv(h, b.edge_index, b.edge_attr)
    n = b.num_nodes
    A = torch.zeros(n, n)
    A.index_put_((b.edge_index[1], b.edge_index[0]), torch.ones(b.edge_index.size(1)), accumulate=True)
    s = torch.zeros(n).index_add_(0, b.edge_index[1], (conv.edge_embedding1.weight[b.edge_attr[:, 0], 0]
                                                         + conv.edge_embedding2.weight[b.edge_attr[:, 1], 0]))
    s = s + conv.edge_embedding1.weight[4, 0] + conv.edge_embedding2.weight[0, 0]
    want = (A + torch.eye(n)) @ (h @ conv.weight) + s[:, None] + conv.bias
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-5)
