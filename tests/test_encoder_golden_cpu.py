"""CPU tests: the oracle's encoder restatement against golden vectors produced by the reference's OWN model classes
(tests/golden/make_encoder_golden.py runs models/ginet_molclr.py, models/gcn_molclr.py and models/ginet_finetune.py unmodified
on a restated torch-geometric 1.6.3 base).  Same ATen primitives in the same order => tight fp32 tolerances."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gnn as ognn
from oracle.nt_xent import NTXentRestated
from tests.util import golden_weights, golden_batch, check_golden_grads, max_rel

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL, TOL_GRAD = 2e-6, 2e-5          # fp32, identical op order up to reduction details of scatter/index_add


def _load(model, g):
    model.load_state_dict(golden_weights(model.state_dict(), int(g["weight_seed"])))
    return model.train()


@pytest.mark.parametrize("name,cls", [("enc_gin_pretrain", "GINet"), ("enc_gcn_pretrain", "GCN")])
def test_oracle_pretrain_step_matches_reference_models(name, cls):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = _load(getattr(ognn, cls)(5, 300, 512, 0, "mean"), g)
    bi, bj = golden_batch(g, "i"), golden_batch(g, "j")
    ris, zis = m(bi)
    rjs, zjs = m(bj)
    loss = NTXentRestated("cpu", int(g["batch_size"]), 0.1, True)(F.normalize(zis, dim=1), F.normalize(zjs, dim=1))
    loss.backward()
    for got, key in ((ris, "h_i"), (zis, "out_i"), (rjs, "h_j"), (zjs, "out_j")):
        assert max_rel(got, torch.from_numpy(g[key])) < TOL, key
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-6)
    assert not check_golden_grads(m, g, TOL_GRAD)
    for l in (0, 4):
        np.testing.assert_allclose(m.batch_norms[l].running_mean.numpy(), g[f"running_mean.{l}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(m.batch_norms[l].running_var.numpy(), g[f"running_var.{l}"], rtol=1e-5, atol=1e-7)
    assert int(m.batch_norms[0].num_batches_tracked) == int(g["num_batches_tracked"])
    m.eval()
    with torch.no_grad():
        he, oe = m(bi)
    assert max_rel(he, torch.from_numpy(g["h_i_eval"])) < TOL and max_rel(oe, torch.from_numpy(g["out_i_eval"])) < TOL


SMALL = sorted(glob.glob(os.path.join(GOLDEN, "enc_g*_small_*.npz")))


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[:-4] for p in SMALL])
def test_oracle_small_models_all_pools_match_reference_models(path):
    g = np.load(path)
    cls = ognn.GCN if "gcn" in os.path.basename(path) else ognn.GINet
    m = _load(cls(int(g["layers"]), int(g["emb"]), int(g["feat"]), 0, str(g["pool"])), g)
    h, o = m(golden_batch(g, "b"))
    (o.square().sum() + 0.5 * h.sum()).backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < TOL and max_rel(o, torch.from_numpy(g["out"])) < TOL
    assert not check_golden_grads(m, g, TOL_GRAD)


@pytest.mark.parametrize("task,gcn", [("cls", False), ("reg", False), ("cls", True), ("reg", True)])
def test_oracle_finetune_matches_reference_model(task, gcn):
    g = np.load(os.path.join(GOLDEN, f"enc_{'gcn_' if gcn else ''}finetune_{task}.npz"))
    m = _load(ognn.GCNFinetune(str(g["task"]), 5, 300, 256, 0, "mean") if gcn else ognn.GINetFinetune(str(g["task"]), 5, 300, 512, 0, "mean"), g)
    h, pred = m(golden_batch(g, "b"))
    y = torch.from_numpy(g["y"])
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten()) if task == "cls" else torch.nn.MSELoss()(pred, y)
    loss.backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < TOL and max_rel(pred, torch.from_numpy(g["pred"])) < TOL
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-6)
    assert not check_golden_grads(m, g, TOL_GRAD)


@pytest.mark.parametrize("task", ["cls", "reg"])
def test_oracle_motif_model_matches_reference_model(task):
    """models/ginet_finetune_mp.py (motif embedding + GlobalAttention): the oracle of the next widening row,
    pinned before a kernel path exists."""
    g = np.load(os.path.join(GOLDEN, f"enc_motif_{task}.npz"))
    m = _load(ognn.GINetMotif(int(g["num_motifs"]), str(g["task"]), 5, 300, 512, 0, "mean"), g)
    h, pred = m(golden_batch(g, "b"), torch.from_numpy(g["mol_idx"]), torch.from_numpy(g["clique_idx"]))
    y = torch.from_numpy(g["y"])
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten()) if task == "cls" else torch.nn.MSELoss()(pred, y)
    loss.backward()
    assert h.shape[1] == 2 * 512
    assert max_rel(h, torch.from_numpy(g["h"])) < TOL and max_rel(pred, torch.from_numpy(g["pred"])) < TOL
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-6)
    assert not check_golden_grads(m, g, TOL_GRAD)


@pytest.mark.parametrize("task", ["cls", "reg"])
def test_motif_dropin_head_and_layout_match_reference_model(task):
    """The drop-in ``molclr_b200.ginet_finetune_mp.GINet``: same state_dict layout as the reference class (through the oracle, whose
    layout the fixture loader pins), and its motif branch -- plain tensor operations, device-agnostic -- reproduces the reference's
    outputs and head gradients when fed the reference's molecule features ``h[:, :feat_dim]``.  (The encoder in front of it needs the GPU.)"""
    from molclr_b200 import ginet_finetune_mp
    g = np.load(os.path.join(GOLDEN, f"enc_motif_{task}.npz"))
    m = ginet_finetune_mp.GINet(int(g["num_motifs"]), str(g["task"]), 5, 300, 512, 0, "mean")
    o = ognn.GINetMotif(int(g["num_motifs"]), str(g["task"]), 5, 300, 512, 0, "mean")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    _load(m, g)
    feat = torch.from_numpy(g["h"])[:, :512].clone().requires_grad_(True)
    h, pred = m.motif_head(feat, torch.from_numpy(g["mol_idx"]), torch.from_numpy(g["clique_idx"]))
    y = torch.from_numpy(g["y"])
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten()) if task == "cls" else torch.nn.MSELoss()(pred, y)
    loss.backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < TOL and max_rel(pred, torch.from_numpy(g["pred"])) < TOL
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-6)
    head = [k for k, _ in m.named_parameters() if k.startswith(("motif_", "pred_head"))]
    assert len(head) >= 9
    bad = check_golden_grads(m, g, TOL_GRAD, skip=tuple(k for k, _ in m.named_parameters() if k not in head))
    assert not bad, bad
    with pytest.raises(RuntimeError):        # no CPU path for the encoder
        m(golden_batch(g, "b"), torch.from_numpy(g["mol_idx"]), torch.from_numpy(g["clique_idx"]))


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference not mounted")
def test_fixtures_reproduce_from_the_live_reference():
    """Dev container only: re-running the unmodified reference GINet on the restated PyG base reproduces the fixture bit for bit."""
    import subprocess, sys
    code = ("import sys, numpy as np, torch; sys.path.insert(0, %r); sys.path.insert(0, %r); import pyg163_stub; pyg163_stub.install();"
            "sys.path.insert(0, '/root/reference'); from models.ginet_molclr import GINet; from tests.util import golden_weights, golden_batch;"
            "g = np.load(%r); m = GINet(3, 64, 64, 0, 'max'); m.load_state_dict(golden_weights(m.state_dict(), int(g['weight_seed'])));"
            "h, o = m(golden_batch(g, 'b')); assert np.array_equal(o.detach().numpy(), g['out']); print('ok')"
            % (os.path.dirname(os.path.dirname(__file__)), GOLDEN, os.path.join(GOLDEN, "enc_gin_small_max.npz")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
