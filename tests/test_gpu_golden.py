"""GPU parity tests against golden vectors produced by the reference's OWN model classes (tests/golden/make_encoder_golden.py:
models/ginet_molclr.py, models/gcn_molclr.py, models/ginet_finetune.py unmodified, MolCLR._step with the reference NTXentLoss).
The CUDA drop-ins are compared with the reference's outputs directly -- no oracle in between."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.util import golden_weights, golden_batch, check_golden_grads, max_rel, tol, SMALL_BATCH_TABLE_TOL, SMALL_BATCH_RTOL_GRAD

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import GINet, GCN, NTXentLoss, pretrain_loss, ginet_finetune, gcn_finetune

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
# compensated (tf32x3) forward ~ fp32, single-pass TF32 backward; the fixtures are SMALL batches (24 pairs, 12-14 graphs): gradient
# tolerance tests/util.py:SMALL_BATCH_RTOL_GRAD (measured worst 8.1e-3 here, 1.29e-2 in the 64-graph test of test_gpu_model.py); the 5e-3 bar is asserted at the BASELINE sizes
RTOL_OUT, RTOL_LOSS, RTOL_GRAD = 2e-5, 1e-4, tol("RTOL_GRAD", SMALL_BATCH_RTOL_GRAD)
ZERO_GRADS = ("mlp.2.bias",)          # a bias in front of a BatchNorm: true gradient 0, both sides hold rounding noise


def _load(model, g):
    model.load_state_dict(golden_weights(model.state_dict(), int(g["weight_seed"])))
    return model.to(DEV).train()


def _zero_grad_skips(model):
    return ZERO_GRADS + tuple(f"gnns.{l}.bias" for l in range(len(model.gnns)))


@pytest.mark.parametrize("name,cls", [("enc_gin_pretrain", "GINet"), ("enc_gcn_pretrain", "GCN")])
def test_pretrain_step_matches_reference_models(name, cls):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = _load((GINet if cls == "GINet" else GCN)(5, 300, 512, 0, "mean"), g)
    bi, bj = golden_batch(g, "i").to(DEV), golden_batch(g, "j").to(DEV)
    ris, zis = m(bi)
    for got, key in ((ris, "h_i"), (zis, "out_i")):
        assert max_rel(got, torch.from_numpy(g[key])) < RTOL_OUT, (key, max_rel(got, torch.from_numpy(g[key])))
    m.zero_grad()
    m.load_state_dict(golden_weights(m.state_dict(), int(g["weight_seed"])))     # fresh running statistics for the step proper
    loss = pretrain_loss(m, NTXentLoss(DEV, int(g["batch_size"]), 0.1, True), bi, bj)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < RTOL_LOSS * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    bad = check_golden_grads(m, g, RTOL_GRAD, skip=_zero_grad_skips(m), overrides=SMALL_BATCH_TABLE_TOL)      # 24 pairs
    assert not bad, bad
    for l in (0, 4):     # two forwards (view i, view j) = two momentum updates, as in the reference
        assert max_rel(m.batch_norms[l].running_mean, torch.from_numpy(g[f"running_mean.{l}"])) < 10 * RTOL_OUT
        assert max_rel(m.batch_norms[l].running_var, torch.from_numpy(g[f"running_var.{l}"])) < 10 * RTOL_OUT
    assert int(m.batch_norms[0].num_batches_tracked) == int(g["num_batches_tracked"])
    m.eval()
    with torch.no_grad():
        he, oe = m(bi)
    # eval mode normalises with the running statistics, which carry the (10 x RTOL_OUT) error of the two training forwards
    assert max_rel(he, torch.from_numpy(g["h_i_eval"])) < 5 * RTOL_OUT and max_rel(oe, torch.from_numpy(g["out_i_eval"])) < 5 * RTOL_OUT


SMALL = sorted(glob.glob(os.path.join(GOLDEN, "enc_g*_small_*.npz")))


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[:-4] for p in SMALL])
def test_small_models_all_pools_match_reference_models(path):
    g = np.load(path)
    cls = GCN if "gcn" in os.path.basename(path) else GINet
    m = _load(cls(int(g["layers"]), int(g["emb"]), int(g["feat"]), 0, str(g["pool"])), g)
    h, o = m(golden_batch(g, "b").to(DEV))
    (o.square().sum() + 0.5 * h.sum()).backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < RTOL_OUT and max_rel(o, torch.from_numpy(g["out"])) < RTOL_OUT
    bad = check_golden_grads(m, g, RTOL_GRAD, skip=_zero_grad_skips(m))
    assert not bad, bad


@pytest.mark.parametrize("task,gcn", [("cls", False), ("reg", False), ("cls", True), ("reg", True)])
def test_finetune_matches_reference_model(task, gcn):
    g = np.load(os.path.join(GOLDEN, f"enc_{'gcn_' if gcn else ''}finetune_{task}.npz"))
    m = _load(gcn_finetune.GCN(str(g["task"]), 5, 300, 256, 0, "mean") if gcn else ginet_finetune.GINet(str(g["task"]), 5, 300, 512, 0, "mean"), g)
    h, pred = m(golden_batch(g, "b").to(DEV))
    y = torch.from_numpy(g["y"]).to(DEV)
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten()) if task == "cls" else torch.nn.MSELoss()(pred, y)
    loss.backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < RTOL_OUT and max_rel(pred, torch.from_numpy(g["pred"])) < RTOL_OUT
    assert abs(loss.item() - float(g["loss"])) < RTOL_LOSS * abs(float(g["loss"]))
    bad = check_golden_grads(m, g, RTOL_GRAD, skip=_zero_grad_skips(m))
    assert not bad, bad
