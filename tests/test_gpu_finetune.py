"""GPU parity tests of the fine-tune GINet drop-in (models/ginet_finetune.py; BASELINE config 4 shapes) against the oracle."""
import pytest
import torch

from tests.util import rel_err, max_rel, sync_oracle_from, tol, SMALL_BATCH_RTOL_GRAD

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import ginet_finetune, GINet
    from molclr_b200.synth import make_plain_batch
    from oracle import gnn as ognn

DEV = "cuda:0"
RTOL_OUT, RTOL_GRAD = 5e-5, tol("RTOL_GRAD", SMALL_BATCH_RTOL_GRAD)


def _models(task, act="softplus", n_layer=2, seed=0):
    torch.manual_seed(seed)
    m = ginet_finetune.GINet(task, 5, 300, 512, 0, "mean", n_layer, act).to(DEV)
    with torch.no_grad():
        for bn in m.batch_norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    return m, sync_oracle_from(m, ognn.GINetFinetune(task, 5, 300, 512, 0, "mean", n_layer, act))


@pytest.mark.parametrize("task,act,n_layer,atoms", [("classification", "softplus", 2, 46.0), ("regression", "softplus", 2, 26.0),
                                                     ("classification", "relu", 1, 26.0)])
def test_finetune_forward_backward_matches_oracle(task, act, n_layer, atoms):
    """BBBP-shaped (~46 atoms incl. explicit H, 2 logits, CE) and ESOL-shaped (~26 atoms, 1 output, MSE) batches."""
    m, o = _models(task, act, n_layer)
    G = 96
    b = make_plain_batch(G, seed=5, mean_atoms=atoms, std_atoms=atoms / 3)
    torch.manual_seed(1)
    if task == "classification":
        y = (torch.rand(G) < 0.77).long()
        crit = torch.nn.CrossEntropyLoss()
    else:
        y = torch.randn(G, 1) * 2.1 - 3.05
        crit = torch.nn.MSELoss()
    h, pred = m(b.to(DEV))
    loss = crit(pred, y.to(DEV))
    loss.backward()
    ho, po = o(b)
    lo = crit(po, y)
    lo.backward()
    assert pred.shape == po.shape and max_rel(h, ho) < RTOL_OUT and max_rel(pred, po) < RTOL_OUT, (max_rel(h, ho), max_rel(pred, po))
    assert abs(loss.item() - lo.item()) < 1e-4 * abs(lo.item())
    bad = []
    for (k, p), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, k
        if k.endswith("mlp.2.bias"):            # in front of a BatchNorm: true gradient 0
            continue
        e = rel_err(p.grad, q.grad)
        if not e < RTOL_GRAD:
            bad.append((k, e))
    assert not bad, bad


def test_finetune_loads_pretrained_encoder_and_eval_mode():
    """load_my_state_dict (ginet_finetune.py:149-157): pre-trained encoder weights are taken over, the head is kept."""
    torch.manual_seed(3)
    pre = GINet(5, 300, 512, 0, "mean").to(DEV)
    m, o = _models("regression")
    head_before = m.pred_head[0].weight.detach().clone()
    m.load_my_state_dict(pre.state_dict())
    assert torch.equal(m.gnns[3].mlp[0].weight, pre.gnns[3].mlp[0].weight) and torch.equal(m.pred_head[0].weight, head_before)
    sync_oracle_from(m, o)
    m.eval(); o.eval()
    b = make_plain_batch(17, seed=9)
    with torch.no_grad():
        h, pred = m(b.to(DEV))
        ho, po = o(b)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(pred, po) < RTOL_OUT


def test_finetune_constructor_errors():
    with pytest.raises(ValueError):
        ginet_finetune.GINet("classification", pred_act="tanh")


@pytest.mark.parametrize("task", ["cls", "reg"])
def test_motif_model_matches_reference_model(task):
    """models/ginet_finetune_mp.py (motif embedding + GlobalAttention) against golden vectors produced by the reference class
    (passed twice on the round-1 driver box while still marked xfail; a plain test since round 2)."""
    import os
    import numpy as np
    from molclr_b200 import ginet_finetune_mp
    from tests.util import golden_weights, golden_batch, check_golden_grads
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"enc_motif_{task}.npz"))
    m = ginet_finetune_mp.GINet(int(g["num_motifs"]), str(g["task"]), 5, 300, 512, 0, "mean")
    m.load_state_dict(golden_weights(m.state_dict(), int(g["weight_seed"])))
    m = m.to(DEV).train()
    h, pred = m(golden_batch(g, "b").to(DEV), torch.from_numpy(g["mol_idx"]).to(DEV), torch.from_numpy(g["clique_idx"]).to(DEV))
    y = torch.from_numpy(g["y"]).to(DEV)
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten()) if task == "cls" else torch.nn.MSELoss()(pred, y)
    loss.backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < 2e-5 and max_rel(pred, torch.from_numpy(g["pred"])) < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    # zero-gradient parameters (a bias in front of a BatchNorm; the attention gate bias shifts every logit of a softmax group
    # alike): both sides hold rounding noise only
    bad = check_golden_grads(m, g, RTOL_GRAD, skip=("mlp.2.bias", "motif_pool.gate_nn.0.bias"))
    assert not bad, bad


def test_feat_dim_must_keep_head_widths_tensor_core_aligned():
    """feat_dim // 2 is the width of the hidden head activations: a GEMM operand, so it must be a multiple of 4."""
    with pytest.raises(ValueError):
        ginet_finetune.GINet("classification", 5, 300, 300)
