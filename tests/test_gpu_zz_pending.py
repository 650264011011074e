"""GPU parity tests of drop-ins that were written after the round's GPU budget was spent (non-strict xfail until each has passed
once).  The file name sorts LAST among the test files on purpose: these are the only tests whose kernel-launch sequence has not
been exercised on a GPU, and a CUDA error raised here must not be able to poison the context of verified tests that follow."""
import os

import numpy as np
import pytest
import torch

from tests.util import golden_weights, golden_batch, check_golden_grads, max_rel

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import ginet_finetune_mp, ginet_finetune_link

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL_OUT, RTOL_LOSS, RTOL_GRAD = 2e-5, 1e-4, 2e-2          # the policy of tests/test_gpu_golden.py
ZERO_GRADS = ("mlp.2.bias",)          # a bias in front of a BatchNorm: true gradient 0, both sides hold rounding noise


def _load(model, g):
    model.load_state_dict(golden_weights(model.state_dict(), int(g["weight_seed"])))
    return model.to(DEV).train()


def _zero_grad_skips(model):
    return ZERO_GRADS + tuple(f"gnns.{l}.bias" for l in range(len(model.gnns)))


@pytest.mark.xfail(strict=False, reason="written when the round's GPU budget was spent: one 5-second GPU run failed before the zero-gradient "
                                        "gate bias was excluded from the gradient comparison; not re-run since (the motif head is pinned on CPU)")
@pytest.mark.parametrize("task", ["cls", "reg"])
def test_motif_model_matches_reference_model(task):
    """models/ginet_finetune_mp.py (motif embedding + GlobalAttention) against the reference class's golden vectors."""
    g = np.load(os.path.join(GOLDEN, f"enc_motif_{task}.npz"))
    m = _load(ginet_finetune_mp.GINet(int(g["num_motifs"]), str(g["task"]), 5, 300, 512, 0, "mean"), g)
    h, pred = m(golden_batch(g, "b").to(DEV), torch.from_numpy(g["mol_idx"]).to(DEV), torch.from_numpy(g["clique_idx"]).to(DEV))
    y = torch.from_numpy(g["y"]).to(DEV)
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten()) if task == "cls" else torch.nn.MSELoss()(pred, y)
    loss.backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < RTOL_OUT and max_rel(pred, torch.from_numpy(g["pred"])) < RTOL_OUT
    assert abs(loss.item() - float(g["loss"])) < RTOL_LOSS * abs(float(g["loss"]))
    # the gate bias shifts every logit of a softmax group alike: its true gradient is 0 and both sides hold rounding noise
    bad = check_golden_grads(m, g, RTOL_GRAD, skip=_zero_grad_skips(m) + ("motif_pool.gate_nn.0.bias",))
    assert not bad, bad


@pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent: never run on a GPU (the label head is pinned on CPU)")
def test_link_model_matches_reference_model():
    """models/ginet_finetune_link.py (label-conditioned head) against the reference class's golden vectors."""
    g = np.load(os.path.join(GOLDEN, "enc_link.npz"))
    m = _load(ginet_finetune_link.GINet("classification", 5, 300, 512, 0, "mean"), g)
    h, pred = m(golden_batch(g, "b").to(DEV), DEV)
    y = torch.from_numpy(g["y"]).to(DEV)
    loss = torch.nn.CrossEntropyLoss()(pred, y.flatten())
    loss.backward()
    assert max_rel(h, torch.from_numpy(g["h"])) < RTOL_OUT and max_rel(pred, torch.from_numpy(g["pred"])) < RTOL_OUT
    assert abs(loss.item() - float(g["loss"])) < RTOL_LOSS * abs(float(g["loss"]))
    bad = check_golden_grads(m, g, RTOL_GRAD, skip=_zero_grad_skips(m))
    assert not bad, bad
