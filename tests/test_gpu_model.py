"""GPU parity tests of the drop-in classes against the CPU oracle and the reference's golden vectors."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.util import rel_err, max_rel, sync_oracle_from, tol, grad_tolerance, SMALL_BATCH_TABLE_TOL

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import molclr_b200
    from molclr_b200 import GINet, NTXentLoss, normalize, pretrain_loss
    from molclr_b200.synth import make_pair_batch
    from oracle import gnn as ognn
    from oracle.nt_xent import NTXentRestated, ntxent_closed_form
    from oracle.step import pretrain_loss as oracle_pretrain_loss

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NTX = sorted(glob.glob(os.path.join(GOLDEN, "ntxent_*.npz")))

# Tolerances.  Precisions "fp16x3" (default) and "tf32x3": forward contractions are error-compensated three-product forms (~fp32),
# backward contractions single-pass TF32 (operands rounded to 10 mantissa bits, fp32 accumulate); everything
# else is fp32.  Gradients of a ReLU network are only sqrt-continuous in the activations (a pre-activation
# within rounding distance of 0 flips its mask), so even the fp32 CPU reference differs from its own fp64
# run by ~1e-3 in the norm of a gradient (tests/parity_report.py prints that floor); RTOL_GRAD sits above it.
RTOL_OUT = 2e-5        # max-relative error of activations / outputs (compensated forward)
RTOL_LOSS = 1e-4       # relative error of the scalar loss
RTOL_GRAD = tol("RTOL_GRAD", 5e-3)       # norm-relative error of EVERY parameter gradient (measured: <= 3.7e-3 from 128 pairs up, profiles/parity_r2.json)
RTOL_OUT_TF32 = 5e-3   # single-pass "tf32" mode: activations
RTOL_GRAD_TF32 = 1.5e-1  # single-pass "tf32" mode: gradients (ReLU mask flips, see above)


@pytest.mark.parametrize("path", NTX, ids=[os.path.basename(p)[:-4] for p in NTX])
def test_ntxent_matches_reference_golden(path):
    g = np.load(path)
    n = int(g["batch_size"])
    zis = torch.tensor(g["zis"], device=DEV, requires_grad=True)
    zjs = torch.tensor(g["zjs"], device=DEV, requires_grad=True)
    crit = NTXentLoss(DEV, n, float(g["temperature"]), bool(g["cosine"]))
    loss = crit(zis, zjs)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    assert rel_err(zis.grad, torch.tensor(g["dzis"])) < RTOL_GRAD
    assert rel_err(zjs.grad, torch.tensor(g["dzjs"])) < RTOL_GRAD


# (6, 12): candidate count 12 = 4 mod 8 (ragged fp16 chunks); (2500, 256): several candidate splits of the fused backward
@pytest.mark.parametrize("n,c,tau,cos", [(512, 256, 0.1, True), (1500, 256, 0.1, True), (256, 64, 0.5, False), (6, 12, 0.1, True), (2500, 256, 0.1, True),
                                         (2500, 64, 0.5, False),
                                         # C > 256: the striped fp16 backward (W stripes + K-major cols^T) instead of the fused kernel
                                         (300, 320, 0.1, True)])
def test_ntxent_matches_closed_form(n, c, tau, cos):
    torch.manual_seed(n)
    a = torch.randn(n, c)
    b = 0.7 * a + 0.7 * torch.randn(n, c)
    if cos:
        a, b = torch.nn.functional.normalize(a, dim=1), torch.nn.functional.normalize(b, dim=1)
    else:
        a, b = a * 0.2, b * 0.2
    a64, b64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = ntxent_closed_form(a64, b64, tau, cos)
    ref.backward()
    zis, zjs = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    loss = NTXentLoss(DEV, n, tau, cos)(zis, zjs)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-3 * abs(ref.item()), (loss.item(), ref.item())
    assert rel_err(zis.grad, a64.grad) < RTOL_GRAD and rel_err(zjs.grad, b64.grad) < RTOL_GRAD


@pytest.mark.parametrize("R,Rc,C", [(1024, 1024, 256), (512, 6000, 256), (64, 200, 40)])
def test_ntxent_fp16_operands_equal_tf32_operands(R, Rc, C):
    """unit_rows=True (fp16 tensor-core operands, fp16 W stripes) against unit_rows=False (TF32 operands, fp32 W): both carry
    11-bit significands, so losses / log-sum-exps / gradients agree to rounding, local rows inside a larger candidate set."""
    from molclr_b200 import ops
    g = torch.Generator().manual_seed(R + Rc)
    cols = ops.round_tf32(torch.nn.functional.normalize(torch.randn(Rc, C, generator=g), dim=1).to(DEV))
    rep = torch.cat([cols[:R // 2], cols[Rc // 2:Rc // 2 + R // 2]]).contiguous()
    res = []
    for unit in (True, False):
        loss, lse, pos = ops.ntxent_fwd(rep, cols, 0, 10.0, Rc // 2, unit_rows=unit)
        col_lse = torch.full((Rc,), float(lse.mean()), device=DEV)
        col_lse[:R // 2], col_lse[Rc // 2:Rc // 2 + R // 2] = lse[:R // 2], lse[R // 2:]
        res.append((loss, lse, pos, ops.ntxent_bwd(rep, cols, 0, 10.0, lse, col_lse, Rc // 2, unit_rows=unit), col_lse))
    (l1, s1, p1, g1, _), (l0, s0, p0, g0, cl) = res
    assert abs(float(l1) - float(l0)) < 2e-5 * abs(float(l0))
    assert (s1 - s0).abs().max() < 1e-4 and (p1 - p0).abs().max() < 1e-4
    assert rel_err(g1, g0) < 2e-3, rel_err(g1, g0)
    # and against the dense fp64 evaluation of the same formula
    lg = (rep.double() @ cols.double().T) * 10.0
    rows = torch.cat([torch.arange(R // 2), torch.arange(R // 2) + Rc // 2]).to(DEV)
    self_mask = torch.arange(Rc, device=DEV)[None, :] == rows[:, None]
    w = torch.exp(lg - s0.double()[:, None]) + torch.exp(lg - cl.double()[None, :])
    w = w.masked_fill(self_mask, 0.0)
    w[torch.arange(R, device=DEV), (rows + Rc // 2) % Rc] -= 2.0
    ref = (10.0 / Rc) * (w @ cols.double())
    assert rel_err(g1, ref) < 2e-3 and rel_err(g0, ref) < 2e-3, (rel_err(g1, ref), rel_err(g0, ref))
    ref_lse = torch.logsumexp(lg.masked_fill(self_mask, float("-inf")), dim=1)
    assert (s1.double() - ref_lse).abs().max() < 2e-3


def test_ntxent_batch_size_mismatch_raises():
    crit = NTXentLoss(DEV, 8, 0.1, True)
    with pytest.raises(RuntimeError):
        crit(torch.randn(6, 16, device=DEV), torch.randn(6, 16, device=DEV))


def _models(num_layer=5, emb=300, feat=512, seed=0, precision=None):
    torch.manual_seed(seed)
    m = GINet(num_layer, emb, feat, 0, "mean").to(DEV)
    if precision is not None:              # (None: the class default)
        m.precision = precision
    with torch.no_grad():                      # non-trivial BN affine so gamma/beta gradients are exercised
        for bn in m.batch_norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    o = sync_oracle_from(m, ognn.GINet(num_layer, emb, feat, 0, "mean"))
    return m, o


def test_ginet_state_dict_matches_oracle_layout():
    m, o = _models()
    sm, so = m.state_dict(), o.state_dict()
    assert list(sm.keys()) == list(so.keys())
    assert all(sm[k].shape == so[k].shape and sm[k].dtype == so[k].dtype for k in sm)


@pytest.mark.parametrize("bs,precision", [(3, "tf32x3"), (64, "tf32x3"), (3, "fp16x3"), (64, "fp16x3"), (64, "tf32")])
def test_ginet_forward_train_and_eval(bs, precision):
    m, o = _models(precision=precision)
    RTOL_OUT = globals()["RTOL_OUT"] if precision != "tf32" else RTOL_OUT_TF32
    bi, _ = make_pair_batch(bs, seed=11)
    h, out = m(bi.to(DEV))
    ho, oo = o(bi)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT, (max_rel(h, ho), max_rel(out, oo))
    # running statistics updated identically (momentum 0.1, unbiased variance), once per forward
    for l in range(5):
        assert int(m.batch_norms[l].num_batches_tracked) == 1
        assert max_rel(m.batch_norms[l].running_mean, o.batch_norms[l].running_mean) < 10 * RTOL_OUT
        assert max_rel(m.batch_norms[l].running_var, o.batch_norms[l].running_var) < 10 * RTOL_OUT
    m.eval(); o.eval()
    with torch.no_grad():
        h, out = m(bi.to(DEV))
        ho, oo = o(bi)
    assert max_rel(h, ho) < RTOL_OUT and max_rel(out, oo) < RTOL_OUT
    assert int(m.batch_norms[0].num_batches_tracked) == 1


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3", "tf32"])
def test_ginet_backward_all_parameter_gradients(precision):
    m, o = _models(precision=precision)
    RTOL_GRAD = globals()["RTOL_GRAD"] if precision != "tf32" else RTOL_GRAD_TF32
    bi, _ = make_pair_batch(64, seed=12)
    torch.manual_seed(5)
    wh, wo = torch.randn(64, 512), torch.randn(64, 256)
    h, out = m(bi.to(DEV))
    ((h * wh.to(DEV)).sum() + (out * wo.to(DEV)).sum()).backward()
    ho, oo = o(bi)
    ((ho * wh).sum() + (oo * wo).sum()).backward()
    bad = []
    for (k, p), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        assert p.grad is not None, k
        if k.endswith("mlp.2.bias"):           # bias in front of a BatchNorm: true gradient is 0, both sides hold rounding noise
            assert float(p.grad.abs().max()) < 1e-3 * float(m.get_parameter(k.replace("mlp.2.bias", "mlp.0.bias")).grad.abs().max() + 1e-6)
            continue
        e = rel_err(p.grad, q.grad)
        if not e < (RTOL_GRAD if precision == "tf32" else grad_tolerance(k, RTOL_GRAD, SMALL_BATCH_TABLE_TOL)):       # 64 graphs
            bad.append((k, e))
    assert not bad, bad


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
def test_pretrain_step_loss_and_gradients_config1_shape(precision):
    """BASELINE config 1 scaled to 128 pairs: MolCLR._step with the unmodified loss formulation."""
    bs = 128
    m, o = _models(seed=3, precision=precision)
    bi, bj = make_pair_batch(bs, seed=21)
    loss = pretrain_loss(m, NTXentLoss(DEV, bs, 0.1, True), bi.to(DEV), bj.to(DEV))
    loss.backward()
    lo = oracle_pretrain_loss(o, NTXentRestated("cpu", bs, 0.1, True), bi, bj)
    lo.backward()
    assert abs(loss.item() - lo.item()) < RTOL_LOSS * abs(lo.item()), (loss.item(), lo.item())
    assert int(m.batch_norms[0].num_batches_tracked) == 2        # two encoder passes per step (molclr.py:57,60)
    bad = []
    for (k, p), (_, q) in zip(m.named_parameters(), o.named_parameters()):
        if k.endswith("mlp.2.bias"):
            continue
        e = rel_err(p.grad, q.grad)
        if not e < RTOL_GRAD:
            bad.append((k, e))
    assert not bad, bad


def test_fp16x3_range_violation_is_reported_by_the_next_forward():
    """An activation beyond fp16's finite range under precision 'fp16x3' is clamped on the device, recorded in a sticky status word
    and raised by the NEXT forward (no stream drain in between); 'tf32x3' computes the same input without complaint."""
    torch.manual_seed(0)
    m = GINet(2, 32, 16).to(DEV)
    m.precision = "fp16x3"
    m.fp16_check_every = 1
    bi, _ = make_pair_batch(4, seed=1)
    h, _ = m(bi.to(DEV))                                    # in range: nothing is flagged
    torch.cuda.synchronize()
    m(bi.to(DEV))
    with torch.no_grad():
        m.x_embedding1.weight.fill_(1.0e5)
    m(bi.to(DEV))
    torch.cuda.synchronize()
    m(bi.to(DEV))                                           # (its own copy of the status word is taken here ...)
    torch.cuda.synchronize()
    with pytest.raises(FloatingPointError):
        m(bi.to(DEV))                                       # ... and examined here at the latest
    m(bi.to(DEV))                                           # still out of range: the blocking check finds it at once
    with pytest.raises(FloatingPointError):
        m.check_fp16_range()
    m.precision = "tf32x3"
    h, _ = m(bi.to(DEV))
    assert bool(torch.isfinite(h).all())
    m.check_fp16_range()                                    # (nothing pending)


def test_unknown_pool_and_cpu_input_fail_loudly():
    m = GINet(2, 32, 16, 0, "bogus").to(DEV)
    bi, _ = make_pair_batch(2, seed=1)
    with pytest.raises(AttributeError):
        m(bi.to(DEV))
    m = GINet(2, 32, 16).to(DEV)
    with pytest.raises(RuntimeError):
        m(bi)                     # CPU tensors: there is no CPU path


def test_ntxent_global_negatives_two_rank_emulation():
    """The multi-GPU NT-Xent kernels (rows = this rank's [zjs; zis] halves inside the all-gathered candidates) run rank by
    rank on ONE GPU against the fp64 closed form over the concatenated batch (dist.py drives the same calls)."""
    from molclr_b200 import ops
    W, B, C, tau = 2, 384, 256, 0.1
    torch.manual_seed(7)
    a = torch.nn.functional.normalize(torch.randn(W * B, C), dim=1)
    b = torch.nn.functional.normalize(0.7 * a + 0.7 * torch.randn(W * B, C), dim=1)
    zis64, zjs64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = ntxent_closed_form(zis64, zjs64, tau, True)
    ref.backward()
    Bg = W * B
    cols = ops.round_tf32(torch.cat([b, a]).to(DEV))                       # [all zjs; all zis]
    shares, lses = [], []
    for r in range(W):
        local = torch.cat([cols[r * B:(r + 1) * B], cols[Bg + r * B:Bg + (r + 1) * B]]).contiguous()
        loss, lse, _ = ops.ntxent_fwd(local, cols, r * B, 1.0 / tau, Bg + r * B)
        shares.append(loss)
        lses.append(lse)
    total = sum(float(s) for s in shares)
    assert abs(total - float(ref)) < RTOL_LOSS * abs(float(ref)), (total, float(ref))
    col_lse = torch.cat([torch.cat([l[:B] for l in lses]), torch.cat([l[B:] for l in lses])])
    for r in range(W):
        local = torch.cat([cols[r * B:(r + 1) * B], cols[Bg + r * B:Bg + (r + 1) * B]]).contiguous()
        g = ops.ntxent_bwd(local, cols, r * B, 1.0 / tau, lses[r], col_lse, Bg + r * B)
        # gradient w.r.t. the (already unit-norm) candidates: compare with the closed form's gradient projected likewise
        # through the cosine normalisation, which is the identity on the tangent space here
        gj, gi = zjs64.grad[r * B:(r + 1) * B], zis64.grad[r * B:(r + 1) * B]
        n_j, n_i = b[r * B:(r + 1) * B].double(), a[r * B:(r + 1) * B].double()
        proj = lambda gg, n: gg - n * (gg * n).sum(1, keepdim=True)
        assert rel_err(proj(g[:B].double().cpu(), n_j), gj) < RTOL_GRAD
        assert rel_err(proj(g[B:].double().cpu(), n_i), gi) < RTOL_GRAD
