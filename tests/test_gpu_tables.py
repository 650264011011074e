"""GPU tests of the weight-shadow kernel, the fp32 table gradients and the ordered (bit-reproducible) weight gradient."""
import pytest
import torch

from tests.util import rel_err, tf32_round

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import ops, get_plan
    from molclr_b200.synth import make_pair_batch

DEV = "cuda:0"


def _trunc_tf32(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def test_prepare_weights_all_operand_forms_bit_exact():
    """hi = tf32(w), lo = tf32(w - hi), raw (also of the transpose), bf16 correction tiles bf16(raw), bf16(raw - trunc_tf32(raw)),
    zero padded -- all from ONE launch over several weights of different shapes."""
    torch.manual_seed(0)
    ws = [torch.randn(600, 300, device=DEV) * 0.1, torch.randn(300, 600, device=DEV) * 3.0, torch.randn(300, 300, device=DEV),
          torch.randn(12, 20, device=DEV)]
    flags = [ops.W_HI | ops.W_RAW | ops.W_B16, ops.W_HI | ops.W_LO, ops.W_HI | ops.W_RAW_T | ops.W_B16, ops.W_HI | ops.W_LO | ops.W_RAW | ops.W_B16]
    outs = ops.prepare_weights(list(zip(ws, flags)))
    for w, f, o in zip(ws, flags, outs):
        hi = tf32_round(w)
        assert torch.equal(o["hi"], hi)
        if f & ops.W_LO:
            assert torch.equal(o["lo"], tf32_round(w - hi))
        else:
            assert o["lo"] is None
        if f & (ops.W_RAW | ops.W_RAW_T):
            raw = w.t().contiguous() if f & ops.W_RAW_T else w
            assert torch.equal(o["raw"], raw) and o["raw"].stride(0) % 32 == 0
            if f & ops.W_B16:
                b = o["b16"]
                R, K = raw.shape
                assert b.dtype == torch.bfloat16 and b.shape[1] % 256 == 0 and b.shape[1] >= R and b.shape[2] % 8 == 0
                assert torch.equal(b[0, :R, :K], raw.to(torch.bfloat16))
                assert torch.equal(b[1, :R, :K], (raw - _trunc_tf32(raw)).to(torch.bfloat16))
                assert float(b[:, R:].abs().max()) == 0.0 and float(b[:, :, K:].abs().max()) == 0.0
        else:
            assert o["raw"] is None and o["b16"] is None


@pytest.mark.parametrize("M,N,K", [(1000, 600, 300), (777, 300, 600), (300, 300, 300)])
def test_compensated_gemm_with_presplit_weight_tiles_is_bit_identical(M, N, K):
    """The TMA-loaded, pre-split bf16 tiles of B are the very values the converter warps derive on chip: same result, bit for bit."""
    torch.manual_seed(M)
    A = ops.padded(M, K, DEV); A.copy_(torch.randn(M, K, device=DEV))
    W = torch.randn(N, K, device=DEV) * 0.1
    o = ops.prepare_weights([(W, ops.W_RAW | ops.W_B16)])[0]
    bias = torch.randn(N, device=DEV)
    c0, c1 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A, o["raw"], M, N, K, compensate=True, out=c0, bias=bias, relu=True)
    ops.gemm(A, o["raw"], M, N, K, compensate=True, B16=o["b16"], out=c1, bias=bias, relu=True)
    assert torch.equal(c0, c1)
    ref = torch.relu(A.double() @ W.double().t() + bias.double())
    assert rel_err(c1, ref) < 5e-6          # (fp32 accumulation over K = 600: measured 3.2e-6)


def test_prepare_weights_fp16_halves_bit_exact():
    """W_H16: h = fp16(2^6 w), l = fp16(2^6 w - h), zero padded; h + l carries 22 significand bits of the weight."""
    torch.manual_seed(1)
    for w, fl in [(torch.randn(600, 300, device=DEV) * 0.1, ops.W_HI | ops.W_H16), (torch.randn(40, 600, device=DEV) * 2.0, ops.W_H16)]:
        o = ops.prepare_weights([(w, fl)])[0]
        b = o["b16"]
        R, K = w.shape
        assert b.dtype == torch.float16 and b.shape[1] % 256 == 0 and b.shape[1] >= R and b.shape[2] % 8 == 0 and o["raw"] is None
        s = w * 64.0
        h = s.to(torch.float16)
        assert torch.equal(b[0, :R, :K], h)
        assert torch.equal(b[1, :R, :K], (s - h.float()).to(torch.float16))
        assert float(b[:, R:].abs().max()) == 0.0 and float(b[:, :, K:].abs().max()) == 0.0
        assert float(((b[0, :R, :K].double() + b[1, :R, :K].double()) / 64.0 - w.double()).abs().max()) <= float(w.abs().max()) * 2.0 ** -21


@pytest.mark.parametrize("M,N,K", [(1000, 600, 300), (777, 300, 600), (300, 300, 300), (5000, 600, 300)])
@pytest.mark.parametrize("scale", [0.05, 1.0, 40.0])
def test_fp16_three_product_gemm_is_fp32_accurate(M, N, K, scale):
    """compensate = 2: A split into two fp16 tiles on chip, B pre-split (W_H16), three kind::f16 products per K = 16.  At least as
    accurate as the TF32 + bf16 form (compensate = 1), bias / ReLU / ReLU bits / BatchNorm tile statistics as in the plain epilogue,
    bit-reproducible, status word untouched."""
    torch.manual_seed(M + K)
    A = ops.padded(M, K, DEV); A.copy_(torch.randn(M, K, device=DEV).relu_() * scale + torch.randn(M, K, device=DEV) * 0.1 * scale)
    W = torch.randn(N, K, device=DEV) * 0.1
    o = ops.prepare_weights([(W, ops.W_RAW | ops.W_B16)])[0]
    h = ops.prepare_weights([(W, ops.W_H16)])[0]
    bias = torch.randn(N, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    c1, c2, c3 = ops.padded(M, N, DEV), ops.padded(M, N, DEV), ops.padded(M, N, DEV)
    bits = ops.relu_bits_buffer(M, N, DEV)
    ops.gemm(A, o["raw"], M, N, K, compensate=1, B16=o["b16"], out=c1, bias=bias, relu=True)
    ops.gemm(A, None, M, N, K, compensate=2, B16=h["b16"], out=c2, bias=bias, relu=True, relu_bits=bits, status=status)
    ops.gemm(A, None, M, N, K, compensate=2, B16=h["b16"], out=c3, bias=bias, relu=True, status=status)
    assert torch.equal(c2, c3)
    ref = torch.relu(A.double() @ W.double().t() + bias.double())
    e1, e2 = rel_err(c1, ref), rel_err(c2, ref)
    assert e2 < 5e-6 and e2 < 1.5 * e1 + 1e-7, (e1, e2)
    assert int(status.item()) == 0
    words = bits[:, :(N + 31) // 32].cpu().numpy().astype("uint32")
    got = ((words[:, :, None] >> torch.arange(32).numpy().astype("uint32")[None, None, :]) & 1).reshape(M, -1)[:, :N]
    assert (got == (c2[:, :N] > 0).cpu().numpy()).all()
    # BatchNorm tile statistics (no ReLU): (mean, M2) per 32-row group merge to the column mean / variance of the result
    T = ops.colstat_tiles(M)
    part = torch.zeros(T, 2 * N, device=DEV)
    z = ops.padded(M, N, DEV)
    ops.gemm(A, None, M, N, K, compensate=2, B16=h["b16"], out=z, bias=bias, colstat=part, colstat_mode=2)
    zr = A.double() @ W.double().t() + bias.double()
    assert rel_err(z, zr) < 5e-6
    rows = torch.tensor([min(32, max(0, M - 32 * g)) for g in range(T)], device=DEV, dtype=torch.float64)
    mean_t, m2_t = part.view(T, 2, N)[:, 0].double(), part.view(T, 2, N)[:, 1].double()
    mean = (mean_t * rows[:, None]).sum(0) / M
    var = (m2_t.sum(0) + (rows[:, None] * (mean_t - mean) ** 2).sum(0)) / M
    assert rel_err(mean, zr.mean(0)) < 1e-5 and rel_err(var, zr.var(0, unbiased=False)) < 1e-5


def test_fp16_three_product_gemm_reports_out_of_range_operands():
    """An A element beyond fp16's finite range is clamped (finite result) and reported through the sticky status word."""
    torch.manual_seed(0)
    M, N, K = 512, 300, 300
    A = ops.padded(M, K, DEV); A.copy_(torch.randn(M, K, device=DEV))
    A[100, 7] = 1.0e5
    W = torch.randn(N, K, device=DEV) * 0.1
    h = ops.prepare_weights([(W, ops.W_H16)])[0]
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    c = ops.padded(M, N, DEV)
    ops.gemm(A, None, M, N, K, compensate=2, B16=h["b16"], out=c, status=status)
    assert int(status.item()) & 1
    assert bool(torch.isfinite(c[:, :N]).all())
    ok = torch.ones(M, dtype=torch.bool, device=DEV); ok[100] = False
    assert rel_err(c[ok][:, :N], (A.double() @ W.double().t())[ok]) < 5e-6           # the other rows are unaffected
    # tiny operands keep an ABSOLUTE accuracy of ~2^-25 per element: never worse than single-pass TF32
    A.copy_(torch.randn(M, K, device=DEV) * 1e-4)
    status.zero_()
    ops.gemm(A, None, M, N, K, compensate=2, B16=h["b16"], out=c, status=status)
    assert int(status.item()) == 0
    assert rel_err(c[:, :N], A.double() @ W.double().t()) < 2e-3


def test_table_gradients_fp32_exact_and_reproducible():
    """dB = cnt^T g (bond tables) and dE = onehot^T g (atom / chirality tables): fp32 with a fixed order -- equal to the fp64
    scatter sums to fp32 rounding, identical run to run."""
    bi, _ = make_pair_batch(700, seed=3)
    plan = get_plan(bi.to(DEV))
    N, D = plan.N, 300
    torch.manual_seed(1)
    g = ops.padded(N, D, DEV); g.copy_(torch.randn(N, D, device=DEV) * (1 + 10 * torch.rand(N, 1, device=DEV)))
    dB = ops.edge_table_grad_raw(plan, g)
    cnt = plan.cnt[:8 * N].view(N, 8).double()
    assert rel_err(dB, cnt.t() @ g.double()) < 5e-6
    assert torch.equal(dB, ops.edge_table_grad_raw(plan, g))
    d1, d2 = ops.embed_nodes_bwd(plan, g)
    x = bi.x.to(DEV)
    r1 = torch.zeros(119, D, dtype=torch.float64, device=DEV).index_add_(0, x[:, 0], g.double())
    r2 = torch.zeros(3, D, dtype=torch.float64, device=DEV).index_add_(0, x[:, 1], g.double())
    assert rel_err(d1, r1) < 5e-6 and rel_err(d2, r2) < 5e-6
    e1, e2 = ops.embed_nodes_bwd(plan, g)
    assert torch.equal(d1, e1) and torch.equal(d2, e2)
    # a tiny batch (fewer rows than CTAs)
    bi, _ = make_pair_batch(1, seed=4)
    plan = get_plan(bi.to(DEV))
    g = torch.randn(plan.N, 64, device=DEV)
    d1, d2 = ops.embed_nodes_bwd(plan, g)
    x = bi.x.to(DEV)
    assert rel_err(d1, torch.zeros(119, 64, dtype=torch.float64, device=DEV).index_add_(0, x[:, 0], g.double())) < 1e-6
    assert rel_err(ops.edge_table_grad_raw(plan, g), plan.cnt[:8 * plan.N].view(plan.N, 8).double().t() @ g.double()) < 1e-6


@pytest.mark.parametrize("R,O,I", [(20000, 600, 300), (20000, 300, 600), (4096, 512, 512), (4096, 256, 512), (96, 4, 256), (5000, 300, 300)])
def test_ordered_weight_gradient_matches_and_is_bit_reproducible(R, O, I):
    torch.manual_seed(R + O)
    dY = ops.padded(R, O, DEV); dY.copy_(tf32_round(torch.randn(R, O, device=DEV)))
    X = ops.padded(R, I, DEV); X.copy_(tf32_round(torch.randn(R, I, device=DEV)))
    ref = dY.double().t() @ X.double()
    a = ops.gemm_dw(dY, X, ordered=True)
    b = ops.gemm_dw(dY, X, ordered=True)
    assert torch.equal(a, b)
    assert rel_err(a, ref) < 1e-5
    c = ops.gemm_dw(dY, X)
    assert rel_err(c, ref) < 1e-5
    acc = c.clone()
    ops.gemm_dw(dY, X, accumulate_into=acc)                   # dW += dY^T X (no zero-fill)
    assert rel_err(acc, 2 * ref) < 1e-5


def test_prepare_weights_transposed_tf32_copy():
    torch.manual_seed(3)
    w = torch.randn(600, 300, device=DEV)
    o = ops.prepare_weights([(w, ops.W_HI | ops.W_HI_T)])[0]
    assert torch.equal(o["hi_t"], tf32_round(w.t().contiguous())) and o["hi_t"].stride(0) % 32 == 0 and o["hi_t"].shape == (300, 600)
