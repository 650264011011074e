"""CPU restatement of the ARITHMETIC of the fp16 three-product GEMM (molclr_gemm_args.compensate = 2; csrc/gemm.cu K_PLAIN_H3,
csrc/tables.cu b16_kind = 1): A = A_h + A_l with A_h = fp16(a), A_l = fp16(a - A_h); B_h, B_l the same split of 2^6 W;
C = (A_h B_h + A_l B_h + A_h B_l) * 2^-6 with fp32 accumulation.  Products of two fp16 numbers are exact in fp32, so plain fp32
matmuls of the halves restate what the tensor core computes up to the order of its fp32 additions.  These tests pin the claims
DESIGN.md section 3.1 makes about that form -- accuracy against fp64 next to the TF32 + bf16 form and to a plain fp32 product,
the weight scale, and what happens at both ends of fp16's range -- without a GPU; tests/test_gpu_tables.py checks the kernel
against the same fp64 products."""
import torch

SCALE = 64.0            # MOLCLR_H16_SCALE
F16_MAX = 65504.0


def split_fp16(x, saturate=True):
    x = x.float()
    xs = x.clamp(-F16_MAX, F16_MAX) if saturate else x          # cvt.rn.satfinite
    h = xs.to(torch.float16)
    r = x - h.float()
    l = (r.clamp(-F16_MAX, F16_MAX) if saturate else r).to(torch.float16)
    return h, l


def gemm_fp16x3(A, W):
    ah, al = split_fp16(A)
    bh, bl = split_fp16(W * SCALE)
    ah, al, bh, bl = ah.float(), al.float(), bh.float(), bl.float()
    return (ah @ bh.t() + al @ bh.t() + ah @ bl.t()) / SCALE


def _trunc_tf32(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def gemm_tf32x3(A, W):
    """compensate = 1: TF32 pass on the truncated operands + two bf16 correction passes."""
    ah, wh = _trunc_tf32(A), _trunc_tf32(W)
    bf = lambda t: t.to(torch.bfloat16).float()
    return ah @ wh.t() + bf(A - ah) @ bf(W).t() + bf(A) @ bf(W - wh).t()


def _err(c, ref):
    return float((c.double() - ref).norm() / ref.norm())


def _case(scale_a, scale_w, M=512, K=300, N=600, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g).relu_() * scale_a + torch.randn(M, K, generator=g) * 0.1 * scale_a       # post-ReLU-like rows
    W = torch.randn(N, K, generator=g) * scale_w
    return A, W, A.double() @ W.double().t()


def test_split_carries_22_bits_and_the_halves_are_exact():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(10000, generator=g) * 3.0
    h, l = split_fp16(x)
    assert float((h.double() + l.double() - x.double()).abs().max()) <= float(x.abs().max()) * 2.0 ** -22
    # the residual a - fp16(a) is exactly representable in fp32 (Sterbenz-type difference of neighbours)
    assert torch.equal((x - h.float()).double(), x.double() - h.double())


def test_three_product_form_is_fp32_accurate_over_the_working_range():
    for sa in (0.05, 1.0, 30.0, 3000.0):
        for sw in (0.004, 0.05, 1.0):
            A, W, ref = _case(sa, sw)
            e_h3, e_t3, e_f32 = _err(gemm_fp16x3(A, W), ref), _err(gemm_tf32x3(A, W), ref), _err(A @ W.t(), ref)
            single = _err(_trunc_tf32(A) @ _trunc_tf32(W).t(), ref)
            assert e_h3 < 2.5 * e_f32 + 1e-7, (sa, sw, e_h3, e_f32)                # as good as a plain fp32 product
            assert e_h3 < e_t3, (sa, sw, e_h3, e_t3)                                # better than the TF32 + bf16 form
            assert e_h3 < 1e-2 * single, (sa, sw, e_h3, single)                     # two orders below single-pass TF32


def test_weight_scale_keeps_the_low_halves_of_small_weights():
    """Without the 2^6 scale the low half of a weight of magnitude 0.01 is an fp16 subnormal (step 6e-8 against a value of
    2.4e-6): the scaled split is several times more accurate there (measured 8x), and exact to undo (a power of two)."""
    g = torch.Generator().manual_seed(2)
    w = torch.randn(20000, generator=g) * 0.01
    h, l = split_fp16(w)
    hs, ls = split_fp16(w * SCALE)
    e_plain = float((h.double() + l.double() - w.double()).abs().max())
    e_scaled = float(((hs.double() + ls.double()) / SCALE - w.double()).abs().max())
    assert e_scaled < 0.25 * e_plain, (e_scaled, e_plain)
    assert float((w * SCALE / SCALE - w).abs().max()) == 0.0
    assert float((torch.tensor([1023.0]) * SCALE).to(torch.float16)) < float("inf")            # |W| < 1024 stays finite


def test_range_ends_saturate_and_degrade_gracefully():
    A, W, ref = _case(1.0, 0.1, M=64)
    A[3, 7] = 1.0e6                                   # far beyond fp16: both halves clamp -- finite, wrong in that row only, and detectable
    ref = A.double() @ W.double().t()
    c = gemm_fp16x3(A, W)
    assert bool(torch.isfinite(c).all()) and float(A.abs().max()) > F16_MAX
    ok = torch.ones(64, dtype=torch.bool); ok[3] = False
    assert _err(c[ok], ref[ok]) < 1e-6 and _err(c[3:4], ref[3:4]) > 1e-3
    h, l = split_fp16(torch.tensor([1.0e5, -7.0e4, 131008.0]))
    assert bool(torch.isfinite(h.float()).all()) and bool(torch.isfinite(l.float()).all())
    assert float(h[0]) == F16_MAX and float(h[0]) + float(l[0]) == 1.0e5        # up to 2 x 65504 the two halves still add up
    # small rows: an absolute error of ~2^-25 per element -- between fp32 and single-pass TF32, never worse than the latter
    for sa, bound in ((1e-3, 1e-3), (1e-4, 1e-2)):
        A, W, ref = _case(sa, 0.1)
        e = _err(gemm_fp16x3(A, W), ref)
        single = _err(_trunc_tf32(A) @ _trunc_tf32(W).t(), ref)
        assert e < single and e < bound, (sa, e, single)
