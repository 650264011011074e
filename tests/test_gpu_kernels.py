"""GPU parity tests, kernel by kernel, through the C ABI (python -m pytest tests -m gpu)."""
import numpy as np
import pytest
import torch

from tests.util import tf32_round, rel_err, max_rel

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import ops, Batch
    from molclr_b200.graph import GraphPlan
    from molclr_b200.synth import make_pair_batch, make_plain_batch
    from oracle.csr import build_csr, build_graph_segments
    from oracle import gnn as ognn

DEV = "cuda:0"


def _plan_arrays(plan):
    N, E, G = plan.N, plan.E, plan.G
    return {
        "rowptr": plan.rowptr[:N + 1].cpu().numpy(), "col": plan.col[:E].cpu().numpy(),
        "eattr": plan.eattr[:E].cpu().numpy(), "rowptr_t": plan.rowptr_t[:N + 1].cpu().numpy(),
        "col_t": plan.col_t[:E].cpu().numpy(), "cnt": plan.cnt[:8 * N].cpu().numpy().reshape(N, 8),
        "gptr": plan.gptr[:G + 1].cpu().numpy(), "gperm": plan.gperm[:N].cpu().numpy(),
    }


@pytest.mark.parametrize("bs,seed", [(1, 0), (7, 1), (512, 2)])
def test_plan_bit_exact(bs, seed):
    bi, _ = make_pair_batch(bs, seed=seed)
    plan = GraphPlan(bi.to(DEV))
    got = _plan_arrays(plan)
    want = build_csr(bi.edge_index.numpy(), bi.edge_attr.numpy(), bi.num_nodes)
    for k in ("rowptr", "col", "eattr", "rowptr_t", "col_t", "cnt"):
        assert np.array_equal(got[k], want[k]), k
    gptr, gperm = build_graph_segments(bi.batch.numpy(), bi.num_graphs)
    assert np.array_equal(got["gptr"], gptr) and np.array_equal(got["gperm"], gperm)
    xp = plan.xpacked[:plan.N].cpu().numpy()
    assert np.array_equal(xp & 0xff, bi.x[:, 0].numpy()) and np.array_equal(xp >> 8, bi.x[:, 1].numpy())
    assert np.array_equal(plan.nbr[:8 * plan.N].cpu().numpy().view(np.uint32).reshape(-1, 8), _nbr_table(want, plan.N))


def _nbr_table(csr, N):
    """Fixed-width neighbour table restated from the CSR (include/molclr_b200.h: molclr_plan_build, `nbr`)."""
    t = np.full((N, 8), 0xFFFFFFFF, dtype=np.uint32)
    rp, col, ea = csr["rowptr"], csr["col"], csr["eattr"]
    for n in range(N):
        b, e = int(rp[n]), int(rp[n + 1])
        if e - b <= 8:
            t[n, :e - b] = (col[b:e].astype(np.uint32) << 4) | ea[b:e].astype(np.uint32)
        else:
            t[n, 0] = 0xFFFFFFFE
    return t


@pytest.mark.parametrize("D,drop", [(300, 0.0), (300, 0.3), (512, 0.0), (64, 0.0)])
def test_aggregate_bwd_tile_kernel_equals_row_kernel(D, drop):
    """Backward: gy of the tile kernel (out-edge table given) is bitwise the warp-per-row kernel's; the BatchNorm statistics
    partials are summed in a different (fixed) order, so their totals agree to fp32 rounding."""
    from tests.test_gpu_properties import _random_batch
    g = torch.Generator().manual_seed(D + 7)
    coef = torch.randn(4, D, generator=g)
    coef[0] = coef[0].abs() + 0.5; coef[3] = coef[3].abs() + 0.5
    for b in (make_pair_batch(200, seed=6)[0], _random_batch(np.random.default_rng(4), 30, 70), make_plain_batch(1, seed=2)):
        plan = GraphPlan(b.to(DEV))
        want = build_csr(b.edge_index.numpy(), b.edge_attr.numpy(), b.num_nodes)
        nt = plan.nbr_t[:8 * plan.N].cpu().numpy().view(np.uint32).reshape(-1, 8)
        assert np.array_equal(nt, _nbr_table({"rowptr": want["rowptr_t"], "col": want["col_t"], "eattr": np.zeros_like(want["col_t"])}, plan.N))
        ga, z = torch.randn(plan.N, D, generator=g).to(DEV), torch.randn(plan.N, D, generator=g).to(DEV)
        a, _, _ = ops.gine_aggregate_bwd(plan, ga)
        r, _, _ = ops.gine_aggregate_bwd(plan, ga, use_nbr=False)
        assert torch.equal(a, r)
        dp = (77, drop)
        a, pa, na = ops.gine_aggregate_bwd(plan, ga, z_prev=z, bn_coef=coef.to(DEV), drop=dp, round_out=True)
        r, pr, nr = ops.gine_aggregate_bwd(plan, ga, z_prev=z, bn_coef=coef.to(DEV), drop=dp, round_out=True, use_nbr=False)
        assert torch.equal(a, r)
        sa, sr = pa[:na].double().sum(0), pr[:nr].double().sum(0)
        assert rel_err(sa, sr) < 1e-5, rel_err(sa, sr)


@pytest.mark.parametrize("D,drop", [(300, 0.0), (300, 0.3), (512, 0.0), (64, 0.0)])
def test_aggregate_tile_kernel_equals_row_kernel(D, drop):
    """The shared-memory tile kernel (neighbour table given) and the warp-per-row CSR kernel compute the same sums in the same
    order: bitwise equal outputs, with and without the fused BatchNorm/ReLU/dropout, on molecules and on irregular graphs
    (rows longer than the table, neighbours outside the staged tile, a ragged last tile)."""
    from tests.test_gpu_properties import _random_batch
    g = torch.Generator().manual_seed(D)
    B1, B2 = torch.randn(5, D, generator=g).to(DEV), torch.randn(3, D, generator=g).to(DEV)
    coef = torch.randn(4, D, generator=g)
    coef[0] = coef[0].abs() + 0.5
    for b in (make_pair_batch(200, seed=5)[0], _random_batch(np.random.default_rng(3), 30, 70), make_plain_batch(1, seed=1)):
        plan = GraphPlan(b.to(DEV))
        h = torch.randn(plan.N, D, generator=g).to(DEV)
        for bn in (None, coef.to(DEV)):
            dp = (1234, drop) if bn is not None else (0, 0.0)
            a = ops.gine_aggregate_fwd(plan, h, B1, B2, bn_coef=bn, round_out=False, drop=dp)
            r = ops.gine_aggregate_fwd(plan, h, B1, B2, bn_coef=bn, round_out=False, drop=dp, use_nbr=False)
            assert torch.equal(a[:, :D], r[:, :D])
        a, lo = ops.gine_aggregate_fwd(plan, h, B1, B2, round_out=True, want_lo=True)
        r, rlo = ops.gine_aggregate_fwd(plan, h, B1, B2, round_out=True, want_lo=True, use_nbr=False)
        assert torch.equal(a[:, :D], r[:, :D]) and torch.equal(lo[:, :D], rlo[:, :D])


def test_plan_edge_cases():
    # no edges at all; unsorted batch vector; shuffled edge order (stability)
    x = torch.tensor([[5, 0], [118, 0], [7, 1], [6, 2]])
    b = Batch(x, torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 2, dtype=torch.long), torch.tensor([1, 0, 1, 0]), num_graphs=2)
    plan = GraphPlan(b.to(DEV))
    got = _plan_arrays(plan)
    assert got["rowptr"].tolist() == [0] * 5 and got["gptr"].tolist() == [0, 2, 4] and got["gperm"].tolist() == [1, 3, 0, 2]
    assert (got["cnt"][:, 4] == 1).all() and (got["cnt"][:, 5] == 1).all()
    big = make_plain_batch(40, seed=9)
    perm = torch.randperm(big.edge_index.size(1), generator=torch.Generator().manual_seed(0))
    sh = Batch(big.x, big.edge_index[:, perm], big.edge_attr[perm], big.batch, big.num_graphs)
    got = _plan_arrays(GraphPlan(sh.to(DEV)))
    want = build_csr(sh.edge_index.numpy(), sh.edge_attr.numpy(), sh.num_nodes)
    for k in ("rowptr", "col", "eattr", "rowptr_t", "col_t", "cnt"):
        assert np.array_equal(got[k], want[k]), k


def test_plan_rejects_out_of_range():
    x = torch.tensor([[5, 0], [119, 0]])
    b = Batch(x, torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 2, dtype=torch.long), torch.zeros(2, dtype=torch.long))
    with pytest.raises(IndexError):
        GraphPlan(b.to(DEV))
    x = torch.tensor([[5, 0], [6, 0]])
    b = Batch(x, torch.tensor([[0], [2]]), torch.tensor([[0, 0]]), torch.zeros(2, dtype=torch.long))
    with pytest.raises(IndexError):
        GraphPlan(b.to(DEV))


@pytest.mark.parametrize("D", [300, 32, 128])
def test_embed_and_aggregate_fwd_bit_exact(D):
    torch.manual_seed(0)
    bi, _ = make_pair_batch(96, seed=3)
    plan = GraphPlan(bi.to(DEV))
    conv = ognn.GINEConv(D)
    E1, E2 = torch.randn(119, D), torch.randn(3, D)
    h0_ref = E1[bi.x[:, 0]] + E2[bi.x[:, 1]]
    h0 = ops.embed_nodes_fwd(plan, E1.to(DEV), E2.to(DEV))
    assert torch.equal(h0.cpu(), h0_ref)
    with torch.no_grad():
        ref = conv.aggregate(h0_ref, bi.edge_index, bi.edge_attr)
    got = ops.gine_aggregate_fwd(plan, h0, conv.edge_embedding1.weight.detach().to(DEV), conv.edge_embedding2.weight.detach().to(DEV),
                                 bn_coef=None, round_out=False)
    assert torch.equal(got.cpu(), ref), f"max abs diff {(got.cpu() - ref).abs().max()}"      # bit-exact
    got_r = ops.gine_aggregate_fwd(plan, h0, conv.edge_embedding1.weight.detach().to(DEV), conv.edge_embedding2.weight.detach().to(DEV),
                                   bn_coef=None, round_out=True)
    assert torch.equal(got_r.cpu(), tf32_round(ref))
    # fused BatchNorm + ReLU of the producer layer
    coef = torch.randn(4, D)
    coef[0] = coef[0].abs() + 0.5
    with torch.no_grad():
        ref_bn = conv.aggregate(torch.relu(h0_ref * coef[0] + coef[1]), bi.edge_index, bi.edge_attr)
    got_bn = ops.gine_aggregate_fwd(plan, h0, conv.edge_embedding1.weight.detach().to(DEV), conv.edge_embedding2.weight.detach().to(DEV),
                                    bn_coef=coef.to(DEV), relu=True, round_out=False)
    torch.testing.assert_close(got_bn.cpu(), ref_bn, rtol=1e-5, atol=1e-5)


def test_aggregate_bwd_and_table_grads():
    torch.manual_seed(1)
    D = 300
    bi, _ = make_pair_batch(64, seed=4)
    plan = GraphPlan(bi.to(DEV))
    conv = ognn.GINEConv(D).double()
    h = torch.randn(bi.num_nodes, D, dtype=torch.float64, requires_grad=True)
    ga = torch.randn(bi.num_nodes, D)
    conv.aggregate(h, bi.edge_index, bi.edge_attr).backward(ga.double())
    gy, _, _ = ops.gine_aggregate_bwd(plan, ga.to(DEV))
    assert rel_err(gy, h.grad) < 1e-6
    dB1, dB2 = ops.edge_table_grad(plan, ga.to(DEV))
    # table gradients are skinny split-K TF32 contractions (count / one-hot operand exact, gradient operand truncated to TF32)
    assert rel_err(dB1, conv.edge_embedding1.weight.grad) < 1e-3
    assert rel_err(dB2, conv.edge_embedding2.weight.grad) < 1e-3
    # node-embedding table gradient
    E1 = torch.randn(119, D, dtype=torch.float64, requires_grad=True)
    E2 = torch.randn(3, D, dtype=torch.float64, requires_grad=True)
    (E1[bi.x[:, 0]] + E2[bi.x[:, 1]]).backward(ga.double())
    dE1, dE2 = ops.embed_nodes_bwd(plan, ga.to(DEV))
    assert rel_err(dE1, E1.grad) < 1e-3 and rel_err(dE2, E2.grad) < 1e-3
    # fused ReLU / BatchNorm-statistics variant
    z = torch.randn(bi.num_nodes, D)
    coef = torch.randn(4, D)
    coef[3] = coef[3].abs() + 0.1
    gy2, partials, P = ops.gine_aggregate_bwd(plan, ga.to(DEV), z_prev=z.to(DEV), bn_coef=coef.to(DEV), relu=True)
    mask = (z * coef[0] + coef[1] > 0).double()
    want = h.grad * mask
    assert rel_err(gy2, want) < 1e-6
    s = partials[:P].double().sum(0).cpu()
    xhat = (z.double() - coef[2].double()) * coef[3].double()
    assert rel_err(s[0], want.sum(0)) < 1e-4 and rel_err(s[1], (want * xhat).sum(0)) < 1e-4


def _gemm_ref(A, B, a_mn, b_mn):
    A = A.double().cpu().T if a_mn else A.double().cpu()
    B = B.double().cpu().T if b_mn else B.double().cpu()
    return A @ B.T


@pytest.mark.parametrize("M,N,K", [(128, 160, 32), (300, 600, 300), (1000, 300, 600), (76, 512, 300), (4096, 256, 512)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_layouts(M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = tf32_round(torch.randn((K, M) if a_mn else (M, K), generator=g)).to(DEV)
    B = tf32_round(torch.randn((K, N) if b_mn else (N, K), generator=g)).to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out)
    ref = _gemm_ref(A, B, a_mn, b_mn)
    err = max_rel(out, ref)
    assert err < 2e-5, f"gemm M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: max rel err {err}"


def test_gemm_epilogues():
    g = torch.Generator().manual_seed(5)
    M, N, K = 1000, 600, 300
    A, B = tf32_round(torch.randn(M, K, generator=g)).to(DEV), tf32_round(torch.randn(N, K, generator=g)).to(DEV)
    bias, addend, mask = torch.randn(N, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV)
    ref = _gemm_ref(A, B, False, False)
    T = ops.colstat_tiles(M)
    # bias + relu + rounded output + column sums
    out, part = torch.empty(M, N, device=DEV), torch.empty(T, N, device=DEV)
    ops.gemm(A, B, M, N, K, out=out, bias=bias, relu=True, round_out=True, colstat=part, colstat_mode=1)
    want = torch.relu(ref + bias.double().cpu())
    assert max_rel(out, want) < 1e-3 and torch.equal(out, tf32_round(out))
    assert rel_err(part.double().sum(0), want.sum(0)) < 1e-5          # statistics of the exact result, not of the rounded copy
    # addend + mask, exact output + rounded copy
    out, out2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A, B, M, N, K, out=out, out2=out2, addend=addend, mask=mask)
    want = (ref + addend.double().cpu()) * (mask.cpu() > 0)
    assert max_rel(out, want) < 2e-5 and torch.equal(out2, tf32_round(out))
    # BatchNorm tile statistics (mean, M2)
    out, st = torch.empty(M, N, device=DEV), torch.empty(T, 2, N, device=DEV)
    ops.gemm(A, B, M, N, K, out=out, bias=bias, colstat=st, colstat_mode=2)
    want = ref + bias.double().cpu()
    R = ops.colstat_tile_rows()
    for t in range(-(-M // R)):
        blk = want[t * R:(t + 1) * R]
        assert rel_err(st[t, 0], blk.mean(0)) < 1e-4
        assert rel_err(st[t, 1], ((blk - blk.mean(0)) ** 2).sum(0)) < 1e-4


def _unpack_bits(bits, N):
    """[M, W] int32 words -> bool [M, N] (bit c % 32 of word c // 32)."""
    w = bits.cpu().numpy().view(np.uint32)
    cols = np.arange(N)
    return torch.from_numpy(((w[:, cols // 32] >> (cols % 32).astype(np.uint32)) & 1).astype(bool))


@pytest.mark.parametrize("M,N,K,comp", [(1000, 600, 300, False), (1000, 600, 300, True), (333, 512, 300, False), (200, 300, 64, True)])
def test_gemm_relu_bit_mask_roundtrip(M, N, K, comp):
    """Forward GEMM emits [result > 0] as bits (bit-exact against its own fp32 output); the backward GEMM applies them."""
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g).to(DEV), torch.randn(N, K, generator=g).to(DEV)
    (A_hi, A_lo), (B_hi, B_lo) = ops.split_tf32(A), ops.split_tf32(B)
    bias = torch.randn(N, generator=g).to(DEV)
    bits = ops.relu_bits_buffer(M, N, DEV)
    bits.fill_(-1)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A_hi, B_hi, M, N, K, A_lo=A_lo if comp else None, B_lo=B_lo if comp else None, out=out, bias=bias, relu=True,
             relu_bits=bits)
    assert torch.equal(_unpack_bits(bits, N), out.cpu() > 0)
    # backward-style product masked by the bits, with column sums
    G_, W_ = tf32_round(torch.randn(M, K, generator=g)).to(DEV), tf32_round(torch.randn(N, K, generator=g)).to(DEV)
    T = ops.colstat_tiles(M)
    gu, part = torch.empty(M, N, device=DEV), torch.empty(T, N, device=DEV)
    ops.gemm(G_, W_, M, N, K, out=gu, mask_bits=bits, round_out=True, colstat=part, colstat_mode=1)
    want = _gemm_ref(G_, W_, False, False) * (out.cpu() > 0)
    assert max_rel(gu, want) < 1e-3 and torch.equal(gu == 0, (want == 0).to(DEV) | (gu == 0))
    assert torch.equal((gu != 0).cpu() & ~(out.cpu() > 0), torch.zeros(M, N, dtype=torch.bool))
    assert rel_err(part.double().sum(0), want.sum(0)) < 1e-4          # sums of the exact (unrounded) masked product


@pytest.mark.parametrize("M,N,K,b_mn", [(1000, 600, 300, False), (333, 300, 600, False), (512, 512, 300, True)])
def test_gemm_compensated_three_pass_is_fp32_accurate(M, N, K, b_mn):
    """A_hi*B_hi + A_lo*B_hi + A_hi*B_lo on unrounded fp32 inputs: ~fp32 accuracy (used by the forward pass)."""
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(DEV)
    (A_hi, A_lo), (B_hi, B_lo) = ops.split_tf32(A), ops.split_tf32(B)
    assert torch.equal(A_hi.cpu(), tf32_round(A.cpu())) and torch.equal(A_lo.cpu(), tf32_round(A.cpu() - tf32_round(A.cpu())))
    out, out_lo = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A_hi, B_hi, M, N, K, b_mn=b_mn, A_lo=A_lo, B_lo=B_lo, out=out, out_lo=out_lo)
    ref = _gemm_ref(A, B, False, b_mn)
    single = torch.empty(M, N, device=DEV)
    ops.gemm(A_hi, B_hi, M, N, K, b_mn=b_mn, out=single)
    e3, e1 = rel_err(out, ref), rel_err(single, ref)
    assert e3 < 1e-5 and e1 > 20 * e3, (e3, e1)     # floor: the tensor core's truncating fp32 accumulation
    assert torch.equal(out_lo.cpu(), tf32_round(out.cpu() - tf32_round(out.cpu())))


@pytest.mark.parametrize("M,N,K,b_mn", [(1000, 600, 300, False), (5000, 300, 600, False), (333, 512, 300, True), (130, 160, 40, False)])
def test_gemm_compensated_with_low_half_derived_on_chip(M, N, K, b_mn):
    """B_lo without A_lo: A is UNROUNDED fp32 and the kernel's converter warps form A_lo = tf32(A - trunc_tf32(A)) in shared
    memory -- same ~fp32 accuracy as the explicit (A_hi, A_lo) pair, with one A tensor instead of two."""
    g = torch.Generator().manual_seed(M + N + 1)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(DEV)
    A_pad = ops.padded(M, K, DEV)
    A_pad.copy_(A)
    B_hi, B_lo = ops.split_tf32(B)
    bias = torch.randn(N, generator=g).to(DEV)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A_pad, B_hi, M, N, K, b_mn=b_mn, B_lo=B_lo, out=out, bias=bias)
    ref = _gemm_ref(A, B, False, b_mn) + bias.double().cpu()
    single = torch.empty(M, N, device=DEV)
    ops.gemm(A_pad, B_hi, M, N, K, b_mn=b_mn, out=single, bias=bias)
    e3, e1 = rel_err(out, ref), rel_err(single, ref)
    assert e3 < 1e-5 and e1 > 20 * e3, (e3, e1)


@pytest.mark.parametrize("M,N,K", [(1000, 600, 300), (5000, 300, 600), (130, 160, 40), (4097, 512, 300), (50, 256, 512)])
def test_gemm_compensated_on_chip_bf16_corrections(M, N, K):
    """compensate=True: A and B are the UNROUNDED fp32 tensors.  Pass 1 runs in TF32 on the raw tiles (the tensor core
    truncates), the corrections (A - trunc A) * B and A * (B - trunc B) as kind::f16 MMAs on bf16 tiles the converter warps form
    in shared memory.  The corrections are 2^-10 of the product and their bf16 rounding 2^-9 of that: ~1e-6 relative."""
    g = torch.Generator().manual_seed(M + N + 3)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    A_pad, B_pad = ops.padded(M, K, DEV), ops.padded(N, K, DEV)
    A_pad.copy_(A); B_pad.copy_(B)
    bias = torch.randn(N, generator=g).to(DEV)
    out = torch.empty(M, N, device=DEV)
    bits = ops.relu_bits_buffer(M, N, DEV)
    ops.gemm(A_pad, B_pad, M, N, K, compensate=True, out=out, bias=bias, relu=True, relu_bits=bits)
    ref = torch.relu(_gemm_ref(A, B, False, False) + bias.double().cpu())
    single = torch.empty(M, N, device=DEV)
    ops.gemm(ops.split_tf32(A)[0], ops.split_tf32(B)[0], M, N, K, out=single, bias=bias, relu=True)
    e3, e1 = rel_err(out, ref), rel_err(single, ref)
    assert e3 < 5e-6 and e1 > 20 * e3, (e3, e1)
    with pytest.raises(RuntimeError):       # K-major operands only, no lo tensors
        ops.gemm(A_pad, B_pad, M, N, K, compensate=True, b_mn=True, out=out)


@pytest.mark.parametrize("R,O,I", [(5000, 300, 600), (5000, 600, 300), (4096, 256, 512), (700, 512, 300)])
def test_gemm_weight_gradient_split_k(R, O, I):
    g = torch.Generator().manual_seed(R + O)
    dY, X = tf32_round(torch.randn(R, O, generator=g)).to(DEV), tf32_round(torch.randn(R, I, generator=g)).to(DEV)
    dW = ops.gemm_dw(dY, X)
    ref = dY.double().cpu().T @ X.double().cpu()
    assert max_rel(dW, ref) < 2e-5


def test_bn_finalize_and_pool():
    torch.manual_seed(2)
    D = 300
    bi, _ = make_pair_batch(50, seed=6)
    plan = GraphPlan(bi.to(DEV))
    N = plan.N
    z = (torch.randn(N, D) * 2 + 0.7)
    T = ops.colstat_tiles(N)
    stats = torch.zeros(T, 2, D)
    R = ops.colstat_tile_rows()
    for t in range(-(-N // R)):
        blk = z[t * R:(t + 1) * R].double()
        stats[t, 0], stats[t, 1] = blk.mean(0), ((blk - blk.mean(0)) ** 2).sum(0)
    bn = torch.nn.BatchNorm1d(D)
    bn.weight.data.normal_(); bn.bias.data.normal_()
    rm, rv, nbt = bn.running_mean.clone().to(DEV), bn.running_var.clone().to(DEV), bn.num_batches_tracked.clone().to(DEV)
    coef = ops.bn_fwd_finalize(stats.to(DEV), T, N, bn.weight.detach().to(DEV), bn.bias.detach().to(DEV), rm, rv, nbt, 0.1, 1e-5)
    y_ref = bn(z)
    y = z.to(DEV) * coef[0] + coef[1]
    torch.testing.assert_close(y.cpu(), y_ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rm.cpu(), bn.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv.cpu(), bn.running_var, rtol=1e-5, atol=1e-6)
    assert int(nbt) == 1
    # pool (mean / add) fused with the BN apply
    for name, fn in (("mean", ognn.global_mean_pool), ("add", ognn.global_add_pool)):
        p = ops.pool_fwd(plan, z.to(DEV), coef, ops.POOL_MODES[name], relu=False, round_out=False)
        torch.testing.assert_close(p.cpu(), fn(y_ref.detach(), bi.batch), rtol=1e-4, atol=1e-5)
    # BN backward through the pool
    zz = z.clone().double().requires_grad_(True)
    bn64 = torch.nn.BatchNorm1d(D).double()
    bn64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in bn.state_dict().items()})
    bn64.running_mean.zero_(); bn64.running_var.fill_(1)
    gp = torch.randn(plan.G, D)
    ognn.global_mean_pool(bn64(zz), bi.batch).backward(gp.double())
    partials, P = ops.pool_bwd_stats(plan, gp.to(DEV), z.to(DEV), coef, 0)
    dgamma, dbeta, bcoef = ops.bn_bwd_finalize(partials, P, N, bn.weight.detach().to(DEV), coef, True)
    gz, dbias = ops.bn_bwd_apply(z.to(DEV), bcoef, gp=gp.to(DEV), plan=plan, pool_mode=0)
    assert rel_err(dgamma, bn64.weight.grad) < 1e-4 and rel_err(dbeta, bn64.bias.grad) < 1e-4
    assert rel_err(gz, zz.grad) < 2e-3          # gz is stored tf32-rounded
    assert float(dbias.abs().max()) < 1e-2 * float(gz.abs().max()) * N ** 0.5


def test_l2_normalize():
    torch.manual_seed(0)
    z = torch.randn(300, 256, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(300, 256)
    torch.nn.functional.normalize(z, dim=1).backward(gy.double())
    y, inv = ops.l2_normalize_fwd(z.detach().float().to(DEV), 1e-12)
    gz = ops.l2_normalize_bwd(gy.to(DEV), y, inv, 1e-12)
    assert rel_err(y, torch.nn.functional.normalize(z.detach(), dim=1)) < 1e-6 and rel_err(gz, z.grad) < 1e-5


def test_plan_deferred_check_raises_one_call_late_without_a_sync():
    """get_plan (what model.forward calls) validates without draining the stream: the status word travels to pinned memory behind
    the plan kernels and the IndexError surfaces at the next poll -- the kernels sanitise what they flag, so nothing reads out
    of bounds in between."""
    from molclr_b200.graph import get_plan, poll_checks
    poll_checks(block=True)
    x = torch.tensor([[5, 0], [119, 0]])
    bad = Batch(x, torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 2, dtype=torch.long), torch.zeros(2, dtype=torch.long)).to(DEV)
    plan = get_plan(bad)                      # no exception here
    assert plan.N == 2
    with pytest.raises(IndexError):
        poll_checks(block=True)
    poll_checks(block=True)                   # reported once
    good = make_plain_batch(3, seed=1).to(DEV)
    get_plan(good)
    get_plan(make_plain_batch(3, seed=2).to(DEV))
    poll_checks(block=True)
    with pytest.raises(IndexError):           # explicit synchronous validation still available
        get_plan(Batch(bad.x, bad.edge_index, bad.edge_attr, bad.batch), validate=True)
