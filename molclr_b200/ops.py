"""Thin Python wrappers over the C ABI: allocate outputs with torch, pass raw pointers, raise on error.

Nothing here computes on the CPU or with PyTorch operators; every function enqueues hand-written
sm_100a kernels on the current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, WeightDesc, check, ptr, ptr2d, stream

F32 = torch.float32


def _empty(*shape, device):
    return torch.empty(*shape, dtype=F32, device=device)


def padded(rows, cols, device):
    """[rows, cols] fp32 view whose row stride is a multiple of 32 floats (128-byte aligned rows for TMA)."""
    ld = (cols + 31) // 32 * 32
    return torch.empty(rows, ld, dtype=F32, device=device)[:, :cols]


def max_blocks():
    return _lib.load().molclr_rowwise_max_blocks()


# ------------------------------------------------------------------------------------------- GEMM
def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, A_lo=None, B_lo=None, out=None, out2=None, out_lo=None, bias=None,
         addend=None, mask=None, relu=False, round_out=False, colstat=None, colstat_mode=0, transpose_out=False,
         split_k=1, lda=None, ldb=None, relu_bits=None, mask_bits=None, compensate=False, B16=None, status=None):
    """C[M,N] = sum_k A(m,k) B(n,k) on tcgen05 (TF32 in, FP32 accumulate) with the fused epilogue of
    ``molclr_gemm_tf32`` (see include/molclr_b200.h)."""
    lib = _lib.load()
    a = GemmArgs()
    a.A, a.lda, a.a_mn = ptr2d(A), (A.stride(0) if lda is None else lda), int(a_mn)
    a.B, a.ldb, a.b_mn = ptr2d(B), ((B.stride(0) if B is not None else 0) if ldb is None else ldb), int(b_mn)
    a.A_lo, a.B_lo = ptr2d(A_lo), ptr2d(B_lo)
    a.M, a.N, a.K = M, N, K
    a.out, a.ldo, a.transpose_out = ptr2d(out), (out.stride(0) if out is not None else 0), int(transpose_out)
    a.out2, a.ldo2 = ptr2d(out2), (out2.stride(0) if out2 is not None else 0)
    a.out_lo, a.ldo_lo = ptr2d(out_lo), (out_lo.stride(0) if out_lo is not None else 0)
    a.bias = ptr(bias)
    a.addend, a.ldadd = ptr2d(addend), (addend.stride(0) if addend is not None else 0)
    a.mask, a.ldmask = ptr2d(mask), (mask.stride(0) if mask is not None else 0)
    a.relu, a.round_out = int(relu), int(round_out)
    a.colstat, a.colstat_mode = ptr(colstat), int(colstat_mode)
    a.split_k = int(split_k)
    bits = relu_bits if relu_bits is not None else mask_bits
    a.relu_bits, a.mask_bits = ptr(relu_bits, torch.int32), ptr(mask_bits, torch.int32)
    a.ld_bits = bits.stride(0) if bits is not None else 0
    # compensate: 1 = A, B unrounded fp32 (K-major): ~fp32-accurate product, TF32 pass + bf16 corrections derived on chip;
    # 2 = the fp16 three-product form (B16 = the fp16 tiles of B: prepare_weights W_H16; B may be None)
    a.compensate = int(compensate)
    if B16 is not None:                 # (16-bit tensor [2, rows16, ld16]): B's tiles pre-split by prepare_weights
        a.B16, a.ld16, a.rows16 = B16.data_ptr(), B16.stride(1), B16.shape[1]
    a.status = ptr(status, torch.int32)
    check(lib.molclr_gemm_tf32(C.byref(a), stream()), "gemm_tf32")
    return out


def relu_bits_buffer(M, N, device):
    """[M, words] int32 buffer for the ReLU bit mask of an [M, N] GEMM result."""
    return torch.empty(M, _lib.load().molclr_gemm_mask_words(N), dtype=torch.int32, device=device)


def colstat_tiles(M):
    return _lib.load().molclr_gemm_colstat_tiles(M)


def colstat_tile_rows():
    return _lib.load().molclr_gemm_colstat_tile_rows()


def gemm_dw(dY, X, ordered=False, accumulate_into=None):
    """dW[O,I] = dY^T X for row-major dY [R,O], X [R,I] (both tf32-rounded): the weight gradient of a
    Linear (autograd of ginet_molclr.py:19-23,90-96).  Both operands are consumed MN-major in place;
    the reduction over R is split across CTAs and accumulated atomically (default), or -- ``ordered`` -- written as
    per-split partials and summed in split order (bit-reproducible run to run)."""
    R, O = dY.shape
    I = X.shape[1]
    lib = _lib.load()
    if accumulate_into is not None:       # dW += dY^T X (atomic accumulation, no zero-fill)
        dW = accumulate_into
        check(lib.molclr_gemm_dw_acc(ptr2d(dY), dY.stride(0), ptr2d(X), X.stride(0), R, O, I, ptr2d(dW), dW.stride(0), stream()), "gemm_dw_acc")
        return dW
    dW = _empty(O, I, device=dY.device)
    if ordered:
        nbytes = lib.molclr_gemm_dw_workspace_bytes(R, O, I)
        ws = torch.empty(max(nbytes, 16) // 4, dtype=F32, device=dY.device)
        check(lib.molclr_gemm_dw_ordered(ptr2d(dY), dY.stride(0), ptr2d(X), X.stride(0), R, O, I, ptr(dW), dW.stride(0), ptr(ws), nbytes,
                                         stream()), "gemm_dw_ordered")
    else:
        check(lib.molclr_gemm_dw(ptr2d(dY), dY.stride(0), ptr2d(X), X.stride(0), R, O, I, ptr(dW), dW.stride(0), stream()), "gemm_dw")
    return dW


W_HI, W_LO, W_RAW, W_RAW_T, W_B16, W_HI_T, W_H16, W_T16 = 1, 2, 4, 8, 16, 32, 64, 128


def prepare_weights(specs, want_relaunch=False):
    """ONE launch deriving the tensor-core operand forms of several weights (molclr_prepare_weights).  specs: [(w, flags)] with w a
    2-D fp32 matrix and flags a combination of W_HI (tf32(w)), W_LO (tf32 residual), W_RAW (unrounded copy, 128-byte rows),
    W_RAW_T (W_RAW of w^T: K-major copy of a weight stored [in, out]), W_B16 (bf16 correction tiles [2, rows16, ld16] of the raw
    orientation) or W_H16 (instead: the fp16 halves of 2^6 w, same shape, returned under 'b16' as a float16 view; with W_T16 the 16-bit
    tiles are those of w^T without a raw copy: weights stored [in, out]), W_HI_T (tf32(w^T): the K-major operand of the backward dX product).  Returns a list of dicts with the keys 'hi',
    'lo', 'raw', 'b16', 'hi_t' (None where not requested); with
    ``want_relaunch`` also a callable that re-derives every output from the CURRENT values of the sources into the same buffers."""
    if not specs:
        return []
    dev = specs[0][0].device
    r32 = lambda n: (n + 31) // 32 * 32
    plans, n32, n16 = [], 0, 0
    for w, flags in specs:
        rows, cols = w.shape
        tr = bool(flags & (W_RAW_T | W_T16))
        rt, ct = (cols, rows) if tr else (rows, cols)
        ld_hi, ld_raw, ld16, rows16 = r32(cols), r32(ct), (ct + 63) // 64 * 64, (rt + 255) // 256 * 256
        ld_hi_t = r32(rows)
        o_hi = o_lo = o_raw = o_16 = o_hit = None
        if flags & W_HI_T:
            o_hit, n32 = n32, n32 + cols * ld_hi_t
        if flags & W_HI:
            o_hi, n32 = n32, n32 + rows * ld_hi
        if flags & W_LO:
            o_lo, n32 = n32, n32 + rows * ld_hi
        if flags & (W_RAW | W_RAW_T):
            o_raw, n32 = n32, n32 + rt * ld_raw
        if flags & (W_B16 | W_H16):
            assert not (flags & W_B16 and flags & W_H16), "W_B16 and W_H16 share the 16-bit slot"
            o_16, n16 = n16, n16 + 2 * rows16 * ld16
        plans.append((rows, cols, tr, rt, ct, ld_hi, ld_raw, ld16, rows16, o_hi, o_lo, o_raw, o_16, ld_hi_t, o_hit))
    buf32 = torch.empty(max(n32, 1), dtype=F32, device=dev)
    buf16 = torch.empty(max(n16, 1), dtype=torch.bfloat16, device=dev)
    descs = (WeightDesc * len(specs))()
    out = []
    b32, b16 = buf32.data_ptr(), buf16.data_ptr()
    for d, (w, flags), (rows, cols, tr, rt, ct, ld_hi, ld_raw, ld16, rows16, o_hi, o_lo, o_raw, o_16, ld_hi_t, o_hit) in zip(descs, specs, plans):
        d.src, d.ld_src, d.rows, d.cols = ptr2d(w), w.stride(0), rows, cols
        d.ld_hi, d.ld_raw, d.transpose_raw, d.ld16, d.rows16 = ld_hi, ld_raw, int(tr), ld16, rows16
        d.hi = b32 + 4 * o_hi if o_hi is not None else None
        d.lo = b32 + 4 * o_lo if o_lo is not None else None
        d.raw = b32 + 4 * o_raw if o_raw is not None else None
        d.b16 = b16 + 2 * o_16 if o_16 is not None else None
        d.b16_kind = 1 if flags & W_H16 else 0
        d.hi_t, d.ld_hi_t = (b32 + 4 * o_hit if o_hit is not None else None), ld_hi_t
        view = lambda o, r, ld, c: None if o is None else buf32[o:o + r * ld].view(r, ld)[:, :c]
        out.append({"hi": view(o_hi, rows, ld_hi, cols), "lo": view(o_lo, rows, ld_hi, cols), "raw": view(o_raw, rt, ld_raw, ct),
                    "hi_t": view(o_hit, cols, ld_hi_t, rows),
                    "b16": None if o_16 is None else
                    (buf16[o_16:o_16 + 2 * rows16 * ld16].view(torch.float16) if flags & W_H16 else buf16[o_16:o_16 + 2 * rows16 * ld16]).view(2, rows16, ld16)})
    fn, n = _lib.load().molclr_prepare_weights, len(specs)

    def launch():          # (descs holds raw pointers into buf32 / buf16 / the sources: the closure keeps all of them alive)
        check(fn(descs, n, stream()), "prepare_weights")
    launch._keep = (buf32, buf16, [w for w, _ in specs])
    launch()
    return (out, launch) if want_relaunch else out


def colsum(Mx):
    """Column sums of a small row-major matrix (deterministic)."""
    R, Cc = Mx.shape
    out = _empty(Cc, device=Mx.device)
    check(_lib.load().molclr_reduce_partials(ptr(Mx), R, Cc, 1.0, 0, ptr(out), stream()), "reduce_partials")
    return out


def reduce_partials(partials, P, length, out):
    check(_lib.load().molclr_reduce_partials(ptr(partials), P, length, 1.0, 0, ptr(out), stream()), "reduce_partials")
    return out


def round_tf32(x):
    out = torch.empty_like(x)
    check(_lib.load().molclr_round_tf32(ptr(x), ptr(out), None, x.numel(), stream()), "round_tf32")
    return out


def split_tf32(x):
    """(hi, lo) with hi = tf32(x), lo = tf32(x - hi): operands of the error-compensated product.  2-D inputs
    get 128-byte aligned rows."""
    if x.dim() == 2:
        hi, lo = padded(x.shape[0], x.shape[1], x.device), padded(x.shape[0], x.shape[1], x.device)
        check(_lib.load().molclr_round_tf32_2d(ptr2d(x), x.stride(0), ptr2d(hi), ptr2d(lo), hi.stride(0), x.shape[0], x.shape[1],
                                               stream()), "round_tf32_2d")
        return hi, lo
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(_lib.load().molclr_round_tf32(ptr(x), ptr(hi), ptr(lo), x.numel(), stream()), "round_tf32")
    return hi, lo


# ------------------------------------------------------------------------------------------- row-wise
def embed_nodes_fwd(plan, E1, E2):
    D = E1.shape[1]
    out = _empty(plan.N, D, device=E1.device)
    check(_lib.load().molclr_embed_nodes_fwd(ptr(plan.xpacked, torch.int32), ptr(E1), ptr(E2), plan.N, D, ptr(out), stream()),
          "embed_nodes_fwd")
    return out


def embed_nodes_bwd(plan, g):
    lib = _lib.load()
    D = g.shape[1]
    dE = _empty(119 + 3, D, device=g.device)
    ws = torch.empty(lib.molclr_embed_nodes_bwd_workspace_bytes(plan.N) // 4, dtype=F32, device=g.device)
    check(lib.molclr_embed_nodes_bwd(ptr(plan.xpacked, torch.int32), ptr2d(g), g.stride(0), plan.N, D, ptr(dE), ptr(ws), stream()),
          "embed_nodes_bwd")
    return dE[:119], dE[119:]


def gine_aggregate_fwd(plan, src, B1, B2, bn_coef=None, relu=True, round_out=True, want_lo=False, drop=(0, 0.0), use_nbr=True):
    """Returns the aggregate (tf32-rounded if round_out), plus its tf32 residual when want_lo.  use_nbr=False withholds the
    fixed-width neighbour table, which selects the warp-per-row CSR kernel instead of the shared-memory tile kernel."""
    D = src.shape[1]
    out = padded(plan.N, D, src.device)          # a GEMM operand: 128-byte aligned rows
    lo = padded(plan.N, D, src.device) if want_lo else None
    check(_lib.load().molclr_gine_aggregate_fwd(ptr(src), ptr(bn_coef), int(relu), ptr(plan.rowptr, torch.int32),
                                                ptr(plan.col, torch.int32), ptr(plan.eattr, torch.uint8),
                                                ptr(plan.nbr if use_nbr else None, torch.int32), ptr(B1), ptr(B2),
                                                plan.N, D, ptr2d(out), out.stride(0), int(round_out), ptr2d(lo), drop[0], drop[1],
                                                stream()), "gine_aggregate_fwd")
    return (out, lo) if want_lo else out


def gine_aggregate_bwd(plan, ga, z_prev=None, bn_coef=None, relu=True, round_out=False, drop=(0, 0.0), use_nbr=True):
    """Returns (gy, partials, P): partials/P are None/0 when z_prev is None.  use_nbr=False withholds the out-edge table, which
    selects the warp-per-row CSR kernel instead of the shared-memory tile kernel."""
    D = ga.shape[1]
    gy = torch.empty_like(ga)
    partials = _empty(max_blocks(), 2, D, device=ga.device) if z_prev is not None else None
    n = C.c_int(0)
    check(_lib.load().molclr_gine_aggregate_bwd(ptr(ga), ptr(plan.rowptr_t, torch.int32), ptr(plan.col_t, torch.int32),
                                                ptr(plan.nbr_t if use_nbr else None, torch.int32), ptr(z_prev), ptr(bn_coef), int(relu), plan.N, D, ptr(gy), int(round_out),
                                                ptr(partials), C.byref(n), drop[0], drop[1], stream()), "gine_aggregate_bwd")
    return gy, partials, n.value


def relu_bn_bwd_stats(g, z_prev, bn_coef, relu=True, drop=(0, 0.0)):
    """gy = g * [relu(BN(z_prev)) > 0] and the (sum gy, sum gy*xhat) partials.  Returns (gy, partials, P)."""
    N, D = g.shape
    gy = torch.empty_like(g)
    partials = _empty(max_blocks(), 2, D, device=g.device)
    n = C.c_int(0)
    check(_lib.load().molclr_relu_bn_bwd_stats(ptr(g), ptr(z_prev), ptr(bn_coef), int(relu), N, D, ptr(gy), ptr(partials),
                                               C.byref(n), drop[0], drop[1], stream()), "relu_bn_bwd_stats")
    return gy, partials, n.value


def gcn_aggregate_fwd(plan, src, b1, b2, bias):
    """GCNConv propagate + bias (gcn_molclr.py:79-82) on y = x @ weight; exact fp32 output."""
    D = src.shape[1]
    out = _empty(plan.N, D, device=src.device)
    check(_lib.load().molclr_gcn_aggregate_fwd(ptr2d(src), ptr(plan.rowptr, torch.int32), ptr(plan.col, torch.int32),
                                               ptr(plan.eattr, torch.uint8), ptr(b1.reshape(-1)), ptr(b2.reshape(-1)), ptr(bias),
                                               plan.N, D, ptr(out), D, stream()), "gcn_aggregate_fwd")
    return out


def row_sum(x):
    R, Cc = x.shape
    out = _empty(R, device=x.device)
    check(_lib.load().molclr_row_sum(ptr(x), R, Cc, ptr(out), stream()), "row_sum")
    return out


def bn_apply_fwd(z, bn_coef, relu, want_lo, drop=(0, 0.0), round_hi=True):
    """(hi, lo) tensor-core operand pair of relu(BN(z)) (bn_coef None: of z itself), 128-byte aligned rows; round_hi=False
    leaves hi unrounded (operand of the compensated GEMM that derives its low halves on chip)."""
    N, D = z.shape
    hi = padded(N, D, z.device)
    lo = padded(N, D, z.device) if want_lo else None
    check(_lib.load().molclr_bn_apply_fwd(ptr(z), ptr(bn_coef), int(relu), N, D, ptr2d(hi), ptr2d(lo), hi.stride(0), int(round_hi),
                                          drop[0], drop[1], stream()), "bn_apply_fwd")
    return hi, lo


def bn_tile_stats(z):
    N, D = z.shape
    T = colstat_tiles(N)
    st = _empty(T, 2, D, device=z.device)
    check(_lib.load().molclr_bn_tile_stats(ptr(z), N, D, T, ptr(st), stream()), "bn_tile_stats")
    return st, T


def edge_table_grad_raw(plan, ga):
    """dB [8, D]: rows 0..4 = gradient of the bond-type table, rows 5..7 = of the bond-direction table (exact fp32, fixed order)."""
    D = ga.shape[1]
    lib = _lib.load()
    dB = _empty(8, D, device=ga.device)
    ws = torch.empty(lib.molclr_edge_table_grad_workspace_bytes(D) // 4, dtype=F32, device=ga.device)
    check(lib.molclr_edge_table_grad(ptr2d(ga), ga.stride(0), ptr(plan.cnt), plan.N, D, ptr(dB), ptr(ws), stream()), "edge_table_grad")
    return dB


def edge_table_grad(plan, ga):
    dB = edge_table_grad_raw(plan, ga)
    return dB[:5], dB[5:]


def bn_fwd_finalize(tile_stats, T, N, gamma, beta, running_mean, running_var, nbt, momentum, eps):
    D = gamma.shape[0]
    coef = _empty(4, D, device=gamma.device)
    lib = _lib.load()
    ws = torch.empty(lib.molclr_bn_finalize_workspace_bytes(D) // 8, dtype=torch.float64, device=gamma.device)
    check(lib.molclr_bn_fwd_finalize(ptr(tile_stats), T, colstat_tile_rows(), N, D, ptr(gamma), ptr(beta), ptr(running_mean),
                                     ptr(running_var), ptr(nbt, torch.int64), momentum, eps, ptr(coef), ptr(ws, torch.float64), stream()),
          "bn_fwd_finalize")
    return coef


def bn_eval_coef(gamma, beta, running_mean, running_var, eps):
    D = gamma.shape[0]
    coef = _empty(4, D, device=gamma.device)
    check(_lib.load().molclr_bn_eval_coef(ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), eps, D, ptr(coef), stream()),
          "bn_eval_coef")
    return coef


def bn_bwd_finalize(partials, P, N, gamma, coef, use_batch_stats):
    D = gamma.shape[0]
    dgamma, dbeta, bcoef = _empty(D, device=gamma.device), _empty(D, device=gamma.device), _empty(3, D, device=gamma.device)
    check(_lib.load().molclr_bn_bwd_finalize(ptr(partials), P, N, D, ptr(gamma), ptr(coef), int(use_batch_stats), ptr(dgamma),
                                             ptr(dbeta), ptr(bcoef), stream()), "bn_bwd_finalize")
    return dgamma, dbeta, bcoef


def bn_bwd_apply(z, bcoef, gy=None, gp=None, plan=None, pool_mode=0, round_out=True, argmax=None, drop=(0, 0.0)):
    N, D = z.shape
    gz = padded(N, D, z.device) if round_out else _empty(N, D, device=z.device)   # GEMM operand: 128-byte aligned rows
    dbias = _empty(D, device=z.device)
    partials = _empty(max_blocks(), D, device=z.device)
    n2g = ptr(plan.node2graph, torch.int32) if gp is not None else None
    gptr = ptr(plan.gptr, torch.int32) if gp is not None else None
    check(_lib.load().molclr_bn_bwd_apply(ptr(gy), ptr(gp), n2g, gptr, int(pool_mode), ptr(argmax, torch.int32), ptr(z), ptr(bcoef), N, D,
                                          ptr2d(gz), gz.stride(0), int(round_out), ptr(dbias), ptr(partials), drop[0], drop[1], stream()),
          "bn_bwd_apply")
    return gz, dbias


POOL_MODES = {"mean": 0, "add": 1, "max": 2}


def pool_fwd(plan, z, bn_coef, pool_mode, relu=False, round_out=True, want_lo=False, argmax=None, drop=(0, 0.0)):
    D = z.shape[1]
    out = padded(plan.G, D, z.device)
    lo = padded(plan.G, D, z.device) if want_lo else None
    check(_lib.load().molclr_pool_fwd(ptr(z), ptr(bn_coef), int(relu), ptr(plan.gptr, torch.int32), ptr(plan.gperm, torch.int32),
                                      pool_mode, plan.G, D, ptr2d(out), out.stride(0), int(round_out), ptr2d(lo),
                                      ptr(argmax, torch.int32), drop[0], drop[1], stream()), "pool_fwd")
    return (out, lo) if want_lo else out


def pool_bwd_stats(plan, gp, z, bn_coef, pool_mode, argmax=None, drop=(0, 0.0)):
    D = z.shape[1]
    partials = _empty(max_blocks(), 2, D, device=z.device)
    n = C.c_int(0)
    check(_lib.load().molclr_pool_bwd_stats(ptr(gp), ptr(plan.node2graph, torch.int32), ptr(plan.gptr, torch.int32), pool_mode,
                                            ptr(argmax, torch.int32), ptr(z), ptr(bn_coef), plan.N, D, ptr(partials), C.byref(n),
                                            drop[0], drop[1], stream()), "pool_bwd_stats")
    return partials, n.value


def copy_rows(src, dst, rows):
    """dst[:rows, :cols] = src[:rows, :cols] (2-D fp32, contiguous rows): tiny pad/unpad copies via cudaMemcpy2DAsync."""
    import ctypes as _C
    cols = src.shape[1]
    _cudart_memcpy2d(dst.data_ptr(), dst.stride(0) * 4, src.data_ptr(), src.stride(0) * 4, cols * 4, rows)


def copy_cols(src, dst, cols):
    """dst[:, :cols] = src[:, :cols]."""
    _cudart_memcpy2d(dst.data_ptr(), dst.stride(0) * 4, src.data_ptr(), src.stride(0) * 4, cols * 4, src.shape[0])


def _cudart_memcpy2d(dst, dpitch, src, spitch, width, height):
    check(_lib.load().molclr_copy_2d(dst, dpitch, src, spitch, width, height, stream()), "copy_2d")


def dropout_mask(seed, p, N, D, device):
    """The mask (0 or 1/(1-p)) the fused kernels apply for (seed, p): test / diagnostics hook."""
    out = _empty(N, D, device=device)
    check(_lib.load().molclr_dropout_mask(seed, p, N, D, ptr(out), stream()), "dropout_mask")
    return out


ACT_MODES = {"softplus": 0, "relu": 1}


def act_fwd(x, mode, want_lo):
    """(hi, lo) tensor-core operand pair of act(x) for a contiguous 2-D x whose width is a multiple of 32."""
    hi = torch.empty_like(x)
    lo = torch.empty_like(x) if want_lo else None
    check(_lib.load().molclr_act_fwd(ptr(x), mode, x.numel(), ptr(hi), ptr(lo), stream()), "act_fwd")
    return hi, lo


def act_bwd(gy, x, mode):
    gx = torch.empty_like(x)
    check(_lib.load().molclr_act_bwd(ptr(gy), ptr(x), mode, x.numel(), ptr(gx), stream()), "act_bwd")
    return gx


# ------------------------------------------------------------------------------------------- normalize / NT-Xent
def l2_normalize_fwd(z, eps):
    R, Cc = z.shape
    y, inv = torch.empty_like(z), _empty(R, device=z.device)
    check(_lib.load().molclr_l2_normalize_fwd(ptr(z), R, Cc, eps, ptr(y), ptr(inv), stream()), "l2_normalize_fwd")
    return y, inv


def l2_normalize_bwd(gy, y, inv, eps, gscale=None):
    """gscale: optional 0-dim / 1-element DEVICE tensor multiplied into gy first."""
    R, Cc = y.shape
    gz = torch.empty_like(y)
    if gscale is None:
        check(_lib.load().molclr_l2_normalize_bwd(ptr(gy), ptr(y), ptr(inv), R, Cc, eps, ptr(gz), stream()), "l2_normalize_bwd")
    else:
        check(_lib.load().molclr_l2_normalize_bwd_scaled(ptr(gy), ptr(y), ptr(inv), R, Cc, eps, ptr(gscale.reshape(1)), ptr(gz), stream()),
              "l2_normalize_bwd_scaled")
    return gz


def ntxent_rows_fwd(zA, zB, eps, normalise, want_r=True, want16=False):
    """rep = cat([zA, zB]) with rows optionally L2-normalised: returns (y or None, y_r tf32-rounded or None, inv_norm or None[, y16])."""
    RA, Cc = zA.shape
    RB = zB.shape[0]
    y = _empty(RA + RB, Cc, device=zA.device) if normalise else None
    y_r = _empty(RA + RB, Cc, device=zA.device) if want_r else None
    inv = _empty(RA + RB, device=zA.device) if normalise else None
    ld16 = (Cc + 7) // 8 * 8
    y16 = torch.empty(RA + RB, ld16, dtype=torch.float16, device=zA.device) if want16 else None
    check(_lib.load().molclr_ntxent_rows_fwd(ptr(zA), ptr(zB), RA, RB, Cc, eps, int(bool(normalise)), ptr(y), ptr(y_r), ptr(inv),
                                             ptr(y16, torch.float16), ld16, stream()), "ntxent_rows_fwd")
    return (y, y_r, inv, y16) if want16 else (y, y_r, inv)


def ntxent_h_supported(C_, inv_temperature):
    return bool(_lib.load().molclr_ntxent_h_supported(int(C_), float(inv_temperature)))


def ntxent_fwd_h(rep16, cols16, C_, row_offset, row_offset2, inv_temperature):
    """ntxent_fwd on fp16 rows (rep16 may be a row slice of cols16).  Returns (loss[1], row_lse[R], row_pos[R])."""
    lib = _lib.load()
    R, Rc, ld16 = rep16.shape[0], cols16.shape[0], cols16.stride(0)
    nbytes = lib.molclr_ntxent_workspace_bytes(R, Rc, C_)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cols16.device)
    row_lse, row_pos, loss = _empty(R, device=cols16.device), _empty(R, device=cols16.device), _empty(1, device=cols16.device)
    check(lib.molclr_ntxent_fwd_h(rep16.data_ptr(), cols16.data_ptr(), ld16, R, Rc, C_, row_offset, row_offset2, inv_temperature, ptr(row_lse),
                                  ptr(row_pos), ptr(loss), ptr(ws, torch.uint8), nbytes, stream()), "ntxent_fwd_h")
    return loss, row_lse, row_pos


def ntxent_bwd_h(rep16, cols16, C_, row_offset, row_offset2, inv_temperature, row_lse, col_lse):
    lib = _lib.load()
    R, Rc, ld16 = rep16.shape[0], cols16.shape[0], cols16.stride(0)
    nbytes = lib.molclr_ntxent_workspace_bytes(R, Rc, C_)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cols16.device)
    g = _empty(R, C_, device=cols16.device)
    check(lib.molclr_ntxent_bwd_h(rep16.data_ptr(), cols16.data_ptr(), ld16, R, Rc, C_, row_offset, row_offset2, inv_temperature, ptr(row_lse),
                                  ptr(col_lse), 1.0 / Rc, ptr(g), ptr(ws, torch.uint8), nbytes, stream()), "ntxent_bwd_h")
    return g


def ntxent_fwd(rep, cols, row_offset, inv_temperature, row_offset2=None, unit_rows=False):
    """Returns (loss[1], row_lse[R], row_pos[R]) for local rows `rep` ([zjs; zis] halves at candidate rows row_offset /
    row_offset2) against candidates `cols`.  unit_rows: all rows have norm <= 1 (cosine similarity) -> fp16 tensor-core operands."""
    lib = _lib.load()
    R, Cc = rep.shape
    Rc = cols.shape[0]
    nbytes = lib.molclr_ntxent_workspace_bytes(R, Rc, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rep.device)
    row_lse, row_pos, loss = _empty(R, device=rep.device), _empty(R, device=rep.device), _empty(1, device=rep.device)
    row_offset2 = row_offset + R // 2 if row_offset2 is None else row_offset2
    check(lib.molclr_ntxent_fwd(ptr(rep), ptr(cols), R, Rc, Cc, row_offset, row_offset2, inv_temperature, int(bool(unit_rows)), ptr(row_lse), ptr(row_pos), ptr(loss),
                                ptr(ws, torch.uint8), nbytes, stream()), "ntxent_fwd")
    return loss, row_lse, row_pos


def ntxent_bwd(rep, cols, row_offset, inv_temperature, row_lse, col_lse, row_offset2=None, unit_rows=False):
    lib = _lib.load()
    R, Cc = rep.shape
    Rc = cols.shape[0]
    nbytes = lib.molclr_ntxent_workspace_bytes(R, Rc, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rep.device)
    g = torch.empty_like(rep)
    row_offset2 = row_offset + R // 2 if row_offset2 is None else row_offset2
    check(lib.molclr_ntxent_bwd(ptr(rep), ptr(cols), R, Rc, Cc, row_offset, row_offset2, inv_temperature, int(bool(unit_rows)), ptr(row_lse), ptr(col_lse), 1.0 / Rc,
                                ptr(g), ptr(ws, torch.uint8), nbytes, stream()), "ntxent_bwd")
    return g
