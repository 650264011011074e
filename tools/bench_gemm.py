"""Times the tensor-core GEMM on the hot-path shapes (CUDA events), prints TFLOP/s and operand bytes/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import ops

dev = "cuda:0"
M = int(os.environ.get("M", 102400))


def timeit(fn, iters=int(os.environ.get('ITERS', 10))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3   # us


ONLY = os.environ.get("CASE")


def case(name, M, N, K, comp=False, b_mn=False, relu=False, pad_k=0, dw=False, masked=False, stat=0, derive=False, mixed=False, presplit=False, h3=False):
    if ONLY and not name.startswith(ONLY):
        return
    g = torch.Generator().manual_seed(0)
    if masked or stat:
        A = ops.round_tf32(torch.randn(M, K, generator=g).to(dev))
        B = ops.round_tf32((torch.randn(K, N, generator=g) if b_mn else torch.randn(N, K, generator=g)).to(dev))
        out = ops.padded(M, N, dev)
        bits = ops.relu_bits_buffer(M, N, dev)
        bits.random_(-2**31, 2**31 - 1)
        T = ops.colstat_tiles(M)
        part = torch.empty(T, (2 if stat == 2 else 1) * N, device=dev) if stat else None
        fn = lambda: ops.gemm(A, B, M, N, K, b_mn=b_mn, out=out, mask_bits=bits if masked else None, round_out=True, colstat=part,
                              colstat_mode=stat)
        us = timeit(fn)
        print(f"{name:34s} {us:9.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  out {M * N * 4 / us / 1e3:7.1f} GB/s")
        return
    if dw:
        dY, X = torch.randn(M, N, generator=g).to(dev), torch.randn(M, K, generator=g).to(dev)
        us = timeit(lambda: ops.gemm_dw(dY, X))
        us_o = timeit(lambda: ops.gemm_dw(dY, X, ordered=True))
        flops = 2.0 * M * N * K
        print(f"{name:34s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s   ordered (fixed-order split-K): {us_o:9.1f} us")
        return
    A = torch.randn(M, K + pad_k, generator=g).to(dev)[:, :K]
    B = (torch.randn(K, N, generator=g) if b_mn else torch.randn(N, K + pad_k, generator=g)).to(dev)
    if not b_mn:
        B = B[:, :K]
    out = torch.empty(M, N, device=dev)
    lo = torch.empty(M, N, device=dev) if comp else None
    A_lo = torch.randn(M, K + pad_k, generator=g).to(dev)[:, :K] if (comp and not derive and not mixed) else None
    if derive:
        lo = None
    B_lo = (torch.randn_like(B) if (comp and not mixed) else None)
    if mixed:
        lo = None
    bias = torch.randn(N, device=dev)
    B16 = None
    if h3:                # the fp16 three-product form (compensate = 2): fp16 halves of the weight from molclr_prepare_weights
        B16 = ops.prepare_weights([(B.contiguous(), ops.W_H16)])[0]["b16"]
        A2 = ops.padded(M, K, dev); A2.copy_(A); A = A2
        bits = ops.relu_bits_buffer(M, N, dev)
        out = ops.padded(M, N, dev) if not os.environ.get("UNPADDED_OUT") else torch.empty(M, N, device=dev)      # (row pitch 320 vs 300 floats)
        T = ops.colstat_tiles(M)
        part = torch.empty(T, 2 * N, device=dev) if stat else None
        fn = lambda: ops.gemm(A, None, M, N, K, compensate=2, B16=B16, out=out, bias=bias, relu=relu, relu_bits=bits if relu else None,
                              colstat=part, colstat_mode=stat)
    elif presplit:          # the product path: weight shadows from molclr_prepare_weights (unrounded K-major copy + pre-split bf16 tiles)
        sh = ops.prepare_weights([(B.contiguous(), ops.W_RAW | ops.W_B16)])[0]
        B, B16 = sh["raw"], sh["b16"]
        A2 = ops.padded(M, K, dev); A2.copy_(A); A = A2
        bits = ops.relu_bits_buffer(M, N, dev)
        out = ops.padded(M, N, dev)
        fn = lambda: ops.gemm(A, B, M, N, K, compensate=True, B16=B16, out=out, bias=bias, relu=relu, relu_bits=bits if relu else None)
    else:
        fn = lambda: ops.gemm(A, B, M, N, K, b_mn=b_mn, A_lo=A_lo, B_lo=B_lo, out=out, out_lo=lo, bias=bias, relu=relu, round_out=True,
                              lda=K + pad_k, ldb=(N if b_mn else K + pad_k), compensate=mixed)
    us = timeit(fn)
    flops = 2.0 * M * N * K * (3 if comp else 1)
    print(f"{name:34s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s (tensor work)  out {M * N * 4 * (2 if comp else 1) / us / 1e3:7.1f} GB/s")


print("MOLCLR_GEMM_DEBUG =", os.environ.get("MOLCLR_GEMM_DEBUG"))
if os.environ.get("QUICK"):
    case("fwd1 x1   [M,300]x[600,300]", M, 600, 300)
    case("tinyK x1  [M,32]x[600,32]", M, 600, 32)
    case("sq   x1   [M,512]x[512,512]", M, 512, 512)
    sys.exit(0)
case("fwd1 x1   [M,300]x[600,300]", M, 600, 300)
case("step fwd1 mixed+presplit B, relu bits", M, 600, 300, comp=True, mixed=True, presplit=True, relu=True)
case("step fwd2 mixed+presplit B", M, 300, 600, comp=True, mixed=True, presplit=True)
case("step fwd1 fp16x3, relu bits", M, 600, 300, comp=True, h3=True, relu=True)
case("step fwd2 fp16x3, bn stats", M, 300, 600, comp=True, h3=True, stat=2)
case("fwd1 x3   [M,300]x[600,300]", M, 600, 300, comp=True)
case("fwd1 x3 derive, aligned", M, 600, 300, comp=True, derive=True, pad_k=20)
case("fwd1 mixed (on-chip bf16 corr.)", M, 600, 300, comp=True, mixed=True, pad_k=20)
case("fwd2 mixed (on-chip bf16 corr.)", M, 300, 600, comp=True, mixed=True, pad_k=8)
case("fwd1 x3 explicit lo, aligned", M, 600, 300, comp=True, pad_k=20)
case("fwd2 x3 derive, aligned", M, 300, 600, comp=True, derive=True, pad_k=8)
case("fwd2 x3 explicit lo, aligned", M, 300, 600, comp=True, pad_k=8)
case("fwd2 x1   [M,600]x[300,600]", M, 300, 600)
case("fwd2 x3   [M,600]x[300,600]", M, 300, 600, comp=True)
case("fwd1 x1 K=320 aligned pitch", M, 600, 320)
case("fwd1 x1 pitch 304->320", M, 600, 300, pad_k=20)
case("dX   x1   [M,300]x[300,600]mn", M, 600, 300, b_mn=True)
case("dX   x1   [M,600]x[600,300]mn", M, 300, 600, b_mn=True)
case("sq   x1   [M,512]x[512,512]", M, 512, 512)
case("dU masked+colsum [M,300]x[300,600]mn", M, 600, 300, b_mn=True, masked=True, stat=1)
case("dU masked        [M,300]x[300,600]mn", M, 600, 300, b_mn=True, masked=True)
case("dU colsum        [M,300]x[300,600]mn", M, 600, 300, b_mn=True, stat=1)
case("fwd2 x1 bn-stats [M,600]x[300,600]", M, 300, 600, stat=2)
case("dW   [M,300]^T[M,600]", M, 300, 600, dw=True)
case("dW   [M,600]^T[M,300]", M, 600, 300, dw=True)
