"""Times NT-Xent forward / backward (CUDA events) at the bench shape and at a multi-rank candidate count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import ops

dev = "cuda:0"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for R, Rc in ((8192, 8192), (4096, 16384), (4096, 65536)):
    g = torch.Generator().manual_seed(0)
    cols = ops.round_tf32(torch.nn.functional.normalize(torch.randn(Rc, 256, generator=g), dim=1).to(dev))
    rep = cols[:R].contiguous()
    loss, lse, pos = ops.ntxent_fwd(rep, cols, 0, 10.0)
    col_lse = torch.full((Rc,), float(lse.mean()), device=dev)
    col_lse[:R] = lse
    tf = timeit(lambda: ops.ntxent_fwd(rep, cols, 0, 10.0))
    tb = timeit(lambda: ops.ntxent_bwd(rep, cols, 0, 10.0, lse, col_lse))
    fl = 2.0 * R * Rc * 256
    print(f"R={R} Rc={Rc}: fwd {tf:8.1f} us ({fl / tf / 1e6:6.1f} TFLOP/s)   bwd {tb:8.1f} us ({2 * fl / tb / 1e6:6.1f} TFLOP/s)")
