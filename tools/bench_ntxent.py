"""Times NT-Xent forward / backward (CUDA events) at the bench shape and at the per-rank shape of an 8-GPU job
(R = 2 x 4096 local rows against Rc = 8 x 8192 all-gathered candidates), with fp16 (unit_rows) and TF32 operands.
MOLCLR_NTX_STRIPE16 = 1024 | 2048 | 4096 selects the fp16 W stripe width (read once per process)."""
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


shapes = ((8192, 8192), (8192, 16384), (8192, 65536))
modes = [int(m) for m in os.environ.get("NTX_MODES", "1,0").split(",")]
print("stripe16 =", os.environ.get("MOLCLR_NTX_STRIPE16", "default"))
for R, Rc in shapes:
    g = torch.Generator().manual_seed(0)
    cols = ops.round_tf32(torch.nn.functional.normalize(torch.randn(Rc, 256, generator=g), dim=1).to(dev))
    rep = torch.cat([cols[:R // 2], cols[Rc // 2:Rc // 2 + R // 2]]).contiguous()      # rank 0's [zjs; zis] rows
    out = {}
    for unit in modes:
        loss, lse, pos = ops.ntxent_fwd(rep, cols, 0, 10.0, Rc // 2, unit_rows=unit)
        col_lse = torch.full((Rc,), float(lse.mean()), device=dev)
        col_lse[:R // 2] = lse[:R // 2]
        col_lse[Rc // 2:Rc // 2 + R // 2] = lse[R // 2:]
        gr = ops.ntxent_bwd(rep, cols, 0, 10.0, lse, col_lse, Rc // 2, unit_rows=unit)
        out[unit] = (float(loss), gr)
        tf = timeit(lambda: ops.ntxent_fwd(rep, cols, 0, 10.0, Rc // 2, unit_rows=unit))
        tb = timeit(lambda: ops.ntxent_bwd(rep, cols, 0, 10.0, lse, col_lse, Rc // 2, unit_rows=unit))
        fl = 2.0 * R * Rc * 256
        print(f"R={R} Rc={Rc} {'fp16' if unit else 'tf32'}: fwd {tf:8.1f} us ({fl / tf / 1e6:6.1f} TFLOP/s)   bwd {tb:8.1f} us ({2 * fl / tb / 1e6:6.1f} TFLOP/s)"
              f"   loss {float(loss):.6f}", flush=True)
    if len(out) == 2:
        d = (out[1][1] - out[0][1]).norm() / out[0][1].norm()
        print(f"    fp16 vs tf32: loss diff {abs(out[1][0] - out[0][0]):.2e}, grad rel diff {float(d):.2e}", flush=True)
