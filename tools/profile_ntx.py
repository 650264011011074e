"""One NT-Xent forward + backward at the per-rank shape of an 8-GPU job (R = 8192 local rows, Rc = 65536 candidates),
bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from molclr_b200 import ops

dev = "cuda:0"
R, Rc = int(os.environ.get("R", 8192)), int(os.environ.get("RC", 65536))
g = torch.Generator().manual_seed(0)
cols = ops.round_tf32(torch.nn.functional.normalize(torch.randn(Rc, 256, generator=g), dim=1).to(dev))
rep = torch.cat([cols[:R // 2], cols[Rc // 2:Rc // 2 + R // 2]]).contiguous()


def run():
    loss, lse, pos = ops.ntxent_fwd(rep, cols, 0, 10.0, Rc // 2, unit_rows=True)
    col_lse = torch.full((Rc,), float(lse.mean()), device=dev)
    col_lse[:R // 2] = lse[:R // 2]
    col_lse[Rc // 2:Rc // 2 + R // 2] = lse[R // 2:]
    return ops.ntxent_bwd(rep, cols, 0, 10.0, lse, col_lse, Rc // 2, unit_rows=True)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
