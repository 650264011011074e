"""world_size-2 gloo tests (CPU) of the data-parallel host logic in molclr_b200/dist.py: gather ordering, row offsets,
log-sum-exp exchange, loss shares and gradient scaling -- checked against the single-process fp64 closed form over
the rank-concatenated batch.  The five kernels the collective layer calls are replaced by a
plain-torch stand-in that implements the SAME contracts as the C-ABI entry points (include/molclr_b200.h:
molclr_ntxent_fwd / molclr_ntxent_bwd / molclr_l2_normalize_*), so what is exercised is the product's dist.py."""
import os
import socket
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class TorchKernels:
    """CPU stand-in with the contracts of the CUDA kernels (test infrastructure)."""

    @staticmethod
    def normalize(z, dim=1, eps=1e-12):
        return torch.nn.functional.normalize(z, dim=dim, eps=eps)

    @staticmethod
    def rows_fwd(zjs, zis, eps, normalise, inv_temperature):
        rep = torch.cat([zjs, zis])
        if not normalise:
            return None, rep.clone(), None
        inv = 1.0 / rep.norm(dim=1).clamp_min(eps)
        y = rep * inv[:, None]
        return y, y.clone(), inv

    @staticmethod
    def l2_normalize_bwd(gy, y, inv, eps, gscale=None):
        if gscale is not None:
            gy = gy * gscale
        d = (gy * y).sum(dim=1, keepdim=True)
        d = torch.where((inv >= 1.0 / eps)[:, None], torch.zeros_like(d), d)
        return (gy - y * d) * inv[:, None]

    @staticmethod
    def _logits(rep, cols, row_offset, inv_t, row_offset2):
        R, Rc = rep.shape[0], cols.shape[0]
        lg = (rep.double() @ cols.double().T) * inv_t
        h = R // 2
        rows = torch.cat([torch.arange(h) + row_offset, torch.arange(R - h) + row_offset2])
        self_mask = torch.arange(Rc)[None, :] == rows[:, None]
        pos = torch.cat([torch.arange(h) + row_offset2, torch.arange(R - h) + row_offset])     # partner row of the other block
        return lg, self_mask, pos

    @classmethod
    def ntxent_fwd(cls, rep, cols, C, row_offset, inv_t, row_offset2, unit_rows):
        lg, self_mask, pos = cls._logits(rep, cols, row_offset, inv_t, row_offset2)
        row_pos = lg[torch.arange(rep.shape[0]), pos]
        row_lse = torch.logsumexp(lg.masked_fill(self_mask, float("-inf")), dim=1)
        loss = ((row_lse - row_pos).sum() / cols.shape[0]).reshape(1)
        return loss.to(rep.dtype), row_lse.to(rep.dtype), row_pos.to(rep.dtype)

    @classmethod
    def ntxent_bwd(cls, rep, cols, C, row_offset, inv_t, row_lse, col_lse, row_offset2, unit_rows):
        lg, self_mask, pos = cls._logits(rep, cols, row_offset, inv_t, row_offset2)
        Rc = cols.shape[0]
        w = torch.exp(lg - row_lse.double()[:, None]) + torch.exp(lg - col_lse.double()[None, :])
        w = w.masked_fill(self_mask, 0.0)
        w[torch.arange(rep.shape[0]), pos] -= 2.0
        return ((inv_t / Rc) * (w @ cols.double())).to(rep.dtype)


class ToyEncoder(torch.nn.Module):
    """forward(x) -> (h, out): stands in for the encoder so the parameter-gradient exchange can be checked on CPU."""

    def __init__(self, c_in=12, c_out=8):
        super().__init__()
        self.lin = torch.nn.Linear(c_in, c_out)

    def forward(self, x):
        h = torch.tanh(self.lin(x))
        return h, h


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, global_neg, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from molclr_b200.dist import DataParallelStep
        from oracle.nt_xent import NTXentRestated
        torch.manual_seed(100 + rank)                         # deliberately DIFFERENT initial weights per rank
        model = ToyEncoder().double()
        B = 6
        crit = NTXentRestated("cpu", B, 0.1, True)
        stepper = DataParallelStep(model, B, 0.1, True, global_negatives=global_neg, kern=TorchKernels(),
                                   local_criterion=crit)
        g = torch.Generator().manual_seed(7)
        data = torch.randn(world, 2, B, 12, generator=g, dtype=torch.float64)   # same on all ranks; each takes its slice
        loss = stepper.loss(data[rank, 0], data[rank, 1])
        loss.backward()
        stepper.allreduce_gradients()
        total = stepper.global_loss(loss)
        torch.save({"loss": total, "share": loss.detach(), "grads": [p.grad.clone() for p in model.parameters()],
                    "params": [p.detach().clone() for p in model.parameters()], "data": data},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _run(global_neg, world=2):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), global_neg, d), nprocs=world, join=True)
        return [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(world)]


def _single_process_reference(res, global_neg):
    from oracle.nt_xent import ntxent_closed_form
    world = len(res)
    model = ToyEncoder().double()
    with torch.no_grad():
        for p, q in zip(model.parameters(), res[0]["params"]):
            p.copy_(q)
    data = res[0]["data"]
    norm = torch.nn.functional.normalize
    if global_neg:
        zis = torch.cat([norm(model(data[r, 0])[1], dim=1) for r in range(world)])
        zjs = torch.cat([norm(model(data[r, 1])[1], dim=1) for r in range(world)])
        loss = ntxent_closed_form(zis, zjs, 0.1, True)
    else:
        loss = sum(ntxent_closed_form(norm(model(data[r, 0])[1], dim=1), norm(model(data[r, 1])[1], dim=1), 0.1, True)
                   for r in range(world)) / world
    loss.backward()
    return loss.detach(), [p.grad for p in model.parameters()]


def test_replicas_start_identical_and_end_with_identical_gradients():
    res = _run(True)
    for a, b in zip(res[0]["params"], res[1]["params"]):
        assert torch.equal(a, b)                              # rank 0's weights were broadcast
    for a, b in zip(res[0]["grads"], res[1]["grads"]):
        assert torch.equal(a, b)


def test_global_negatives_match_single_process_closed_form():
    res = _run(True)
    loss, grads = _single_process_reference(res, True)
    assert abs(float(res[0]["loss"]) - float(loss)) < 1e-10 * abs(float(loss))
    assert abs(float(res[0]["share"]) + float(res[1]["share"]) - float(loss)) < 1e-10 * abs(float(loss))
    for g, ref in zip(res[0]["grads"], grads):
        assert float((g - ref).norm() / ref.norm()) < 1e-9


def test_local_negatives_average_gradients():
    res = _run(False)
    loss, grads = _single_process_reference(res, False)
    assert abs(float(res[0]["loss"]) - float(loss)) < 1e-10 * abs(float(loss))
    for g, ref in zip(res[0]["grads"], grads):
        assert float((g - ref).norm() / ref.norm()) < 1e-9
