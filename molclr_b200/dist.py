"""Data-parallel MolCLR pre-training over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Molecules are independent
through the whole encoder, so every rank runs the encoder kernels on its own batch of B pairs with its own
BatchNorm statistics; the path has exactly two exchange steps:

1. **NT-Xent with global negatives** -- the (cosine-normalised, tf32-rounded) projections of all ranks
   are all-gathered into the candidate matrix ``cols [2][W*B][C]`` (rows ``[all zjs ; all zis]``, the
   reference's ordering of nt_xent.py:48 extended over ranks), each rank computes the log-sum-exp of ITS
   2B anchor rows against all 2*W*B candidates, the 2B row log-sum-exps are all-gathered, and the backward
   uses the symmetry of the loss (``molclr_ntxent_bwd``: the column-softmax term of a local row is
   recomputed from the same similarity tile with the other row's gathered log-sum-exp), so the gradient of
   the GLOBAL mean loss w.r.t. the local projections needs no reduce-scatter of a [2*W*B, C] gradient.
2. **Gradient all-reduce** -- one flat fp32 buffer (2.4 M elements for GIN-5/300/512), summed over ranks.

Loss / gradient scale: the objective is the mean over all 2*W*B anchors.  With global negatives every rank
holds the exact partial derivative of that objective through its own projections, so parameter gradients
are SUMMED over ranks; with local negatives each rank's loss is the mean over its own 2B anchors and the
gradients are AVERAGED (the usual DDP convention).  ``loss()`` returns this rank's share; the global value
is ``global_loss(share)`` (one scalar all-reduce, for logging only).

The collective logic here is device-agnostic: ``kern`` supplies the five kernels it calls.  The product
default is the CUDA kernels (``molclr_b200.ops``); the CPU tests (gloo, world_size 2) inject a plain-torch
stand-in so that the gather ordering, row offsets and gradient scaling are checked without a GPU.
"""
import torch
import torch.distributed as dist


class CudaKernels:
    """The sm_100a kernels behind the C ABI (no fallback: raises without a CUDA device / the library)."""

    def __init__(self):
        from . import functional, ops
        self._ops = ops
        self.normalize = functional.normalize
        self.l2_normalize_fwd = ops.l2_normalize_fwd
        self.l2_normalize_bwd = ops.l2_normalize_bwd
        self.round_tf32 = ops.round_tf32
        self.ntxent_fwd = ops.ntxent_fwd
        self.ntxent_bwd = ops.ntxent_bwd


def _gather_rows(dst, src, group):
    """dst [W*R, ...] <- concatenation over ranks of src [R, ...] (rank order)."""
    dist.all_gather_into_tensor(dst, src.contiguous(), group=group)


class _GlobalNTXentFunction(torch.autograd.Function):
    """NT-Xent (nt_xent.py:47-65) of this rank's 2B anchors against the candidates of all ranks."""

    @staticmethod
    def forward(ctx, zis, zjs, temperature, use_cosine, group, kern):
        W, r = dist.get_world_size(group), dist.get_rank(group)
        B, C = zis.shape
        local = torch.cat([zjs, zis], dim=0).contiguous()                    # nt_xent.py:48 (zjs FIRST)
        if use_cosine:                                                       # CosineSimilarity eps 1e-8 (nt_xent.py:19,44)
            local_n, inv = kern.l2_normalize_fwd(local, 1e-8)
        else:
            local_n, inv = local, None
        local_r = kern.round_tf32(local_n)
        Bg = W * B
        cols = torch.empty(2 * Bg, C, dtype=local_r.dtype, device=local_r.device)
        _gather_rows(cols[:Bg], local_r[:B], group)                          # all zjs
        _gather_rows(cols[Bg:], local_r[B:], group)                          # all zis
        # local rows [zjs; zis] are candidates r*B.. and Bg + r*B.. of the global ordering
        share, lse, _pos = kern.ntxent_fwd(local_r, cols, r * B, 1.0 / temperature, Bg + r * B, unit_rows=use_cosine)
        col_lse = torch.empty(2 * Bg, dtype=lse.dtype, device=lse.device)
        _gather_rows(col_lse[:Bg], lse[:B], group)
        _gather_rows(col_lse[Bg:], lse[B:], group)
        ctx.save_for_backward(local_n, inv, local_r, cols, lse, col_lse)
        ctx.meta = (B, Bg, r, temperature, use_cosine, kern)
        return share[0]

    @staticmethod
    def backward(ctx, g_loss):
        local_n, inv, local_r, cols, lse, col_lse = ctx.saved_tensors
        B, Bg, r, temperature, use_cosine, kern = ctx.meta
        g = kern.ntxent_bwd(local_r, cols, r * B, 1.0 / temperature, lse, col_lse, Bg + r * B, unit_rows=use_cosine) * g_loss
        if use_cosine:
            g = kern.l2_normalize_bwd(g.contiguous(), local_n, inv, 1e-8)
        return g[B:], g[:B], None, None, None, None


def global_ntxent(zis, zjs, temperature, use_cosine_similarity=True, group=None, kern=None):
    """This rank's share of the NT-Xent loss over the global batch; the sum over ranks is the loss of
    ``NTXentLoss(batch_size=W*B)`` applied to the rank-concatenated projections."""
    if zis.shape != zjs.shape or zis.dim() != 2:
        raise RuntimeError(f"global_ntxent: zis {tuple(zis.shape)} and zjs {tuple(zjs.shape)} must be equal 2-D shapes")
    return _GlobalNTXentFunction.apply(zis, zjs, float(temperature), bool(use_cosine_similarity), group, kern or CudaKernels())


class DataParallelStep:
    """The loop body of molclr.py:109-127 for one rank of a data-parallel job.

        stepper = DataParallelStep(model, batch_size, temperature, use_cosine_similarity)
        optimizer.zero_grad(); loss = stepper.loss(xis, xjs); loss.backward()
        stepper.allreduce_gradients(); optimizer.step()

    Construction broadcasts rank 0's parameters and buffers so that all replicas start identical.
    """

    def __init__(self, model, batch_size, temperature, use_cosine_similarity, global_negatives=True, group=None, kern=None,
                 local_criterion=None):
        if not dist.is_initialized():
            raise RuntimeError("DataParallelStep: torch.distributed is not initialised")
        self.model, self.batch_size, self.temperature = model, batch_size, float(temperature)
        self.use_cosine, self.global_negatives, self.group = bool(use_cosine_similarity), bool(global_negatives), group
        self.kern = kern or CudaKernels()
        self.world = dist.get_world_size(group)
        self._criterion = local_criterion
        if not self.global_negatives and self._criterion is None:
            from .nt_xent import NTXentLoss
            self._criterion = NTXentLoss(None, batch_size, temperature, use_cosine_similarity)
        self._flat, self._views = None, None
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t, src=0, group=group)

    # gradients are summed (global negatives) or averaged (local negatives) over ranks -- see the module docstring
    @property
    def grad_scale(self):
        return 1.0 if self.global_negatives else 1.0 / self.world

    def loss(self, xis, xjs):
        _ris, zis = self.model(xis)                         # molclr.py:57
        _rjs, zjs = self.model(xjs)                         # molclr.py:60
        zis = self.kern.normalize(zis, dim=1)               # molclr.py:63-64
        zjs = self.kern.normalize(zjs, dim=1)
        if zis.shape[0] != self.batch_size:
            raise RuntimeError(f"DataParallelStep: got {zis.shape[0]} graphs but batch_size={self.batch_size} "
                               "(the reference requires drop_last=True, dataset.py:180)")
        if self.global_negatives:
            return global_ntxent(zis, zjs, self.temperature, self.use_cosine, self.group, self.kern)
        return self._criterion(zis, zjs)

    def global_loss(self, share):
        """The logged loss value: mean over all anchors of all ranks (one scalar all-reduce)."""
        t = share.detach().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t * self.grad_scale

    def allreduce_gradients(self):
        """One all-reduce of a flat fp32 buffer holding every parameter gradient; ``p.grad`` then aliases it."""
        params = [p for p in self.model.parameters() if p.requires_grad]
        if self._flat is None or self._flat.device != params[0].device:
            total = sum(p.numel() for p in params)
            self._flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
            self._views, off = [], 0
            for p in params:
                self._views.append(self._flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        have = [(v, p.grad) for v, p in zip(self._views, params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self._views, params):
            if p.grad is None:
                v.zero_()
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.grad_scale != 1.0:
            self._flat.mul_(self.grad_scale)
        for v, p in zip(self._views, params):
            p.grad = v
