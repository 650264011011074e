"""NTXentLoss: drop-in for ``utils/nt_xent.py`` (same constructor and ``forward(zis, zjs)``), computed by
the fused similarity-GEMM + masked log-sum-exp kernels; the 2N x 2N matrix is never materialised.

With ``torch.distributed`` initialised and ``global_negatives=True`` the projections of all ranks are
all-gathered (NCCL) so every anchor sees the negatives of the whole global batch.
"""
import torch

from . import ops


class _NTXentFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, zis, zjs, temperature, use_cosine, group):
        n = zis.shape[0]
        # rep = cat([zjs, zis]) (nt_xent.py:48, zjs FIRST), rows divided by max(|row|, 1e-8) for the cosine similarity
        # (torch.nn.CosineSimilarity(dim=-1), nt_xent.py:19,44), and the tf32-rounded operand copy: one kernel
        rep_n, rep_r, inv = ops.ntxent_rows_fwd(zjs.contiguous(), zis.contiguous(), 1e-8, use_cosine)
        loss, row_lse, _row_pos = ops.ntxent_fwd(rep_r, rep_r, 0, 1.0 / temperature, unit_rows=use_cosine)
        ctx.save_for_backward(rep_n, inv, rep_r, row_lse)
        ctx.n, ctx.temperature, ctx.use_cosine = n, temperature, use_cosine
        return loss[0]

    @staticmethod
    def backward(ctx, g_loss):
        rep_n, inv, rep_r, row_lse = ctx.saved_tensors
        g = ops.ntxent_bwd(rep_r, rep_r, 0, 1.0 / ctx.temperature, row_lse, row_lse, unit_rows=ctx.use_cosine)
        if ctx.use_cosine:
            g = ops.l2_normalize_bwd(g, rep_n, inv, 1e-8, gscale=g_loss.contiguous())      # (* g_loss folded in: no eager multiply)
        else:
            g = g * g_loss
        n = ctx.n
        return g[n:], g[:n], None, None, None


class NTXentLoss(torch.nn.Module):
    """nt_xent.py:5-65.  ``loss = CrossEntropy(sum)([pos | negatives]/T, 0) / 2N`` over the rows
    ``[zjs; zis]`` with the cosine (or dot, nt_xent.py:32-38) similarity."""

    def __init__(self, device, batch_size, temperature, use_cosine_similarity):
        super().__init__()
        self.batch_size = batch_size
        self.temperature = temperature
        self.device = device
        self.use_cosine_similarity = bool(use_cosine_similarity)

    def forward(self, zis, zjs):
        if zis.shape != zjs.shape or zis.dim() != 2:
            raise RuntimeError(f"NTXentLoss: zis {tuple(zis.shape)} and zjs {tuple(zjs.shape)} must be equal 2-D shapes")
        if zis.shape[0] != self.batch_size:
            # the reference bakes batch_size into its mask and fails in .view() (nt_xent.py:55-57)
            raise RuntimeError(f"NTXentLoss: got {zis.shape[0]} rows but batch_size={self.batch_size} "
                               "(the reference requires drop_last=True, dataset.py:180)")
        return _NTXentFunction.apply(zis, zjs, float(self.temperature), self.use_cosine_similarity, None)
