"""Caller-side pieces of the hot path (molclr.py:55-67) on the same kernels."""
import torch

from . import ops


class _Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, eps):
        z = z.contiguous()
        y, inv = ops.l2_normalize_fwd(z, eps)
        ctx.save_for_backward(y, inv)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, gy):
        y, inv = ctx.saved_tensors
        return ops.l2_normalize_bwd(gy.contiguous(), y, inv, ctx.eps), None


def normalize(z, dim=1, eps=1e-12):
    """``F.normalize(z, dim=1)`` as used in molclr.py:63-64."""
    if dim not in (1, -1) or z.dim() != 2:
        raise NotImplementedError("molclr_b200.normalize supports 2-D inputs with dim=1")
    return _Normalize.apply(z, eps)


def pretrain_loss(model, criterion, xis, xjs):
    """``MolCLR._step`` (molclr.py:55-67): two encoder passes (view i, then view j), L2 normalisation, NT-Xent.  Models that offer
    ``forward_pair`` run the two passes as one autograd node (same values; fewer gradient-accumulation launches)."""
    if hasattr(model, "forward_pair"):
        (_ris, zis), (_rjs, zjs) = model.forward_pair(xis, xjs)
    else:
        _ris, zis = model(xis)
        _rjs, zjs = model(xjs)
    zis = normalize(zis, dim=1)
    zjs = normalize(zjs, dim=1)
    return criterion(zis, zjs)
