"""Oracle: the caller contract of the hot path.  TEST INFRASTRUCTURE -- see oracle/__init__.py."""
import torch
import torch.nn.functional as F


def pretrain_loss(model, criterion, xis, xjs):
    """MolCLR._step, molclr.py:55-67: two SEPARATE encoder passes, F.normalize, NT-Xent."""
    _ris, zis = model(xis)
    _rjs, zjs = model(xjs)
    zis = F.normalize(zis, dim=1)
    zjs = F.normalize(zjs, dim=1)
    return criterion(zis, zjs)


def train_step(model, criterion, optimizer, xis, xjs):
    """One iteration of the loop body molclr.py:109-127 (without logging)."""
    optimizer.zero_grad()
    loss = pretrain_loss(model, criterion, xis, xjs)
    loss.backward()
    optimizer.step()
    return loss
