"""Shared helpers for the parity tests."""
import numpy as np
import torch


def tf32_round(x):
    """cvt.rna.tf32.f32 emulated on any device: round to nearest (ties away) to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def rel_err(a, b):
    """||a-b|| / ||b|| in fp64 (norm-relative error)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def max_rel(a, b):
    """max|a-b| / max|b|."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def sync_oracle_from(model, oracle):
    """Copies the state_dict of the CUDA drop-in into the CPU oracle (shared weights)."""
    oracle.load_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    return oracle
