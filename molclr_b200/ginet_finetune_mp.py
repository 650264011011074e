"""GINet (motif-level fine-tune variant): drop-in for ``models/ginet_finetune_mp.py`` -- the GIN-E encoder and ``feat_lin`` of
the fine-tune model, then a motif branch: ``hp = motif_lin(GlobalAttention([motif_embedding[clique_idx]; h], mol_idx))`` and
``pred_head(cat(h, hp))`` (ginet_finetune_mp.py:141-163).  ``forward(data, mol_idx, clique_idx)`` returns ``(cat(h, hp), pred)``.

The encoder and ``feat_lin`` (all of the work that scales with atoms and bonds) run on the kernels of ``ginet.py`` through
``_EncoderFeatFunction``.  The motif branch works on ``G + #cliques`` rows of width ``feat_dim`` -- a few hundred kFLOP per
molecule against 0.5 GFLOP in the encoder -- and is written with torch tensor operations (``motif_head``); it is the part of
this file the CPU tests can pin against the reference-generated golden vectors.
"""
import torch
from torch import nn

from . import ops
from .ginet import (GINEConv, _EncoderBase, _RoundedWeights, _check_precision, _encoder_backward, _encoder_forward, _lo,
                    num_atom_type, num_chirality_tag)
from .graph import get_plan


def group_softmax(src, index, num_groups):
    """``torch_geometric.utils.softmax`` (1.6.3) over dim 0: ``exp(src - max_g) / (sum_g exp + 1e-16)`` within the groups of ``index``."""
    idx = index.view(-1, 1).expand_as(src)
    mx = torch.full((num_groups, src.shape[1]), float("-inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    out = (src - mx[index]).exp()
    den = torch.zeros(num_groups, src.shape[1], dtype=src.dtype, device=src.device).scatter_add_(0, idx, out)
    return out / (den[index] + 1e-16)


class GlobalAttention(nn.Module):
    """``torch_geometric.nn.GlobalAttention`` (1.6.3) with the reference's attribute names (``gate_nn``, ``nn``), so that
    ``state_dict`` keys match: ``r_g = sum_{n in g} softmax_g(gate_nn(x_n)) * nn(x_n)``; the number of groups defaults to
    ``batch[-1] + 1``."""

    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn

    def forward(self, x, batch, size=None):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        size = int(batch[-1]) + 1 if size is None else size
        gate = self.gate_nn(x).view(-1, 1)
        x = self.nn(x) if self.nn is not None else x
        if gate.dim() != x.dim() or gate.size(0) != x.size(0):
            raise AssertionError("GlobalAttention: gate and features disagree in shape")
        gate = group_softmax(gate, batch, size)
        return torch.zeros(size, x.shape[1], dtype=x.dtype, device=x.device).scatter_add_(0, batch.view(-1, 1).expand_as(x), gate * x)


class GINet(_EncoderBase):
    """ginet_finetune_mp.py:52-163.

    Args:
        num_motifs (int): rows of the motif embedding table
        task (str): 'classification' (2 logits) | 'regression' (1 output)
        num_layer, emb_dim, feat_dim, drop_ratio, pool: as the pre-training model
        pred_n_layer (int): number of hidden layers of the prediction head (>= 1)
        pred_act (str): 'softplus' | 'relu'
    """

    def __init__(self, num_motifs, task="classification", num_layer=5, emb_dim=300, feat_dim=512, drop_ratio=0, pool="mean",
                 pred_n_layer=2, pred_act="softplus"):
        super().__init__()
        self.num_motifs, self.num_layer, self.emb_dim, self.feat_dim = num_motifs, num_layer, emb_dim, feat_dim
        if feat_dim % 8 != 0 or emb_dim % 4 != 0:
            # feat_dim // 2 is the width of the head's hidden activations -- tensor-core operands, whose widths are multiples of 4
            raise ValueError(f"molclr_b200: feat_dim must be a multiple of 8 and emb_dim a multiple of 4, got feat_dim={feat_dim}, emb_dim={emb_dim}")
        self.drop_ratio, self.task = drop_ratio, task
        self.pool_name = pool
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.motif_embedding = nn.Embedding(num_motifs, feat_dim)                      # ginet_finetune_mp.py:79
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        if task == "classification":
            out_dim = 2
        elif task == "regression":
            out_dim = 1
        else:                              # the reference hits an unbound `out_dim` here (ginet_finetune_mp.py:99-102)
            raise UnboundLocalError("local variable 'out_dim' referenced before assignment")
        self.motif_lin = nn.Linear(feat_dim, feat_dim)                                 # :104-105
        nn.init.xavier_uniform_(self.motif_lin.weight.data)
        self.motif_pool = GlobalAttention(gate_nn=nn.Sequential(nn.Linear(feat_dim, 1)))   # :107
        self.pred_n_layer = max(1, pred_n_layer)
        if pred_act == "relu":
            act = lambda: nn.ReLU(inplace=True)
        elif pred_act == "softplus":
            act = nn.Softplus
        else:
            raise ValueError("Undefined activation function")                          # :132-133
        head = [nn.Linear(2 * feat_dim, feat_dim // 2), act()]
        for _ in range(self.pred_n_layer - 1):
            head.extend([nn.Linear(feat_dim // 2, feat_dim // 2), act()])
        head.append(nn.Linear(feat_dim // 2, out_dim))
        self.pred_head = nn.Sequential(*head)
        self._rounded = _RoundedWeights()

    def init_motif_emb(self, init):
        """ginet_finetune_mp.py:138-140."""
        with torch.no_grad():
            self.motif_embedding.weight.data = nn.Parameter(init)

    def _params(self):
        ps = [self.x_embedding1.weight, self.x_embedding2.weight]
        for g, bn in zip(self.gnns, self.batch_norms):
            ps += [g.mlp[0].weight, g.mlp[0].bias, g.mlp[2].weight, g.mlp[2].bias,
                   g.edge_embedding1.weight, g.edge_embedding2.weight, bn.weight, bn.bias]
        ps += [self.feat_lin.weight, self.feat_lin.bias]
        return ps

    def motif_head(self, h, mol_idx, clique_idx):
        """ginet_finetune_mp.py:157-163 on the molecule features ``h = feat_lin(pool(...))``: returns ``(cat(h, hp), pred)``."""
        hp = self.motif_embedding(clique_idx)
        hp = torch.cat((hp, h), dim=0)
        hp = self.motif_pool(hp, mol_idx)
        hp = self.motif_lin(hp)
        h = torch.cat((h, hp), dim=1)
        return h, self.pred_head(h)

    def forward(self, data, mol_idx, clique_idx):
        self._check_input(data)
        plan = get_plan(data)
        h = _EncoderFeatFunction.apply(self, plan, *self._params())
        return self.motif_head(h, mol_idx, clique_idx)

    def load_my_state_dict(self, state_dict):
        """ginet_finetune_mp.py:165-174: copy the entries whose names exist here (pre-trained encoder -> fine-tune model)."""
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name not in own_state:
                continue
            if isinstance(param, nn.parameter.Parameter):
                param = param.data
            own_state[name].copy_(param)


class _EncoderFeatFunction(torch.autograd.Function):
    """Encoder (ginet.py kernels) followed by ``feat_lin`` on the tcgen05 GEMM; returns ``h`` [G, feat_dim]."""

    @staticmethod
    def forward(ctx, m, plan, *params):
        comp = _check_precision(m)
        training, pool_mode = m.training, ops.POOL_MODES[m.pool_name]
        m._refresh_weights(comp)
        p, p_lo, layers = _encoder_forward(m, plan, comp, training, pool_mode)
        G, D, Fd = p.shape[0], m.emb_dim, m.feat_dim
        Wf, Wf_lo = m._rounded.get(m.feat_lin.weight)
        h = torch.empty(G, Fd, device=p.device)
        ops.gemm(p, Wf, G, Fd, D, A_lo=p_lo, B_lo=_lo(Wf_lo, comp), out=h, bias=m.feat_lin.bias.detach())
        ctx.m, ctx.plan, ctx.layers, ctx.p, ctx.Wf = m, plan, layers, p, Wf
        ctx.training, ctx.pool_mode = training, pool_mode
        return h

    @staticmethod
    def backward(ctx, g_h):
        m, plan, p, Wf = ctx.m, ctx.plan, ctx.p, ctx.Wf
        G, D, Fd = p.shape[0], m.emb_dim, m.feat_dim
        g_h = g_h.contiguous()
        g_r = ops.round_tf32(g_h)
        dWf = ops.gemm_dw(g_r, p, ordered=m.deterministic)                                   # [feat_dim, emb_dim]
        dbf = ops.colsum(g_h)
        g_p = torch.empty(G, D, device=p.device)
        ops.gemm(g_r, Wf, G, D, Fd, b_mn=True, out=g_p)
        grads = _encoder_backward(m, plan, ctx.layers, g_p, ctx.training, ctx.pool_mode)
        ctx.layers = None
        return (None, None, *grads, dWf, dbf)
