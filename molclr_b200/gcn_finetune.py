"""GCN (fine-tune variant): drop-in for ``models/gcn_finetune.py`` -- the GCN encoder of ``gcn.py`` (un-normalised GCNConv as
the reference computes it), ``feat_lin`` and ``pred_lin`` = Linear(feat, feat/2) -> Softplus -> Linear(feat/2, 2 | 1)
(gcn_finetune.py:133-144); ``forward(data)`` returns ``(h, pred_lin(h))`` (:146-163).  The task loss is the caller's
(finetune.py:70-77).
"""
import torch
from torch import nn

from . import ops
from .gcn import GCNConv, _gcn_encoder_backward, _gcn_encoder_forward, _gcn_precision
from .ginet import _EncoderBase, _RoundedWeights, num_atom_type, num_chirality_tag
from .ginet_finetune import finetune_head_backward, finetune_head_forward
from .graph import get_plan


class GCN(_EncoderBase):
    """gcn_finetune.py:94-163.

    Args:
        task (str): 'classification' (2 logits) | 'regression' (1 output)
        num_layer (>= 2), emb_dim, feat_dim, drop_ratio, pool ('mean' | 'add' | 'max'): as the pre-training GCN
    """

    def __init__(self, task="classification", num_layer=5, emb_dim=300, feat_dim=256, drop_ratio=0, pool="mean"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio, self.task = num_layer, emb_dim, feat_dim, drop_ratio, task
        if feat_dim % 8 != 0 or emb_dim % 4 != 0:
            # feat_dim // 2 is the width of the head's hidden activations -- tensor-core operands, whose widths are multiples of 4
            raise ValueError(f"molclr_b200: feat_dim must be a multiple of 8 and emb_dim a multiple of 4, got feat_dim={feat_dim}, emb_dim={emb_dim}")
        if self.num_layer < 2:
            raise ValueError("Number of GNN layers must be greater than 1.")          # gcn_finetune.py:103-104
        if pool not in ("mean", "add", "max"):
            raise ValueError("Not defined pooling!")                                  # gcn_finetune.py:128-129
        self.pool_name = pool
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GCNConv(emb_dim, aggr="add") for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        if task == "classification":
            out_dim = 2
        elif task == "regression":
            out_dim = 1
        else:          # the reference leaves pred_lin undefined (gcn_finetune.py:133-144) and fails at forward
            raise AttributeError("'GCN' object has no attribute 'pred_lin'")
        self.pred_lin = nn.Sequential(nn.Linear(feat_dim, feat_dim // 2), nn.Softplus(), nn.Linear(feat_dim // 2, out_dim))
        self._rounded = _RoundedWeights()

    def _params(self):
        ps = [self.x_embedding1.weight, self.x_embedding2.weight]
        for g, bn in zip(self.gnns, self.batch_norms):
            ps += [g.weight, g.bias, g.edge_embedding1.weight, g.edge_embedding2.weight, bn.weight, bn.bias]
        ps += [self.feat_lin.weight, self.feat_lin.bias, self.pred_lin[0].weight, self.pred_lin[0].bias,
               self.pred_lin[2].weight, self.pred_lin[2].bias]
        return ps

    def forward(self, data):
        self._check_input(data)
        plan = get_plan(data)
        return _GCNFinetuneFunction.apply(self, plan, *self._params())

    def load_my_state_dict(self, state_dict):
        """gcn_finetune.py:165-173: copy the entries whose names exist here (pre-trained encoder -> fine-tune model)."""
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name not in own_state:
                continue
            if isinstance(param, nn.parameter.Parameter):
                param = param.data
            own_state[name].copy_(param)


class _GCNFinetuneFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, m, plan, *params):
        comp = _gcn_precision(m)
        training, pool_mode = m.training, ops.POOL_MODES[m.pool_name]
        m._refresh_weights(comp)
        p, p_lo, saved = _gcn_encoder_forward(m, plan, comp, training, pool_mode)
        mode = ops.ACT_MODES["softplus"]
        h, pred, head_saved, Wf = finetune_head_forward(m, p, p_lo, [m.pred_lin[0], m.pred_lin[2]], mode, comp)
        ctx.m, ctx.plan, ctx.saved, ctx.p, ctx.head_saved, ctx.Wf = m, plan, saved, p, head_saved, Wf
        ctx.training, ctx.pool_mode, ctx.mode = training, pool_mode, mode
        return h, pred

    @staticmethod
    def backward(ctx, g_h, g_pred):
        m, plan, p = ctx.m, ctx.plan, ctx.p
        g_p, dWf, dbf, head_grads = finetune_head_backward(m, p, ctx.head_saved, ctx.Wf, g_h, g_pred, ctx.mode)
        grads = _gcn_encoder_backward(m, plan, ctx.saved, g_p, ctx.training, ctx.pool_mode)
        ctx.saved = None
        return (None, None, *grads, dWf, dbf, *head_grads)
