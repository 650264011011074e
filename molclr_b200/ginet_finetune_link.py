"""GINet (label-conditioned fine-tune variant): drop-in for ``models/ginet_finetune_link.py`` -- the GIN-E encoder and ``feat_lin``
of the fine-tune model; the molecule feature is concatenated with ``label_lin(label_embedding[c])`` for c = 0 and c = 1 and the
same ``pred_head`` scores both (ginet_finetune_link.py:134-158).  ``forward(data, device)`` returns ``(h, cat(score_0, score_1))``.

Encoder and ``feat_lin`` run on the kernels of ``ginet.py`` (``ginet_finetune_mp._EncoderFeatFunction``); the label branch works
on G rows of width ``feat_dim`` and is written with torch tensor operations (``label_head``), which the CPU tests pin against the
reference-generated golden vectors.
"""
import torch
from torch import nn

from .ginet import GINEConv, _EncoderBase, _RoundedWeights, num_atom_type, num_chirality_tag
from .ginet_finetune_mp import _EncoderFeatFunction
from .graph import get_plan


class GINet(_EncoderBase):
    """ginet_finetune_link.py:52-172.  Arguments as ``ginet_finetune.GINet``; the head always has ONE output (:98)."""

    def __init__(self, task="classification", num_layer=5, emb_dim=300, feat_dim=512, drop_ratio=0, pool="mean",
                 pred_n_layer=2, pred_act="softplus"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio, self.task = num_layer, emb_dim, feat_dim, drop_ratio, task
        self.pool_name = pool
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.label_embedding = nn.Embedding(2, feat_dim)                               # ginet_finetune_link.py:78
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        self.label_lin = nn.Linear(feat_dim, feat_dim)                                 # :100-101
        nn.init.xavier_uniform_(self.label_lin.weight)
        self.pred_n_layer = max(1, pred_n_layer)
        if pred_act == "relu":
            act = lambda: nn.ReLU(inplace=True)
        elif pred_act == "softplus":
            act = nn.Softplus
        else:
            raise ValueError("Undefined activation function")                          # :126-127
        head = [nn.Linear(2 * feat_dim, feat_dim // 2), act()]
        for _ in range(self.pred_n_layer - 1):
            head.extend([nn.Linear(feat_dim // 2, feat_dim // 2), act()])
        head.append(nn.Linear(feat_dim // 2, 1))
        self.pred_head = nn.Sequential(*head)
        self._rounded = _RoundedWeights()

    def init_label_emb(self, init):
        """ginet_finetune_link.py:130-132."""
        with torch.no_grad():
            self.label_embedding.weight = nn.Parameter(init)

    def _params(self):
        ps = [self.x_embedding1.weight, self.x_embedding2.weight]
        for g, bn in zip(self.gnns, self.batch_norms):
            ps += [g.mlp[0].weight, g.mlp[0].bias, g.mlp[2].weight, g.mlp[2].bias,
                   g.edge_embedding1.weight, g.edge_embedding2.weight, bn.weight, bn.bias]
        ps += [self.feat_lin.weight, self.feat_lin.bias]
        return ps

    def label_head(self, h):
        """ginet_finetune_link.py:151-158 on the molecule features ``h``: returns ``cat(score_0, score_1)`` [G, 2]."""
        G = h.shape[0]
        h1 = self.label_lin(self.label_embedding(torch.zeros(G, dtype=torch.long, device=h.device)))
        h2 = self.label_lin(self.label_embedding(torch.ones(G, dtype=torch.long, device=h.device)))
        return torch.cat((self.pred_head(torch.cat((h, h1), dim=1)), self.pred_head(torch.cat((h, h2), dim=1))), dim=1)

    def forward(self, data, device=None):
        """``device`` is accepted for signature compatibility (the reference builds its label indices on it, :151,153); the label
        branch runs on the device of the features."""
        self._check_input(data)
        plan = get_plan(data)
        h = _EncoderFeatFunction.apply(self, plan, *self._params())
        return h, self.label_head(h)

    def load_my_state_dict(self, state_dict):
        """ginet_finetune_link.py:162-171: copy the entries whose names exist here (pre-trained encoder -> fine-tune model)."""
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name not in own_state:
                continue
            if isinstance(param, nn.parameter.Parameter):
                param = param.data
            own_state[name].copy_(param)
