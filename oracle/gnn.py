"""Oracle: the reference's GNN encoders restated with the ATen primitives PyG dispatches to.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PARITY UNPINNED for this file: the
reference has no tests and torch-geometric 1.6.3 / torch-scatter 2.0.6 are not installable
here, so the semantics of those libraries are restated from their published behaviour:

* ``add_self_loops(edge_index, num_nodes=N)``  (PyG 1.6.3 ``utils/loop.py``): appends
  ``arange(N)`` loops at the END of ``edge_index``.
* ``MessagePassing()`` defaults ``aggr='add'``, ``flow='source_to_target'``,
  ``node_dim=-2``: ``x_j = x.index_select(0, edge_index[0])``; aggregation is
  ``torch_scatter.scatter(msg, edge_index[1], dim=0, dim_size=N, reduce='sum')`` which is
  ``zeros(N, D).scatter_add_(0, index.expand_as(msg), msg)``.
* ``global_mean_pool(x, batch)`` = ``scatter(x, batch, dim=0, dim_size=batch.max()+1,
  reduce='mean')`` = sum / ``count.clamp(min=1)``; ``global_add_pool`` is the sum,
  ``global_max_pool`` the per-segment max.

Module/parameter names equal the reference's so that ``state_dict()`` keys match
(SURVEY.md section 8b).
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

num_atom_type = 119      # ginet_molclr.py:9
num_chirality_tag = 3    # ginet_molclr.py:10
num_bond_type = 5        # ginet_molclr.py:12  (4 = self-loop bond type)
num_bond_direction = 3   # ginet_molclr.py:13


# ----------------------------------------------------------------------------- PyG restatements
def add_self_loops(edge_index, num_nodes):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1)


def scatter_sum(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_add_(0, idx, src)


def propagate_add(edge_index, x, edge_attr):
    """MessagePassing.propagate with message = x_j + edge_attr, aggr='add'."""
    x_j = x.index_select(0, edge_index[0])
    msg = x_j + edge_attr                       # ginet_molclr.py:43-44 / gcn_molclr.py:86-88
    return scatter_sum(msg, edge_index[1], x.size(0))


def num_graphs_of(batch):
    return int(batch.max().item()) + 1 if batch.numel() else 0


def global_add_pool(x, batch, size=None):
    size = num_graphs_of(batch) if size is None else size
    return scatter_sum(x, batch, size)


def global_mean_pool(x, batch, size=None):
    size = num_graphs_of(batch) if size is None else size
    out = scatter_sum(x, batch, size)
    count = scatter_sum(torch.ones(batch.size(0), dtype=x.dtype, device=x.device), batch, size)
    count.clamp_(1)
    return out / count.view(-1, 1)


def global_max_pool(x, batch, size=None):
    size = num_graphs_of(batch) if size is None else size
    out = torch.full((size, x.size(1)), float("-inf"), dtype=x.dtype, device=x.device)
    idx = batch.view(-1, 1).expand_as(x)
    out = out.scatter_reduce(0, idx, x, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out), torch.zeros_like(out), out)


_POOLS = {"mean": global_mean_pool, "max": global_max_pool, "add": global_add_pool}


def _dropout(model, layer, h):
    """F.dropout(h, drop_ratio, training) -- or, when the test sets ``model.dropout_masks`` (a list of per-layer
    [N, D] tensors holding 0 or 1/(1-p)), multiplication by that explicit mask, so that the CUDA kernels' counter-based
    masks can be checked value by value."""
    masks = getattr(model, "dropout_masks", None)
    if masks is not None and model.training:
        return h * masks[layer].to(h.dtype)
    return F.dropout(h, model.drop_ratio, training=model.training)


def _self_loop_attr(edge_attr, num_nodes):
    # ginet_molclr.py:34-37 / gcn_molclr.py:67-70: rows [4, 0] appended after the real edges
    sl = torch.zeros(num_nodes, 2)
    sl[:, 0] = 4
    sl = sl.to(edge_attr.device).to(edge_attr.dtype)
    return torch.cat((edge_attr, sl), dim=0)


# ----------------------------------------------------------------------------- GIN
class GINEConv(nn.Module):
    """ginet_molclr.py:16-47."""

    def __init__(self, emb_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(emb_dim, 2 * emb_dim), nn.ReLU(), nn.Linear(2 * emb_dim, emb_dim))
        self.edge_embedding1 = nn.Embedding(num_bond_type, emb_dim)
        self.edge_embedding2 = nn.Embedding(num_bond_direction, emb_dim)
        nn.init.xavier_uniform_(self.edge_embedding1.weight.data)
        nn.init.xavier_uniform_(self.edge_embedding2.weight.data)

    def aggregate(self, x, edge_index, edge_attr):
        """Everything before ``update`` (ginet_molclr.py:29-44): the neighbour aggregate."""
        edge_index = add_self_loops(edge_index, x.size(0))
        edge_attr = _self_loop_attr(edge_attr, x.size(0))
        edge_embeddings = self.edge_embedding1(edge_attr[:, 0]) + self.edge_embedding2(edge_attr[:, 1])
        return propagate_add(edge_index, x, edge_embeddings)

    def forward(self, x, edge_index, edge_attr):
        return self.mlp(self.aggregate(x, edge_index, edge_attr))     # update(): ginet_molclr.py:46-47


class GINet(nn.Module):
    """ginet_molclr.py:50-117."""

    def __init__(self, num_layer=5, emb_dim=300, feat_dim=256, drop_ratio=0, pool="mean"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio = num_layer, emb_dim, feat_dim, drop_ratio
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        if pool in _POOLS:                       # the reference leaves self.pool unset otherwise (83-88)
            self.pool = _POOLS[pool]
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        self.out_lin = nn.Sequential(nn.Linear(feat_dim, feat_dim), nn.ReLU(inplace=True),
                                     nn.Linear(feat_dim, feat_dim // 2))

    def node_embedding(self, x):
        return self.x_embedding1(x[:, 0]) + self.x_embedding2(x[:, 1])      # ginet_molclr.py:103

    def encode(self, data, return_layers=False):
        h = self.node_embedding(data.x)
        layers = []
        for layer in range(self.num_layer):                                   # ginet_molclr.py:105-111
            h = self.gnns[layer](h, data.edge_index, data.edge_attr)
            h = self.batch_norms[layer](h)
            if layer == self.num_layer - 1:
                h = _dropout(self, layer, h)
            else:
                h = _dropout(self, layer, F.relu(h))
            layers.append(h)
        return (h, layers) if return_layers else h

    def forward(self, data):
        h = self.encode(data)
        h = self.pool(h, data.batch)                                          # ginet_molclr.py:113
        h = self.feat_lin(h)
        out = self.out_lin(h)
        return h, out


# ----------------------------------------------------------------------------- GCN
def gcn_norm(edge_index, num_nodes):
    """gcn_molclr.py:27-36.  Computed by the reference and then DISCARDED (gcn_molclr.py:74)."""
    edge_weight = torch.ones((edge_index.size(1),), device=edge_index.device)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_sum(edge_weight, col, num_nodes)
    deg_inv_sqrt = deg.pow_(-0.5)
    deg_inv_sqrt.masked_fill_(deg_inv_sqrt == float("inf"), 0)
    return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]


class GCNConv(nn.Module):
    """gcn_molclr.py:39-91 (weight stored [in, out]; scalar edge embeddings; norm is dead code)."""

    def __init__(self, emb_dim, aggr="add"):
        super().__init__()
        self.emb_dim, self.aggr = emb_dim, aggr
        self.weight = nn.Parameter(torch.empty(emb_dim, emb_dim))
        self.bias = nn.Parameter(torch.empty(emb_dim))
        stdv = math.sqrt(6.0 / (emb_dim + emb_dim))                           # gcn_molclr.py:55-60
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.fill_(0)
        self.edge_embedding1 = nn.Embedding(num_bond_type, 1)
        self.edge_embedding2 = nn.Embedding(num_bond_direction, 1)
        nn.init.xavier_uniform_(self.edge_embedding1.weight.data)
        nn.init.xavier_uniform_(self.edge_embedding2.weight.data)

    def forward(self, x, edge_index, edge_attr):
        edge_index = add_self_loops(edge_index, x.size(0))
        edge_attr = _self_loop_attr(edge_attr, x.size(0))
        edge_embeddings = self.edge_embedding1(edge_attr[:, 0]) + self.edge_embedding2(edge_attr[:, 1])
        edge_index, _unused = gcn_norm(edge_index, x.size(0))                 # result discarded, line 74
        x = x @ self.weight
        out = propagate_add(edge_index, x, edge_embeddings)                   # [E',1] broadcasts over D
        out = out + self.bias
        return out


class GCN(nn.Module):
    """gcn_molclr.py:94-158."""

    def __init__(self, num_layer=5, emb_dim=300, feat_dim=256, drop_ratio=0, pool="mean"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio = num_layer, emb_dim, feat_dim, drop_ratio
        if num_layer < 2:
            raise ValueError("Number of GNN layers must be greater than 1.")
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GCNConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        if pool not in _POOLS:
            raise ValueError("Not defined pooling!")
        self.pool = _POOLS[pool]
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        self.out_lin = nn.Sequential(nn.Linear(feat_dim, feat_dim), nn.ReLU(inplace=True),
                                     nn.Linear(feat_dim, feat_dim // 2))

    def forward(self, data):
        h = self.x_embedding1(data.x[:, 0]) + self.x_embedding2(data.x[:, 1])
        for layer in range(self.num_layer):
            h = self.gnns[layer](h, data.edge_index, data.edge_attr)
            h = self.batch_norms[layer](h)
            if layer == self.num_layer - 1:
                h = _dropout(self, layer, h)
            else:
                h = _dropout(self, layer, F.relu(h))
        h = self.pool(h, data.batch)
        h = self.feat_lin(h)
        out = self.out_lin(h)
        return h, out


# ----------------------------------------------------------------------------- fine-tune GINet
class GINetFinetune(nn.Module):
    """models/ginet_finetune.py:52-147: the same GIN-E encoder, ``feat_lin`` and the ``pred_head`` MLP; returns
    ``(h, pred_head(h))``."""

    def __init__(self, task="classification", num_layer=5, emb_dim=300, feat_dim=512, drop_ratio=0, pool="mean",
                 pred_n_layer=2, pred_act="softplus"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio, self.task = num_layer, emb_dim, feat_dim, drop_ratio, task
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        if pool in _POOLS:
            self.pool = _POOLS[pool]
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        out_dim = {"classification": 2, "regression": 1}[task]                      # ginet_finetune.py:96-99
        self.pred_n_layer = max(1, pred_n_layer)
        act = {"relu": lambda: nn.ReLU(inplace=True), "softplus": nn.Softplus}[pred_act]   # ginet_finetune.py:103-124
        head = [nn.Linear(feat_dim, feat_dim // 2), act()]
        for _ in range(self.pred_n_layer - 1):
            head.extend([nn.Linear(feat_dim // 2, feat_dim // 2), act()])
        head.append(nn.Linear(feat_dim // 2, out_dim))
        self.pred_head = nn.Sequential(*head)

    def forward(self, data):
        h = self.x_embedding1(data.x[:, 0]) + self.x_embedding2(data.x[:, 1])
        for layer in range(self.num_layer):                                         # ginet_finetune.py:136-142
            h = self.gnns[layer](h, data.edge_index, data.edge_attr)
            h = self.batch_norms[layer](h)
            if layer == self.num_layer - 1:
                h = _dropout(self, layer, h)
            else:
                h = _dropout(self, layer, F.relu(h))
        h = self.pool(h, data.batch)
        h = self.feat_lin(h)
        return h, self.pred_head(h)


# ----------------------------------------------------------------------------- fine-tune GIN-E with motif attention
def _group_softmax(src, index, num_groups):
    """torch_geometric.utils.softmax 1.6.3 over dim 0: exp(src - max_g) / (sum_g exp + 1e-16); ATen primitives as PyG calls them
    (scatter-max via ``scatter_reduce(amax)``, ``zeros.scatter_add_``)."""
    idx = index.view(-1, 1).expand_as(src)
    mx = torch.full((num_groups, src.shape[1]), float("-inf"), dtype=src.dtype).scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    out = (src - mx[index]).exp()
    den = torch.zeros(num_groups, src.shape[1], dtype=src.dtype).scatter_add_(0, idx, out)
    return out / (den[index] + 1e-16)


class GINetMotif(nn.Module):
    """models/ginet_finetune_mp.py:52-163 (the motif-level fine-tune model; SURVEY 8f item 3, not built as a kernel path yet):
    the GIN-E encoder and ``feat_lin`` of the fine-tune model, then ``hp = motif_lin(GlobalAttention([motif_embedding[clique_idx];
    h], mol_idx))`` and ``pred_head(cat(h, hp))`` with a ``2 * feat_dim`` wide first layer.  ``forward(data, mol_idx, clique_idx)``
    returns ``(cat(h, hp), pred)``.  ``mol_idx`` lists the molecule of every clique followed by ``arange(G)`` (finetune.py:202-210)."""

    def __init__(self, num_motifs, task="classification", num_layer=5, emb_dim=300, feat_dim=512, drop_ratio=0, pool="mean",
                 pred_n_layer=2, pred_act="softplus"):
        super().__init__()
        self.num_motifs, self.num_layer, self.emb_dim, self.feat_dim = num_motifs, num_layer, emb_dim, feat_dim
        self.drop_ratio, self.task = drop_ratio, task
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.motif_embedding = nn.Embedding(num_motifs, feat_dim)                   # ginet_finetune_mp.py:79
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        if pool in _POOLS:
            self.pool = _POOLS[pool]
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        out_dim = {"classification": 2, "regression": 1}[task]
        self.motif_lin = nn.Linear(feat_dim, feat_dim)                              # :104-105
        nn.init.xavier_uniform_(self.motif_lin.weight.data)
        self.motif_pool = nn.Module()                                               # GlobalAttention(gate_nn=Sequential(Linear(feat_dim, 1))), :107
        self.motif_pool.gate_nn = nn.Sequential(nn.Linear(feat_dim, 1))
        self.motif_pool.nn = None
        self.pred_n_layer = max(1, pred_n_layer)
        if pred_act not in ("relu", "softplus"):
            raise ValueError("Undefined activation function")
        act = {"relu": lambda: nn.ReLU(inplace=True), "softplus": nn.Softplus}[pred_act]   # :111-133
        head = [nn.Linear(2 * feat_dim, feat_dim // 2), act()]
        for _ in range(self.pred_n_layer - 1):
            head.extend([nn.Linear(feat_dim // 2, feat_dim // 2), act()])
        head.append(nn.Linear(feat_dim // 2, out_dim))
        self.pred_head = nn.Sequential(*head)

    def forward(self, data, mol_idx, clique_idx):
        h = self.x_embedding1(data.x[:, 0]) + self.x_embedding2(data.x[:, 1])
        for layer in range(self.num_layer):                                         # :146-152
            h = self.gnns[layer](h, data.edge_index, data.edge_attr)
            h = self.batch_norms[layer](h)
            if layer == self.num_layer - 1:
                h = _dropout(self, layer, h)
            else:
                h = _dropout(self, layer, F.relu(h))
        h = self.pool(h, data.batch)
        h = self.feat_lin(h)
        hp = torch.cat((self.motif_embedding(clique_idx), h), dim=0)                # :157-158
        groups = int(mol_idx[-1]) + 1                                               # GlobalAttention: size = batch[-1] + 1
        gate = _group_softmax(self.motif_pool.gate_nn(hp).view(-1, 1), mol_idx, groups)
        hp = torch.zeros(groups, hp.shape[1], dtype=hp.dtype).scatter_add_(0, mol_idx.view(-1, 1).expand_as(hp), gate * hp)
        hp = self.motif_lin(hp)                                                     # :160
        h = torch.cat((h, hp), dim=1)
        return h, self.pred_head(h)


# ----------------------------------------------------------------------------- fine-tune GCN
class GCNFinetune(nn.Module):
    """models/gcn_finetune.py:94-163: the GCN encoder, ``feat_lin`` and ``pred_lin`` = Linear -> Softplus -> Linear(., 2 | 1);
    returns ``(h, pred_lin(h))``."""

    def __init__(self, task="classification", num_layer=5, emb_dim=300, feat_dim=256, drop_ratio=0, pool="mean"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio, self.task = num_layer, emb_dim, feat_dim, drop_ratio, task
        if num_layer < 2:
            raise ValueError("Number of GNN layers must be greater than 1.")
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GCNConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        if pool not in _POOLS:
            raise ValueError("Not defined pooling!")
        self.pool = _POOLS[pool]
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        out_dim = {"classification": 2, "regression": 1}[task]                      # gcn_finetune.py:133-144
        self.pred_lin = nn.Sequential(nn.Linear(feat_dim, feat_dim // 2), nn.Softplus(), nn.Linear(feat_dim // 2, out_dim))

    def forward(self, data):
        h = self.x_embedding1(data.x[:, 0]) + self.x_embedding2(data.x[:, 1])
        for layer in range(self.num_layer):                                         # gcn_finetune.py:152-158
            h = self.gnns[layer](h, data.edge_index, data.edge_attr)
            h = self.batch_norms[layer](h)
            if layer == self.num_layer - 1:
                h = _dropout(self, layer, h)
            else:
                h = _dropout(self, layer, F.relu(h))
        h = self.pool(h, data.batch)
        h = self.feat_lin(h)
        return h, self.pred_lin(h)
