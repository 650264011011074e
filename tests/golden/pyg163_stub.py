"""Minimal restatement of the torch-geometric 1.6.3 / torch-scatter 2.0.6 / torch-sparse 0.6.9 API surface that the
reference's model files import (README.md:37-38 pins those versions; none of them is vendored under /root/reference or
installable offline).  FIXTURE-GENERATION INFRASTRUCTURE ONLY: `make_encoder_golden.py` installs these modules into
``sys.modules`` so that the UNMODIFIED reference classes (`models/ginet_molclr.py`, `models/gcn_molclr.py`,
`models/ginet_finetune.py`) can be imported and run on CPU to produce golden vectors.  Nothing in the product or in the test
suite imports this file; the tests only read the .npz files it helped to produce.

What is restated, with the published algorithm of the pinned versions:

* ``MessagePassing(aggr='add', flow='source_to_target', node_dim=-2)``: ``propagate`` inspects the signatures of
  ``message`` / ``aggregate`` / ``update``, lifts every ``*_j`` argument with ``index_select(node_dim, edge_index[0])`` and
  every ``*_i`` argument with ``edge_index[1]`` (source_to_target), calls ``message``, aggregates with
  ``torch_scatter.scatter(msg, edge_index[1], dim=node_dim, dim_size=N, reduce=aggr)`` and calls ``update``.  The fused
  ``message_and_aggregate`` path is taken only for ``SparseTensor`` adjacency, never for a COO tensor.
* ``torch_scatter.scatter``: sum = ``zeros(size).scatter_add_`` with the index broadcast to src; mean = sum divided by the
  per-segment count clamped to >= 1; max = segment maximum with empty segments left at 0.
* ``add_self_loops``: ``cat([edge_index, arange(N).repeat(2, 1)], dim=1)`` -- the loops go LAST.
* ``global_{add,mean,max}_pool(x, batch)``: ``scatter(x, batch, dim=0, dim_size=batch.max() + 1, reduce=...)``.
* ``GlobalAttention(gate_nn)`` and ``utils.softmax`` (the motif model, ``models/ginet_finetune_mp.py:107,158``): group-wise
  softmax of the gate (max-shifted, ``+ 1e-16`` in the denominator) and ``scatter_add`` of ``gate * x``; the number of groups
  defaults to ``batch[-1] + 1``.
"""
import inspect
import sys
import types

import torch


# ------------------------------------------------------------------------------------------------ torch_scatter 2.0.6
def _broadcast(src, other, dim):
    if dim < 0:
        dim = other.dim() + dim
    if src.dim() == 1:
        for _ in range(0, dim):
            src = src.unsqueeze(0)
    for _ in range(src.dim(), other.dim()):
        src = src.unsqueeze(-1)
    return src.expand_as(other)


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
        return out.scatter_add_(dim, index, src)
    return out.scatter_add_(dim, index, src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter_sum(src, index, dim, out, dim_size)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count.clamp_(1)
    count = _broadcast(count, out, dim)
    out.true_divide_(count)
    return out


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    index_b = _broadcast(index, src, dim)
    size = list(src.size())
    size[dim] = dim_size if dim_size is not None else int(index.max()) + 1
    res = torch.full(size, float("-inf"), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, index_b, src, reduce="amax", include_self=True)
    touched = torch.zeros(size, dtype=torch.bool, device=src.device).scatter_(dim, index_b, torch.ones_like(index_b, dtype=torch.bool))
    return torch.where(touched, res, torch.zeros_like(res)), None          # empty segments stay 0 (the C++ kernel's fill value)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)


# ------------------------------------------------------------------------------------------------ torch_geometric 1.6.3
def maybe_num_nodes(index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(index.max()) + 1


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    N = maybe_num_nodes(edge_index, num_nodes)
    loop_index = torch.arange(0, N, dtype=torch.long, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        loop_weight = edge_weight.new_full((N,), fill_value)
        edge_weight = torch.cat([edge_weight, loop_weight], dim=0)
    edge_index = torch.cat([edge_index, loop_index], dim=1)
    return edge_index, edge_weight


def global_add_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="add")


def global_mean_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="mean")


def global_max_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="max")


class MessagePassing(torch.nn.Module):
    special_args = {"edge_index", "adj_t", "edge_index_i", "edge_index_j", "size", "size_i", "size_j", "ptr", "index", "dim_size"}

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        assert self.aggr in ["add", "mean", "max", None] and self.flow in ["source_to_target", "target_to_source"]

    def _params(self, fn, pop_first=False):
        names = list(inspect.signature(fn).parameters.keys())
        return names[1:] if pop_first else names

    def propagate(self, edge_index, size=None, **kwargs):
        assert isinstance(edge_index, torch.Tensor) and edge_index.dtype == torch.long and edge_index.dim() == 2 and edge_index.size(0) == 2
        size = [None, None] if size is None else list(size)
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        msg_args, aggr_args, upd_args = self._params(self.message), self._params(self.aggregate, True), self._params(self.update, True)
        user_args = set(msg_args + aggr_args + upd_args) - self.special_args
        coll = {}
        for arg in user_args:
            if arg[-2:] not in ("_i", "_j"):
                coll[arg] = kwargs.get(arg, inspect.Parameter.empty)
            else:
                dim = 0 if arg[-2:] == "_j" else 1
                data = kwargs.get(arg[:-2], inspect.Parameter.empty)
                if isinstance(data, torch.Tensor):
                    if size[dim] is None:
                        size[dim] = data.size(self.node_dim)
                    data = data.index_select(self.node_dim, edge_index[j if arg[-2:] == "_j" else i])
                coll[arg] = data
        coll.update(adj_t=None, edge_index=edge_index, edge_index_i=edge_index[i], edge_index_j=edge_index[j], ptr=None)
        coll["index"] = coll["edge_index_i"]
        coll["size"] = size
        coll["size_i"] = size[1] if size[1] is not None else size[0]
        coll["size_j"] = size[0] if size[0] is not None else size[1]
        coll["dim_size"] = coll["size_i"]

        def with_defaults(fn, names, pop_first):
            sig = inspect.signature(fn).parameters
            out = {}
            for n in names:
                v = coll.get(n, inspect.Parameter.empty)
                if v is inspect.Parameter.empty:
                    if sig[n].default is inspect.Parameter.empty:
                        raise TypeError(f"Required parameter {n} is empty.")
                    v = sig[n].default
                out[n] = v
            return out

        out = self.message(**with_defaults(self.message, msg_args, False))
        out = self.aggregate(out, **with_defaults(self.aggregate, aggr_args, True))
        return self.update(out, **with_defaults(self.update, upd_args, True))

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        assert ptr is None
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs


def softmax(src, index, ptr=None, num_nodes=None):
    """torch_geometric.utils.softmax, 1.6.3: group-wise softmax over dim 0 -- subtract the group maximum, exponentiate, divide by
    (group sum + 1e-16)."""
    N = maybe_num_nodes(index, num_nodes)
    out = src - scatter_max(src, index, dim=0, dim_size=N)[0][index]
    out = out.exp()
    return out / (scatter_add(out, index, dim=0, dim_size=N)[index] + 1e-16)


class GlobalAttention(torch.nn.Module):
    """torch_geometric.nn.GlobalAttention, 1.6.3: r_g = sum_{n in g} softmax_g(gate_nn(x_n)) * nn(x_n); ``size`` defaults to
    ``batch[-1].item() + 1`` (the LAST entry of ``batch``, not its maximum)."""

    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn

    def forward(self, x, batch, size=None):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        size = batch[-1].item() + 1 if size is None else size
        gate = self.gate_nn(x).view(-1, 1)
        x = self.nn(x) if self.nn is not None else x
        assert gate.dim() == x.dim() and gate.size(0) == x.size(0)
        gate = softmax(gate, batch, num_nodes=size)
        return scatter_add(gate * x, batch, dim=0, dim_size=size)


def _unused(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name}: imported by the reference but not on the MolCLR hot path (pyg163_stub)")
    f.__name__ = name
    return f


def install():
    """Registers the stub modules; returns nothing.  Idempotent."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    ts = mod("torch_scatter", scatter=scatter, scatter_add=scatter_add, scatter_sum=scatter_sum, scatter_mean=scatter_mean,
             scatter_max=scatter_max)
    tsp = mod("torch_sparse", SparseTensor=type("SparseTensor", (), {}), matmul=_unused("matmul"), fill_diag=_unused("fill_diag"),
              sum=_unused("sum"), mul=_unused("mul"))
    nn = mod("torch_geometric.nn", MessagePassing=MessagePassing, GCNConv=type("GCNConv", (), {}), global_add_pool=global_add_pool,
             global_mean_pool=global_mean_pool, global_max_pool=global_max_pool, GlobalAttention=GlobalAttention,
             Set2Set=_unused("Set2Set"))
    num_nodes = mod("torch_geometric.utils.num_nodes", maybe_num_nodes=maybe_num_nodes)
    utils = mod("torch_geometric.utils", add_self_loops=add_self_loops, degree=_unused("degree"), softmax=softmax,
                num_nodes=num_nodes)
    mod("torch_geometric", nn=nn, utils=utils, __version__="1.6.3-restated")
    return ts, tsp
