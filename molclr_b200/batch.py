"""Duck-typed stand-in for ``torch_geometric.data.Batch`` (PyG is not a dependency).

The reference's models only touch ``data.x``, ``data.edge_index``, ``data.edge_attr``,
``data.batch`` and ``data.to(device)`` (ginet_molclr.py:99-101,113; molclr.py:111-112); a
real PyG ``Batch`` works with the models in this package just as well.
"""
import torch


class Batch:
    """x int64 [N,2]; edge_index int64 [2,E]; edge_attr int64 [E,2]; batch int64 [N]."""

    __slots__ = ("x", "edge_index", "edge_attr", "batch", "num_graphs", "y", "_molclr_plan")

    def __init__(self, x, edge_index, edge_attr, batch, num_graphs=None, y=None):
        self.x, self.edge_index, self.edge_attr, self.batch = x, edge_index, edge_attr, batch
        if num_graphs is None:
            num_graphs = int(batch.max().item()) + 1 if batch.numel() else 0
        self.num_graphs = int(num_graphs)
        self.y = y
        self._molclr_plan = None        # cached CSR plan (molclr_b200.graph.GraphPlan)

    @property
    def num_nodes(self):
        return self.x.size(0)

    def _map(self, fn):
        return Batch(fn(self.x), fn(self.edge_index), fn(self.edge_attr), fn(self.batch), self.num_graphs,
                     None if self.y is None else fn(self.y))

    def to(self, device, non_blocking=False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def __repr__(self):
        return (f"Batch(x=[{self.x.size(0)}, 2], edge_index=[2, {self.edge_index.size(1)}], "
                f"edge_attr=[{self.edge_attr.size(0)}, 2], batch=[{self.batch.size(0)}], num_graphs={self.num_graphs})")
