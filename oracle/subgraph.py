"""Oracle: the reference's subgraph-removal augmentation.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates /root/reference/dataset/dataset_subgraph.py:
  * ``removeSubgraph`` (:70-88): breadth-first removal of floor(percent * N) atoms starting from a centre atom;
  * the view construction of ``MoleculeDataset.__getitem__`` (:125-172): removed atoms are masked to [118, 0] (NOT deleted, so
    node numbering and the ``batch`` vector are those of the whole molecule), and a bond survives iff the tuple
    (begin, end) is ``in list(G.edges)`` of the reduced networkx graph.

The restatement is dependency-free: networkx's ``Graph`` is modelled as an insertion-ordered dict of insertion-ordered
adjacency dicts, which is what decides (a) the order ``G.neighbors`` yields atoms in and (b) the ORIENTATION in which
``G.edges`` reports an edge -- (u, v) with u the endpoint that comes first in node-insertion order.  The reference tests
``(start, end) in G_edges`` with the bond's own orientation (:152-161), so a surviving bond whose END atom entered the graph
before its BEGIN atom is silently dropped: that quirk is part of the reference's output and is reproduced here.
``temp = list(set(neighbors))`` (:87) is kept verbatim: the iteration order of a CPython set of small ints decides which atoms
of the last breadth-first level are removed when the budget runs out inside a level.

``mix_view`` restates the mixed augmentation of /root/reference/dataset/dataset_mix.py:128-215 (subgraph removal with a random
fraction, then random atom masking / bond deletion up to the 25 % budgets) on the same graph model.

Pinned by tests/golden/subgraph_*.npz (made by tests/golden/make_subgraph_golden.py, which executes the reference's own
``removeSubgraph`` / ``remove_subgraph`` on networkx graphs)."""
import numpy as np

MASK_TOKEN = (118, 0)             # [len(ATOM_LIST), 0], dataset_subgraph.py:139,143


def build_graph(bonds):
    """nx.Graph(edges) (dataset_subgraph.py:118-121): nodes in order of first appearance in the bond list, neighbours in order of
    bond appearance.  Atoms without a bond are NOT nodes of the graph."""
    adj = {}
    for s, e in bonds:
        s, e = int(s), int(e)
        adj.setdefault(s, {})
        adj.setdefault(e, {})
        adj[s][e] = None
        adj[e][s] = None
    return adj


def remove_subgraph(adj, center, percent=0.2, stop_when_exhausted=False):
    """dataset_subgraph.py:70-88 on the dict-of-dicts graph.  Returns (reduced graph, removed atoms in removal order).
    Like the reference it raises (KeyError here, NetworkXError there) when ``center`` is not a node of the graph.
    ``stop_when_exhausted``: the variant of dataset_mix.py:45-68, which leaves the loop when a breadth-first level is empty
    (a connected component smaller than the budget) instead of spinning."""
    assert percent <= 1
    # Graph.copy(): nodes in the original order, then add_edges_from over (u in node order, v in u's adjacency order) -- which
    # can permute the adjacency order of a node relative to the original graph
    g = {n: {} for n in adj}
    for u, nbrs in adj.items():
        for v in nbrs:
            g[u][v] = None
            g[v][u] = None
    num = int(np.floor(len(g) * percent))
    removed = []
    temp = [center]
    while len(removed) < num:
        neighbors = []
        if stop_when_exhausted and len(temp) < 1:                            # dataset_mix.py:55-56
            break
        for n in temp:
            neighbors.extend([i for i in g[n] if i not in temp])             # G.neighbors(n): adjacency insertion order
        for n in temp:
            if len(removed) < num:
                for nb in list(g[n]):                                         # G.remove_node(n)
                    del g[nb][n]
                del g[n]
                removed.append(n)
            else:
                break
        temp = list(set(neighbors))
    return g, removed


def edge_list(g):
    """list(G.edges): every edge once, as (u, v) with u the endpoint met first in node order, v in u's adjacency order."""
    seen, out = set(), []
    for u, nbrs in g.items():
        for v in nbrs:
            if v not in seen:
                out.append((u, v))
        seen.add(u)
    return out


def subgraph_view(x, bonds, battr, center, percent=0.25):
    """One augmented view (dataset_subgraph.py:128-172, percent_i = percent_j = 0.25 at :128): x [n,2], bonds [m,2] (begin, end),
    battr [m,2].  Returns x_v [n,2], edge_index_v [2,E], edge_attr_v [E,2], removed."""
    g, removed = remove_subgraph(build_graph(bonds), int(center), percent)
    xv = np.array(x, dtype=np.int64, copy=True)
    for a in removed:                                                         # :137-139
        xv[a, :] = MASK_TOKEN
    g_edges = edge_list(g)
    row, col, feat = [], [], []
    for (s, e), a in zip(bonds, battr):                                       # :149-161
        if (int(s), int(e)) in g_edges:
            row += [int(s), int(e)]
            col += [int(e), int(s)]
            feat += [list(a), list(a)]
    ei = np.array([row, col], dtype=np.int64).reshape(2, len(row))
    ea = np.array(feat, dtype=np.int64).reshape(len(row), 2)
    return xv, ei, ea, removed


def mix_mask_counts(n_atoms, n_bonds, n_removed, n_surviving_bonds):
    """dataset_mix.py:175-178: random masking tops the view up to floor(0.25 N) hidden atoms and down to ceil(0.75 M) bonds."""
    import math
    return max(0, math.floor(0.25 * n_atoms) - n_removed), max(0, n_surviving_bonds - math.ceil(0.75 * n_bonds))


def mix_view(x, bonds, battr, center, percent, mask_nodes, mask_bonds_single):
    """One view of the mixed augmentation (dataset_mix.py:128-215) with every random draw given explicitly: the centre and the
    fraction of the subgraph removal (:137-139, percent ~ U(0, 0.2)), the extra atoms to mask (drawn from the atoms that remain,
    :179) and the extra bonds to delete (indices into the bonds that SURVIVED the removal, :181).  Unlike dataset_subgraph.py the
    survival test accepts either orientation of the bond (:156,161).  Returns x_v, edge_index_v, edge_attr_v, removed."""
    g, removed = remove_subgraph(build_graph(bonds), int(center), percent, stop_when_exhausted=True)
    g_edges = edge_list(g)
    row, col, feat = [], [], []
    for (s, e), a in zip(bonds, battr):                                       # :150-165
        if (int(s), int(e)) in g_edges or (int(e), int(s)) in g_edges:
            row += [int(s), int(e)]
            col += [int(e), int(s)]
            feat += [list(a), list(a)]
    ei = np.array([row, col], dtype=np.int64).reshape(2, len(row))
    ea = np.array(feat, dtype=np.int64).reshape(len(row), 2)
    mask_edges = [2 * i for i in mask_bonds_single] + [2 * i + 1 for i in mask_bonds_single]       # :183-184
    xv = np.array(x, dtype=np.int64, copy=True)
    for a in range(len(x)):                                                   # :187-190
        if a in mask_nodes or a in removed:
            xv[a, :] = MASK_TOKEN
    k = len(mask_bonds_single)
    ei_f = np.zeros((2, ei.shape[1] - 2 * k), dtype=np.int64)                 # :191-198
    ea_f = np.zeros((ea.shape[0] - 2 * k, 2), dtype=np.int64)
    count = 0
    for b in range(ea.shape[0]):
        if b not in mask_edges:
            ei_f[:, count] = ei[:, b]
            ea_f[count, :] = ea[b, :]
            count += 1
    return xv, ei_f, ea_f, removed
