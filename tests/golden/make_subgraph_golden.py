"""Golden vectors for the subgraph-removal augmentation, produced by the reference's own ``removeSubgraph``.

Run in the dev container only (needs /root/reference and networkx):  python tests/golden/make_subgraph_golden.py

``dataset/dataset_subgraph.py`` cannot be imported (rdkit, torch_geometric), so the function definition is taken out of the
file with ``ast`` and executed as it is; the graphs are ``nx.Graph(edges)`` built from the bond lists of synthetic molecules
exactly as :118-121 builds them from RDKit bonds.  Stored per case: the molecule (x, bonds, bond attributes), the centre,
the removed atoms in removal order, ``list(G.edges)`` of the reduced graph, and the surviving directed edges as the loop at
:149-161 emits them (that loop is re-typed here around the reference's data structures: ``(start, end) in G_edges``)."""
import ast
import os
import sys

import networkx as nx
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/dataset/dataset_subgraph.py"
sys.path.insert(0, ROOT)

from molclr_b200.synth import random_molecule   # noqa: E402


REF_MIX = "/root/reference/dataset/dataset_mix.py"


def reference_function(path, name):
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def reference_remove_subgraph():
    return reference_function(REF, "removeSubgraph")


def mix_cases(rng):
    """dataset_mix.py: its own ``remove_subgraph`` (with the empty-level exit) on networkx graphs, then the view loops of
    :150-198 re-typed around those structures with explicit draws (a seeded numpy generator stands in for ``random``)."""
    import math
    remove = reference_function(REF_MIX, "remove_subgraph")
    out, n_cases = {}, 0
    for case in range(40):
        x, bonds, battr = random_molecule(rng, mean_atoms=10.0 + 2.0 * (case % 12), std_atoms=6.0)
        if case % 2:
            flip = rng.random(len(bonds)) < 0.25
            bonds = np.where(flip[:, None], bonds[:, ::-1], bonds)
        if case % 5 == 0 and len(bonds) > 6:          # a disconnected molecule graph: drop a bond of the spanning tree
            keep = np.ones(len(bonds), dtype=bool)
            keep[int(rng.integers(1, len(bonds) // 2))] = False
            bonds, battr = bonds[keep], battr[keep]
        G = nx.Graph([[int(s), int(e)] for s, e in bonds])
        nodes = list(G.nodes)
        if len(nodes) < 2:
            continue
        N, M = len(x), len(bonds)
        center = int(nodes[int(rng.integers(len(nodes)))])
        percent = float(rng.uniform(0, 0.2)) if case % 4 else 0.2
        G2, removed = remove(G, center, percent)
        g_edges = list(G2.edges)
        row, col, feat = [], [], []
        for (s, e), a in zip(bonds, battr):
            s, e = int(s), int(e)
            if (s, e) in g_edges or (e, s) in g_edges:
                row += [s, e]; col += [e, s]; feat += [list(a), list(a)]
        n_surv = len(row) // 2
        k_nodes = max([0, math.floor(0.25 * N) - len(removed)])
        k_edges = max([0, n_surv - math.ceil(0.75 * M)])
        remain = [i for i in range(N) if i not in removed]
        mask_nodes = [int(v) for v in rng.choice(remain, size=k_nodes, replace=False)] if k_nodes else []
        mask_single = [int(v) for v in rng.choice(n_surv, size=k_edges, replace=False)] if k_edges else []
        mask_edges = [2 * i for i in mask_single] + [2 * i + 1 for i in mask_single]
        xv = x.copy()
        for a in range(N):
            if a in mask_nodes or a in removed:
                xv[a, :] = [118, 0]
        ei = np.array([row, col], dtype=np.int64).reshape(2, len(row))
        ea = np.array(feat, dtype=np.int64).reshape(len(row), 2)
        keep_cols = [b for b in range(ea.shape[0]) if b not in mask_edges]
        k = f"m{n_cases}"
        out.update({k + "_x": x, k + "_bonds": bonds, k + "_battr": battr, k + "_center": np.int64(center), k + "_percent": np.float64(percent),
                    k + "_removed": np.asarray(removed, dtype=np.int64), k + "_mask_nodes": np.asarray(mask_nodes, dtype=np.int64),
                    k + "_mask_bonds": np.asarray(mask_single, dtype=np.int64), k + "_xv": xv, k + "_edge_index": ei[:, keep_cols],
                    k + "_edge_attr": ea[keep_cols]})
        n_cases += 1
    out["num_mix_cases"] = np.int64(n_cases)
    return out


def main():
    remove = reference_remove_subgraph()
    rng = np.random.default_rng(20261018)
    out = {}
    n_cases = 0
    for case in range(60):
        x, bonds, battr = random_molecule(rng, mean_atoms=8.0 + 2.0 * (case % 16), std_atoms=6.0)
        if case % 2:      # RDKit bonds are not always (lower index, higher index): ring closures point backwards.  Flip a few.
            flip = rng.random(len(bonds)) < 0.25
            bonds = np.where(flip[:, None], bonds[:, ::-1], bonds)
        G = nx.Graph([[int(s), int(e)] for s, e in bonds])
        nodes = list(G.nodes)
        if len(nodes) < 2:
            continue
        center = int(nodes[int(rng.integers(len(nodes)))])
        percent = (0.25, 0.25, 0.2, 0.5)[case % 4]
        G2, removed = remove(G, center, percent)
        g_edges = list(G2.edges)
        row, col = [], []
        for s, e in bonds:
            if (int(s), int(e)) in g_edges:
                row += [int(s), int(e)]
                col += [int(e), int(s)]
        k = f"c{n_cases}"
        out[k + "_x"] = x
        out[k + "_bonds"] = bonds
        out[k + "_battr"] = battr
        out[k + "_center"] = np.int64(center)
        out[k + "_percent"] = np.float64(percent)
        out[k + "_removed"] = np.asarray(removed, dtype=np.int64)
        out[k + "_gedges"] = np.asarray(g_edges, dtype=np.int64).reshape(len(g_edges), 2)
        out[k + "_edge_index"] = np.asarray([row, col], dtype=np.int64).reshape(2, len(row))
        n_cases += 1
    out["num_cases"] = np.int64(n_cases)
    out.update(mix_cases(rng))
    path = os.path.join(HERE, "subgraph_remove.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, n_cases, "+", int(out["num_mix_cases"]), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
