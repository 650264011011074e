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


def reference_remove_subgraph():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "removeSubgraph")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["removeSubgraph"]


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
    path = os.path.join(HERE, "subgraph_remove.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, n_cases, "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
