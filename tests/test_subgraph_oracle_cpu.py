"""CPU test: the subgraph-removal oracle (oracle/subgraph.py, a restatement of dataset/dataset_subgraph.py:70-88,125-172) against
golden vectors produced by the reference's own ``removeSubgraph`` on networkx graphs (tests/golden/make_subgraph_golden.py).
This pins the oracle of the next SURVEY 8(f) row (on-device subgraph-removal augmentation) before any kernel exists."""
import os

import numpy as np
import pytest

from oracle import subgraph as osub

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "subgraph_remove.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_removed_atoms_and_reduced_edge_list_match_the_reference(golden):
    n = int(golden["num_cases"])
    assert n >= 50
    budget_cut = 0
    for c in range(n):
        k = f"c{c}"
        bonds, center, percent = golden[k + "_bonds"], int(golden[k + "_center"]), float(golden[k + "_percent"])
        g, removed = osub.remove_subgraph(osub.build_graph(bonds), center, percent)
        assert removed == golden[k + "_removed"].tolist(), (c, removed, golden[k + "_removed"].tolist())
        assert osub.edge_list(g) == [tuple(e) for e in golden[k + "_gedges"].tolist()], c
        nodes = len(osub.build_graph(bonds))
        assert len(removed) == int(np.floor(nodes * percent))
        budget_cut += int(len(removed) > 0)
    assert budget_cut >= n // 2          # the fixtures do exercise the removal, not only empty budgets


def test_view_construction_matches_the_reference_loop(golden):
    n = int(golden["num_cases"])
    dropped_by_orientation = 0
    for c in range(n):
        k = f"c{c}"
        x, bonds, battr = golden[k + "_x"], golden[k + "_bonds"], golden[k + "_battr"]
        xv, ei, ea, removed = osub.subgraph_view(x, bonds, battr, int(golden[k + "_center"]), float(golden[k + "_percent"]))
        assert np.array_equal(ei, golden[k + "_edge_index"]), c
        # removed atoms are masked, not deleted; every other atom is untouched
        rm = np.zeros(len(x), dtype=bool)
        rm[removed] = True
        assert np.array_equal(xv[rm], np.tile(np.array(osub.MASK_TOKEN), (int(rm.sum()), 1))) and np.array_equal(xv[~rm], x[~rm])
        # edges come in (s, e), (e, s) pairs with identical attributes, none touches a removed atom
        assert ei.shape[1] % 2 == 0 and np.array_equal(ei[:, 0::2], ei[::-1, 1::2]) and np.array_equal(ea[0::2], ea[1::2])
        assert not rm[ei].any()
        # the reference's orientation quirk: bonds between two surviving atoms that its `(start, end) in G.edges` test loses
        alive = [(int(s), int(e)) for s, e in bonds if not rm[s] and not rm[e]]
        dropped_by_orientation += len(alive) - ei.shape[1] // 2
    assert dropped_by_orientation > 0     # (the quirk is real and the fixtures contain it)


def test_mixed_augmentation_matches_the_reference(golden):
    """dataset_mix.py:45-68 (``remove_subgraph`` with the empty-level exit, executed by the fixture generator) and the view loops of
    :150-198 with explicit draws: removed atoms, masked features and surviving edges, including disconnected molecule graphs and a
    random removal fraction."""
    n = int(golden["num_mix_cases"])
    assert n >= 30
    extra_nodes = extra_bonds = short_budget = 0
    for c in range(n):
        k = f"m{c}"
        x, bonds, battr = golden[k + "_x"], golden[k + "_bonds"], golden[k + "_battr"]
        mask_nodes, mask_bonds = golden[k + "_mask_nodes"].tolist(), golden[k + "_mask_bonds"].tolist()
        xv, ei, ea, removed = osub.mix_view(x, bonds, battr, int(golden[k + "_center"]), float(golden[k + "_percent"]), mask_nodes, mask_bonds)
        assert removed == golden[k + "_removed"].tolist(), c
        assert np.array_equal(xv, golden[k + "_xv"]) and np.array_equal(ei, golden[k + "_edge_index"]) and np.array_equal(ea, golden[k + "_edge_attr"]), c
        # the budgets of dataset_mix.py:175-178, recomputed from the oracle's own intermediate counts
        surviving = ei.shape[1] // 2 + len(mask_bonds)
        kn, ke = osub.mix_mask_counts(len(x), len(bonds), len(removed), surviving)
        assert (kn, ke) == (len(mask_nodes), len(mask_bonds)), c
        nodes = len(osub.build_graph(bonds))
        short_budget += int(len(removed) < int(np.floor(nodes * float(golden[k + "_percent"]))))
        extra_nodes += len(mask_nodes)
        extra_bonds += len(mask_bonds)
    assert extra_nodes > 0 and extra_bonds > 0
