"""On-device batch construction + augmentation (molclr_augment_views) against the oracle's replay of dataset.py:112-145."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from molclr_b200 import GINet
    from molclr_b200.dataset import PackedMolecules, augment_pair
    from molclr_b200.synth import random_molecule
    from oracle import augment as oaug

DEV = "cuda:0"


def _store(n_mols, seed, mean=25.0, std=6.0):
    rng = np.random.default_rng(seed)
    graphs = [random_molecule(rng, mean, std) for _ in range(n_mols)]
    graphs[0] = (graphs[0][0][:1], np.zeros((0, 2), np.int64), np.zeros((0, 2), np.int64))       # a single atom, no bonds
    graphs[1] = (graphs[1][0][:3], np.array([[0, 1], [1, 2]]), np.array([[0, 0], [3, 1]]))       # 3 atoms, 2 bonds: nothing deleted
    return graphs, PackedMolecules.from_graphs(graphs)


@pytest.mark.parametrize("n_mols,B,seed", [(50, 50, 1), (200, 333, 2), (40, 7, 3)])
def test_views_equal_the_reference_loops_on_the_same_subsets(n_mols, B, seed):
    graphs, store = _store(n_mols, seed)
    ids = np.random.default_rng(seed + 10).integers(0, n_mols, B)       # with repeats
    (bi, bj), (sel_n, sel_b) = (lambda r: (r[:2], r[2]))(augment_pair(store.to(DEV), ids, seed=1234 + seed, return_selection=True))
    node_off, edge_off, bond_off, N, E, M = store.batch_layout(ids)
    sel_n, sel_b = sel_n.cpu().numpy(), sel_b.cpu().numpy()
    for v, got in enumerate((bi, bj)):
        views = []
        for s, mol in enumerate(ids):
            x, bonds, battr = graphs[mol]
            n, m = len(x), len(bonds)
            mask_nodes = np.nonzero(sel_n[v, node_off[s]:node_off[s] + n])[0]
            mask_bonds = np.nonzero(sel_b[v, bond_off[s]:bond_off[s] + m])[0]
            k_n, k_m = oaug.num_masked(n, m)
            assert len(mask_nodes) == k_n and len(mask_bonds) == k_m                    # subset sizes of dataset.py:112-113
            views.append(oaug.augment_view(x, bonds, battr, list(mask_nodes), list(mask_bonds)))
        x, ei, ea, batch = oaug.collate(views)
        assert np.array_equal(got.x.cpu().numpy(), x) and np.array_equal(got.edge_index.cpu().numpy(), ei)
        assert np.array_equal(got.edge_attr.cpu().numpy(), ea) and np.array_equal(got.batch.cpu().numpy(), batch)
        assert got.num_graphs == B and got.edge_index.shape[1] == E and got.x.shape[0] == N
    assert not np.array_equal(sel_n[0], sel_n[1])                                        # the two views draw independently


def test_reproducible_and_seed_dependent_and_uniform():
    graphs, store = _store(64, 5)
    store = store.to(DEV)
    ids = np.arange(64)
    a = augment_pair(store, ids, seed=7)
    b = augment_pair(store, ids, seed=7)
    c = augment_pair(store, ids, seed=8)
    assert torch.equal(a[0].x, b[0].x) and torch.equal(a[1].edge_index, b[1].edge_index)
    assert not torch.equal(a[0].x, c[0].x)
    # uniformity: over many seeds every atom of a 20-atom molecule is masked with probability 5/20
    g20 = [(np.stack([np.full(20, 5), np.zeros(20, np.int64)], 1), np.stack([np.arange(19), np.arange(1, 20)], 1), np.zeros((19, 2), np.int64))]
    s20 = PackedMolecules.from_graphs(g20).to(DEV)
    counts, bcounts, T = np.zeros(20), np.zeros(19), 400
    for t in range(T):
        _, _, (sn, sb) = augment_pair(s20, np.zeros(8, np.int64), seed=1000 + t, return_selection=True)
        counts += sn[0].view(8, 20).sum(0).cpu().numpy(); bcounts += sb[1].view(8, 19).sum(0).cpu().numpy()
    p, pb = counts / (8 * T), bcounts / (8 * T)
    assert abs(p - 0.25).max() < 0.04 and abs(pb - 4 / 19).max() < 0.04, (p, pb)


def test_store_round_trip_errors_and_training_step(tmp_path):
    graphs, store = _store(30, 9)
    store.save(tmp_path / "mols.npz")
    again = PackedMolecules.load(tmp_path / "mols.npz")
    for i in (0, 1, 17):
        for u, w in zip(graphs[i], again.molecule(i)):
            assert np.array_equal(np.asarray(u).reshape(w.shape), w)
    with pytest.raises(IndexError):
        augment_pair(again.to(DEV), [0, 30], seed=0)
    with pytest.raises(RuntimeError):
        augment_pair(again, [0], seed=0)                     # store not on the GPU: no CPU path
    with pytest.raises(ValueError):
        PackedMolecules.from_graphs([(np.array([[119, 0]]), np.zeros((0, 2)), np.zeros((0, 2)))])
    # the views feed the encoder like any other Batch
    bi, bj = augment_pair(again.to(DEV), np.arange(30), seed=3)
    m = GINet(2, 32, 32, 0, "mean").to(DEV)
    h, out = m(bi)
    assert out.shape == (30, 16) and torch.isfinite(out).all()


def _structured_graphs(seed, n_random):
    """Random molecules (some with shuffled / flipped bond lists, which changes the networkx insertion orders the result depends
    on) plus hand-made shapes: a ring, a star, a long chain, a hub with six neighbours, a 70-atom molecule."""
    rng = np.random.default_rng(seed)
    graphs = []
    for t in range(n_random):
        x, bonds, battr = random_molecule(rng, float(rng.choice([8, 25, 45, 70])), 8.0)
        if t % 3 == 0 and len(bonds) > 3:
            perm = rng.permutation(len(bonds))
            bonds, battr = bonds[perm], battr[perm]
            flip = rng.random(len(bonds)) < 0.5
            bonds = np.where(flip[:, None], bonds[:, ::-1], bonds)
        graphs.append((x, np.ascontiguousarray(bonds), battr))
    xs = lambda n: np.stack([np.full(n, 5), np.zeros(n, np.int64)], 1)
    attr = lambda m: np.zeros((m, 2), np.int64)
    ring = np.array([[i, (i + 1) % 12] for i in range(12)])
    star = np.array([[0, i] for i in range(1, 5)])
    chain = np.array([[i + 1, i] for i in range(30)])                    # every bond "backwards"
    hub = np.array([[3, i] for i in (0, 1, 2, 4, 5, 6)] + [[6, 7], [7, 8]])
    graphs += [(xs(12), ring, attr(12)), (xs(5), star, attr(4)), (xs(31), chain, attr(30)), (xs(9), hub, attr(8))]
    return graphs


@pytest.mark.parametrize("aug", ["subgraph", "mix"])
def test_subgraph_and_mix_views_equal_the_reference_algorithm(aug):
    """dataset_subgraph.py:70-88,125-172 / dataset_mix.py:45-68,128-215: the kernel's views against oracle/subgraph.py (networkx
    insertion orders and CPython set iteration order on Python's own containers) replayed with the kernel's random draws."""
    from oracle import subgraph as osg
    graphs = _structured_graphs(5, 160)
    store = PackedMolecules.from_graphs(graphs)
    ids = np.concatenate([np.arange(len(graphs)), np.random.default_rng(9).integers(0, len(graphs), 90)])
    bi, bj, sel = augment_pair(store.to(DEV), ids, seed=777, return_selection=True, aug=aug)
    node_off, _e, bond_off, N, _E, M = store.batch_layout(ids)
    center, percent = sel["center"].cpu().numpy(), sel["percent"].cpu().numpy()
    removed, extra, keep = sel["removed"].cpu().numpy(), sel["extra_masked"].cpu().numpy(), sel["bond_keep"].cpu().numpy()
    eoff, ecnt = sel["edge_offset"].cpu().numpy(), sel["edge_count"].cpu().numpy()
    assert (center[0] != center[1]).all()                                 # random.sample(range(N), 2): two different start atoms
    n_removed_total = 0
    for v, got in enumerate((bi, bj)):
        gx, gei, gea, gb = (t.cpu().numpy() for t in (got.x, got.edge_index, got.edge_attr, got.batch))
        assert gei.shape[1] == int(ecnt[v].sum()) and gea.shape[0] == gei.shape[1]
        for s, mol in enumerate(ids):
            x, bonds, battr = graphs[mol]
            n, m = len(x), len(bonds)
            no, bo = node_off[s], bond_off[s]
            c, pct = int(center[v, s]), float(percent[v, s])
            k = keep[v, bo:bo + m]
            if aug == "subgraph":
                assert pct == 0.25
                xv, ei, ea, rem = osg.subgraph_view(x, bonds, battr, c, pct)
                assert (k != 2).all()
            else:
                assert 0.0 <= pct < 0.2
                surviving = np.nonzero(k > 0)[0]                          # survived the removal, in bond order
                mask_bonds_single = [int(np.nonzero(surviving == b)[0][0]) for b in np.nonzero(k == 2)[0]]
                mask_nodes = [int(a) for a in np.nonzero(extra[v, no:no + n])[0]]
                xv, ei, ea, rem = osg.mix_view(x, bonds, battr, c, pct, mask_nodes, mask_bonds_single)
                want_n, want_b = osg.mix_mask_counts(n, m, len(rem), len(surviving))
                assert len(mask_nodes) == want_n and len(mask_bonds_single) == want_b         # budgets of dataset_mix.py:175-178
                assert not set(mask_nodes) & set(rem)                                          # drawn from the atoms that remain
            n_removed_total += len(rem)
            assert sorted(rem) == np.nonzero(removed[v, no:no + n])[0].tolist(), (mol, v)
            assert np.array_equal(gx[no:no + n], xv), (mol, v)
            assert (gb[no:no + n] == s).all()
            e0, e1 = int(eoff[v, s]), int(eoff[v, s]) + int(ecnt[v, s])
            assert e1 - e0 == ei.shape[1], (mol, v, e1 - e0, ei.shape[1])
            assert np.array_equal(gei[:, e0:e1], ei + no) and np.array_equal(gea[e0:e1], ea), (mol, v)
    assert n_removed_total > 500
    # the views feed the encoder like any other Batch, and are a pure function of the seed
    m = GINet(2, 32, 32, 0, "mean").to(DEV)
    assert torch.isfinite(m(bi)[1]).all()
    bi2, bj2 = augment_pair(store.to(DEV), ids, seed=777, aug=aug)
    assert torch.equal(bi2.edge_index, bi.edge_index) and torch.equal(bj2.x, bj.x)


def test_subgraph_start_atom_without_bonds_fails_like_the_reference():
    """networkx raises when the start atom is not a node of the bond graph (dataset_subgraph.py:79 via G.neighbors)."""
    xs = np.stack([np.full(3, 5), np.zeros(3, np.int64)], 1)
    lone = (xs, np.array([[0, 1]]), np.zeros((1, 2), np.int64))          # atom 2 has no bond; 2 graph nodes: floor(0.25 * 2) = 0 removed
    big = (np.stack([np.full(9, 5), np.zeros(9, np.int64)], 1), np.array([[i, i + 1] for i in range(7)]), np.zeros((7, 2), np.int64))   # atom 8 isolated
    store = PackedMolecules.from_graphs([lone, big]).to(DEV)
    augment_pair(store, [0], seed=1, aug="subgraph")                     # budget 0: the start atom is never looked up, as in the reference
    raised = 0
    for seed in range(40):
        try:
            augment_pair(store, [1], seed=seed, aug="subgraph")
        except ValueError:
            raised += 1
    assert 0 < raised < 40                                                # raises exactly when a view starts from the isolated atom
