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
