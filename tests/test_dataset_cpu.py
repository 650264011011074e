"""CPU tests of the packed molecule format and of the augmentation oracle (no GPU, no compute through the library)."""
import copy

import numpy as np
import pytest

from molclr_b200.dataset import PackedMolecules, augment_pair
from molclr_b200.synth import random_molecule, augment
from oracle import augment as oaug


def test_packed_format_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    graphs = [random_molecule(rng) for _ in range(25)]
    store = PackedMolecules.from_graphs(graphs)
    assert len(store) == 25 and store.atoms.dtype.is_floating_point is False
    store.save(tmp_path / "s.npz")
    again = PackedMolecules.load(tmp_path / "s.npz")
    for i, g in enumerate(graphs):
        for u, w in zip(g, again.molecule(i)):
            assert np.array_equal(u, w)
    ids = [3, 3, 0, 24]
    node_off, edge_off, bond_off, N, E, M = store.batch_layout(ids)
    n = [len(graphs[i][0]) for i in ids]; m = [len(graphs[i][1]) for i in ids]
    assert node_off.tolist() == [0, n[0], n[0] + n[1], n[0] + n[1] + n[2]] and N == sum(n) and M == sum(m)
    assert E == sum(2 * (k - k // 4) for k in m)
    with pytest.raises(RuntimeError):
        augment_pair(store, ids, seed=0)          # the product path needs the CUDA library and a device-resident store


def test_oracle_replay_matches_the_synthetic_generator():
    """oracle.augment (the reference's loops, explicit subsets) and synth.augment (vectorised numpy) agree on the same subsets."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        x, bonds, battr = random_molecule(rng)
        state = rng.bit_generator.state
        xv, ei, ea = augment(rng, x, bonds, battr)
        rng.bit_generator.state = state                         # replay the same draws
        k_n, k_m = oaug.num_masked(len(x), len(bonds))
        mask_nodes = rng.choice(len(x), size=k_n, replace=False)
        mask_bonds = rng.choice(len(bonds), size=k_m, replace=False) if k_m else []
        xo, eio, eao = oaug.augment_view(x, bonds, battr, list(mask_nodes), list(mask_bonds))
        assert np.array_equal(xv, xo) and np.array_equal(ei, eio) and np.array_equal(ea, eao)


def test_build_dataset_substitutes_nothing_silently():
    from molclr_b200.trainer import DEFAULT_CONFIG, build_dataset
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["dataset"]["data_path"] = "data/pubchem-10m-clean.txt"          # the reference's config.yaml value: SMILES text needs RDKit
    with pytest.raises(ValueError):
        build_dataset(cfg)
    cfg["dataset"]["data_path"] = "synthetic:100"
    cfg["fp16_precision"] = True
    with pytest.raises(ValueError):
        build_dataset(cfg)
    cfg["fp16_precision"] = False
    cfg["aug"] = "bogus"
    with pytest.raises(ValueError):
        build_dataset(cfg)
