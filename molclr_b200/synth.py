"""Synthetic molecule-like graph batches with the reference's augmentation.

There is no RDKit and no dataset here, so inputs are generated with the statistics of the
reference's data pipeline:

* featurisation layout of ``dataset/dataset.py:61-109``: ``x[:,0]`` = atomic number - 1
  (0..117, 118 = mask token), ``x[:,1]`` = chirality 0..2; every bond emitted as two
  consecutive directed edges with identical ``[bond type 0..3, direction 0..2]``;
* augmentation exactly as ``dataset/dataset.py:112-145``: per view independently, mask
  ``max(1, floor(0.25 N))`` atoms to ``[118, 0]`` and delete ``floor(0.25 M)`` bonds (both
  directions), keeping surviving edges in original order;
* collation like PyG's ``DataLoader`` (``dataset/dataset.py:179-184``): concatenate with
  cumulative node offsets and emit ``batch``.
"""
import numpy as np
import torch

from .batch import Batch

MASK_ATOM = 118                                   # len(ATOM_LIST), dataset.py:26,126
_ATOMS = np.array([5, 6, 7, 8, 15, 16])           # C N O F S Cl as (Z - 1)
_ATOM_P = np.array([0.70, 0.12, 0.12, 0.02, 0.02, 0.02])
_CHIR_P = np.array([0.90, 0.05, 0.05])
_BOND_P = np.array([0.50, 0.10, 0.01, 0.39])      # single double triple aromatic
_BDIR_P = np.array([0.95, 0.03, 0.02])


def random_molecule(rng, mean_atoms=25.0, std_atoms=6.0, min_atoms=4):
    """Returns (x [n,2], bonds [m,2] (begin,end), battr [m,2]) as int64 numpy arrays."""
    n = max(min_atoms, int(round(rng.normal(mean_atoms, std_atoms))))
    deg = np.zeros(n, dtype=np.int64)
    bonds = []
    have = set()
    for a in range(1, n):
        lo = max(0, a - 6)
        cand = [p for p in range(lo, a) if deg[p] < 4]
        if not cand:
            cand = [p for p in range(0, a) if deg[p] < 4] or [a - 1]
        p = cand[int(rng.integers(len(cand)))]
        bonds.append((p, a)); have.add((p, a)); deg[p] += 1; deg[a] += 1
    for _ in range(n // 10):                       # ring closures at index distance 4-5
        i = int(rng.integers(0, max(1, n - 5)))
        j = i + int(rng.integers(4, 6))
        if j < n and deg[i] < 4 and deg[j] < 4 and (i, j) not in have:
            bonds.append((i, j)); have.add((i, j)); deg[i] += 1; deg[j] += 1
    m = len(bonds)
    x = np.stack([rng.choice(_ATOMS, size=n, p=_ATOM_P), rng.choice(3, size=n, p=_CHIR_P)], axis=1)
    battr = np.stack([rng.choice(4, size=m, p=_BOND_P), rng.choice(3, size=m, p=_BDIR_P)], axis=1)
    return x.astype(np.int64), np.asarray(bonds, dtype=np.int64).reshape(m, 2), battr.astype(np.int64)


def _directed(bonds, battr, keep):
    b, a = bonds[keep], battr[keep]
    ei = np.empty((2, 2 * len(b)), dtype=np.int64)
    ei[0, 0::2], ei[1, 0::2] = b[:, 0], b[:, 1]
    ei[0, 1::2], ei[1, 1::2] = b[:, 1], b[:, 0]
    return ei, np.repeat(a, 2, axis=0)


def augment(rng, x, bonds, battr):
    """One augmented view (dataset.py:112-145).  Returns x_v, edge_index_v, edge_attr_v."""
    n, m = len(x), len(bonds)
    n_mask, m_mask = max(1, int(np.floor(0.25 * n))), max(0, int(np.floor(0.25 * m)))
    xv = x.copy()
    xv[rng.choice(n, size=n_mask, replace=False)] = (MASK_ATOM, 0)
    keep = np.ones(m, dtype=bool)
    if m_mask:
        keep[rng.choice(m, size=m_mask, replace=False)] = False
    ei, ea = _directed(bonds, battr, keep)
    return xv, ei, ea


def collate(graphs):
    """graphs: list of (x, edge_index, edge_attr) numpy triples -> Batch of CPU int64 tensors."""
    xs, eis, eas, bs, off = [], [], [], [], 0
    for g, (x, ei, ea) in enumerate(graphs):
        xs.append(x); eis.append(ei + off); eas.append(ea.reshape(-1, 2))
        bs.append(np.full(len(x), g, dtype=np.int64)); off += len(x)
    cat = lambda parts, axis, shape: (np.concatenate(parts, axis=axis) if parts else np.zeros(shape, np.int64))
    return Batch(torch.from_numpy(cat(xs, 0, (0, 2))), torch.from_numpy(cat(eis, 1, (2, 0))),
                 torch.from_numpy(cat(eas, 0, (0, 2))), torch.from_numpy(cat(bs, 0, (0,))),
                 num_graphs=len(graphs))


def make_pair_batch(batch_size, seed=0, mean_atoms=25.0, std_atoms=6.0):
    """(Batch_i, Batch_j): the two augmented views of ``batch_size`` synthetic molecules."""
    rng = np.random.default_rng(seed)
    vi, vj = [], []
    for _ in range(batch_size):
        x, bonds, battr = random_molecule(rng, mean_atoms, std_atoms)
        vi.append(augment(rng, x, bonds, battr))
        vj.append(augment(rng, x, bonds, battr))
    return collate(vi), collate(vj)


def make_plain_batch(num_graphs, seed=0, mean_atoms=25.0, std_atoms=6.0):
    """Un-augmented batch (fine-tune shaped inputs)."""
    rng = np.random.default_rng(seed)
    gs = []
    for _ in range(num_graphs):
        x, bonds, battr = random_molecule(rng, mean_atoms, std_atoms)
        ei, ea = _directed(bonds, battr, np.ones(len(bonds), dtype=bool))
        gs.append((x, ei, ea))
    return collate(gs)
