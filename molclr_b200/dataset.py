"""Packed, pre-tokenised molecule store and on-device construction of augmented batch pairs.

The reference reads SMILES text, parses every item with RDKit in 12 DataLoader workers and augments with Python loops
(``dataset/dataset.py:46-53,61-145``).  Here a dataset is tokenised ONCE into four flat integer arrays ("molclr-packed v1"),
kept resident in HBM, and every step's two augmented views are produced by one kernel launch (``molclr_augment_views``):

* ``atom_ptr [M+1]``, ``atoms [total atoms]``  = atom type index | chirality << 8          (dataset.py:75-76)
* ``bond_ptr [M+1]``, ``bonds [total bonds]``  = begin | end << 12 | type << 24 | dir << 27   (dataset.py:94-106)

RDKit is not available in this image, so stores are built from already-tokenised graphs (``from_graphs``; the synthetic
generator of ``synth.py``) or loaded from a ``.npz`` written by ``save`` -- converting a SMILES file is a one-off offline job
for a machine that has RDKit.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream
from .batch import Batch

MAX_ATOMS = 4096          # molecule-local atom indices are packed into 12 bits


class PackedMolecules:
    """Flat int32 arrays describing M molecules; ``.to(device)`` moves them to HBM (host copies of the pointer arrays are kept
    for the per-batch size bookkeeping, which is a few microseconds of numpy)."""

    def __init__(self, atom_ptr, atoms, bond_ptr, bonds):
        self.atom_ptr, self.atoms, self.bond_ptr, self.bonds = (torch.as_tensor(np.asarray(a), dtype=torch.int32)
                                                                for a in (atom_ptr, atoms, bond_ptr, bonds))
        self._atom_ptr_host = self.atom_ptr.cpu().numpy().astype(np.int64)
        self._bond_ptr_host = self.bond_ptr.cpu().numpy().astype(np.int64)
        n = len(self._atom_ptr_host) - 1
        if n < 0 or len(self._bond_ptr_host) != n + 1 or self._atom_ptr_host[-1] != self.atoms.numel() \
                or self._bond_ptr_host[-1] != self.bonds.numel():
            raise ValueError("PackedMolecules: inconsistent pointer arrays")
        if n and int(np.diff(self._atom_ptr_host).max()) > MAX_ATOMS:
            raise ValueError(f"PackedMolecules: a molecule has more than {MAX_ATOMS} atoms")

    def __len__(self):
        return len(self._atom_ptr_host) - 1

    @property
    def device(self):
        return self.atoms.device

    @classmethod
    def from_graphs(cls, graphs):
        """graphs: iterable of (x [n,2], bonds [m,2] (begin, end), battr [m,2] (type, dir)) integer arrays, one per molecule
        (the featurisation of dataset.py:61-109 before augmentation)."""
        ap, bp, atoms, bonds = [0], [0], [], []
        for x, b, a in graphs:
            x, b, a = np.asarray(x, np.int64).reshape(-1, 2), np.asarray(b, np.int64).reshape(-1, 2), np.asarray(a, np.int64).reshape(-1, 2)
            if len(x) > MAX_ATOMS or (len(b) and (b.min() < 0 or b.max() >= len(x))):
                raise ValueError("from_graphs: atom index out of range")
            if len(x) and (x[:, 0].min() < 0 or x[:, 0].max() > 118 or x[:, 1].min() < 0 or x[:, 1].max() > 2):
                raise ValueError("from_graphs: atom feature out of range")
            if len(a) and (a[:, 0].min() < 0 or a[:, 0].max() > 3 or a[:, 1].min() < 0 or a[:, 1].max() > 2):
                raise ValueError("from_graphs: bond feature out of range")
            atoms.append(x[:, 0] | (x[:, 1] << 8))
            bonds.append(b[:, 0] | (b[:, 1] << 12) | (a[:, 0] << 24) | (a[:, 1] << 27))
            ap.append(ap[-1] + len(x)); bp.append(bp[-1] + len(b))
        cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0, np.int64)
        return cls(np.asarray(ap), cat(atoms), np.asarray(bp), cat(bonds))

    def molecule(self, i):
        """(x, bonds, battr) of molecule i as int64 numpy arrays (the inverse of ``from_graphs``)."""
        a0, a1, b0, b1 = self._atom_ptr_host[i], self._atom_ptr_host[i + 1], self._bond_ptr_host[i], self._bond_ptr_host[i + 1]
        at = self.atoms[a0:a1].cpu().numpy().astype(np.int64)
        bo = self.bonds[b0:b1].cpu().numpy().astype(np.int64)
        return (np.stack([at & 0xff, at >> 8], 1), np.stack([bo & 0xfff, (bo >> 12) & 0xfff], 1),
                np.stack([(bo >> 24) & 7, (bo >> 27) & 3], 1))

    def save(self, path):
        np.savez(path, format=np.str_("molclr-packed v1"), atom_ptr=self.atom_ptr.cpu().numpy(), atoms=self.atoms.cpu().numpy(),
                 bond_ptr=self.bond_ptr.cpu().numpy(), bonds=self.bonds.cpu().numpy())

    @classmethod
    def load(cls, path):
        z = np.load(path)
        if str(z["format"]) != "molclr-packed v1":
            raise ValueError(f"{path}: not a molclr-packed v1 file")
        return cls(z["atom_ptr"], z["atoms"], z["bond_ptr"], z["bonds"])

    def to(self, device):
        out = object.__new__(PackedMolecules)
        out.atom_ptr, out.atoms, out.bond_ptr, out.bonds = (t.to(device) for t in (self.atom_ptr, self.atoms, self.bond_ptr, self.bonds))
        out._atom_ptr_host, out._bond_ptr_host = self._atom_ptr_host, self._bond_ptr_host
        return out

    def batch_layout(self, mol_ids):
        """Host bookkeeping for a batch: exclusive prefix sums of atoms, surviving directed edges and bonds, and the totals."""
        ids = np.asarray(mol_ids, dtype=np.int64)
        if len(ids) and (ids.min() < 0 or ids.max() >= len(self)):
            raise IndexError("molecule id out of range")
        n = self._atom_ptr_host[ids + 1] - self._atom_ptr_host[ids]
        m = self._bond_ptr_host[ids + 1] - self._bond_ptr_host[ids]
        e = 2 * (m - m // 4)                                  # dataset.py:113,128: 2 (M - floor(0.25 M)) directed edges survive
        ex = lambda v: np.concatenate([[0], np.cumsum(v)[:-1]]).astype(np.int32) if len(v) else np.zeros(0, np.int32)
        return ex(n), ex(e), ex(m), int(n.sum()), int(e.sum()), int(m.sum())


def augment_pair(store, mol_ids, seed, return_selection=False, aug="node"):
    """The two augmented views (Batch_i, Batch_j) of the molecules ``mol_ids`` (a sequence of store indices), built on the
    store's device by one kernel launch.  ``seed``: any 64-bit integer; the result is a pure function of (store, mol_ids, seed).
    With ``return_selection`` also returns (node_masked [2, N] uint8, bond_deleted [2, sum M] uint8), the subsets drawn."""
    dev = store.device
    if dev.type != "cuda":
        raise RuntimeError("molclr_b200: augment_pair needs the store on a CUDA device (no CPU path exists)")
    if aug in ("subgraph", "mix"):
        return _subgraph_pair(store, mol_ids, seed, aug, return_selection)
    if aug != "node":
        raise ValueError("Not defined molecule augmentation!")                        # molclr.py:186-191
    node_off, edge_off, bond_off, N, E, M = store.batch_layout(mol_ids)
    B = len(node_off)
    ids = torch.as_tensor(np.asarray(mol_ids, dtype=np.int64)).to(dev, non_blocking=True)
    offs = torch.from_numpy(np.concatenate([node_off, edge_off, bond_off])).to(dev, non_blocking=True)
    new = lambda *shape: torch.empty(*shape, dtype=torch.int64, device=dev)
    xi, xj, eii, eij, eai, eaj, bi, bj = new(N, 2), new(N, 2), new(2, E), new(2, E), new(E, 2), new(E, 2), new(N), new(N)
    sel_n = torch.empty(2, N, dtype=torch.uint8, device=dev) if return_selection else None
    sel_b = torch.empty(2, M, dtype=torch.uint8, device=dev) if return_selection else None
    status = torch.empty(1, dtype=torch.int32, device=dev)
    i32, i64 = torch.int32, torch.int64
    check(_lib.load().molclr_augment_views(
        ptr(store.atom_ptr, i32), ptr(store.atoms, i32), ptr(store.bond_ptr, i32), ptr(store.bonds, i32), len(store), ptr(ids, i64), B,
        ptr(offs[:B], i32), ptr(offs[B:2 * B], i32), ptr(offs[2 * B:], i32), int(seed) & (2 ** 64 - 1), N, E, M,
        ptr(xi, i64), ptr(eii, i64), ptr(eai, i64), ptr(bi, i64), ptr(xj, i64), ptr(eij, i64), ptr(eaj, i64), ptr(bj, i64),
        ptr(sel_n, torch.uint8), ptr(sel_b, torch.uint8), ptr(status, i32), stream()), "augment_views")
    out = (Batch(xi, eii, eai, bi, num_graphs=B), Batch(xj, eij, eaj, bj, num_graphs=B))
    return out + ((sel_n, sel_b),) if return_selection else out


def _subgraph_pair(store, mol_ids, seed, aug, return_selection):
    """The views of ``dataset_subgraph.py`` (aug="subgraph") / ``dataset_mix.py`` (aug="mix"): two kernel launches and ONE read-back
    of the two surviving-edge totals in between (their number depends on the draws)."""
    dev = store.device
    node_off, _edge_off, bond_off, N, _E, M = store.batch_layout(mol_ids)
    B = len(node_off)
    lib = _lib.load()
    ids = torch.as_tensor(np.asarray(mol_ids, dtype=np.int64)).to(dev, non_blocking=True)
    offs = torch.from_numpy(np.concatenate([node_off, bond_off])).to(dev, non_blocking=True)
    i32, i64, u8 = torch.int32, torch.int64, torch.uint8
    new = lambda *shape: torch.empty(*shape, dtype=i64, device=dev)
    xi, xj, bi, bj = new(N, 2), new(N, 2), new(N), new(N)
    keep = torch.empty(2, max(M, 1), dtype=u8, device=dev)
    counts = torch.empty(2, max(B, 1), dtype=i32, device=dev)
    eoff = torch.empty(2, max(B, 1), dtype=i32, device=dev)
    totals = torch.empty(2, dtype=i32, device=dev)
    centers = torch.empty(2, max(B, 1), dtype=i32, device=dev)
    percents = torch.empty(2, max(B, 1), dtype=torch.float64, device=dev)
    removed = torch.empty(2, max(N, 1), dtype=u8, device=dev)
    extra = torch.empty(2, max(N, 1), dtype=u8, device=dev)
    status = torch.empty(1, dtype=i32, device=dev)
    check(lib.molclr_subgraph_select(
        ptr(store.atom_ptr, i32), ptr(store.atoms, i32), ptr(store.bond_ptr, i32), ptr(store.bonds, i32), len(store), ptr(ids, i64), B,
        ptr(offs[:B], i32), ptr(offs[B:], i32), int(seed) & (2 ** 64 - 1), 1 if aug == "subgraph" else 2, N, M,
        ptr(xi, i64), ptr(bi, i64), ptr(xj, i64), ptr(bj, i64), ptr(keep, u8), ptr(counts, i32), ptr(eoff, i32), ptr(totals, i32),
        ptr(centers, i32), ptr(percents, torch.float64), ptr(removed, u8), ptr(extra, u8), ptr(status, i32), stream()), "subgraph_select")
    E_i, E_j, bits = (int(v) for v in torch.cat([totals, status]).tolist())           # the one host read-back (sizes the edge tensors)
    if bits & 1:
        raise IndexError("molecule id out of range")
    if bits & 2:
        raise ValueError("molclr_b200: a start atom of the subgraph removal has no bonds (the reference's networkx call raises here, "
                         "dataset_subgraph.py:79)")
    if bits & 4:
        raise ValueError("molclr_b200: subgraph augmentation supports molecules of up to 128 atoms with at most 8 distinct neighbours per atom")
    eii, eij, eai, eaj = new(2, E_i), new(2, E_j), new(E_i, 2), new(E_j, 2)
    check(lib.molclr_subgraph_fill(ptr(store.bond_ptr, i32), ptr(store.bonds, i32), len(store), ptr(ids, i64), B, ptr(offs[:B], i32), ptr(offs[B:], i32),
                                   ptr(eoff, i32), ptr(keep, u8), M, ptr(eii, i64), ptr(eai, i64), E_i, ptr(eij, i64), ptr(eaj, i64), E_j, stream()),
          "subgraph_fill")
    out = (Batch(xi, eii, eai, bi, num_graphs=B), Batch(xj, eij, eaj, bj, num_graphs=B))
    if return_selection:
        sel = {"center": centers[:, :B], "percent": percents[:, :B], "removed": removed[:, :N], "extra_masked": extra[:, :N], "bond_keep": keep[:, :M],
               "edge_offset": eoff[:, :B], "edge_count": counts[:, :B]}
        return out + (sel,)
    return out
