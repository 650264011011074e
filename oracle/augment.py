"""Oracle: the reference's augmentation and collation with the random subsets given explicitly.  TEST INFRASTRUCTURE -- see
oracle/__init__.py.  Restates /root/reference/dataset/dataset.py:112-145 (mask atoms, delete bonds, copy survivors in order)
and the PyG DataLoader collate the reference relies on (dataset.py:179-184): node offsets added to edge_index, ``batch`` vector.
``random.sample`` (dataset.py:114-117) cannot be reproduced on a GPU bit for bit; the CUDA kernel exports the subsets it drew
and this restatement replays them through the reference's loops."""
import numpy as np

MASK_TOKEN = (118, 0)             # [len(ATOM_LIST), 0], dataset.py:126


def num_masked(n_atoms, n_bonds):
    """dataset.py:112-113: num_mask_nodes = max([1, floor(0.25 N)]), num_mask_edges = max([0, floor(0.25 M)])."""
    return max(1, int(np.floor(0.25 * n_atoms))), max(0, int(np.floor(0.25 * n_bonds)))


def augment_view(x, bonds, battr, mask_nodes, mask_bonds_single):
    """One view: x [n,2], bonds [m,2] (begin, end), battr [m,2]; explicit index lists.  Returns x_v, edge_index_v [2,E], edge_attr_v [E,2]."""
    n, m = len(x), len(bonds)
    # dataset.py:94-106: every bond as two consecutive directed edges with identical attributes
    row, col, feat = [], [], []
    for (s, e), a in zip(bonds, battr):
        row += [s, e]; col += [e, s]; feat += [list(a), list(a)]
    edge_index = np.array([row, col], dtype=np.int64).reshape(2, 2 * m)
    edge_attr = np.array(feat, dtype=np.int64).reshape(2 * m, 2)
    mask_edges = [2 * i for i in mask_bonds_single] + [2 * i + 1 for i in mask_bonds_single]      # dataset.py:118-119
    xv = x.copy()
    for a in mask_nodes:                                                                           # dataset.py:124-126
        xv[a, :] = MASK_TOKEN
    k = len(mask_bonds_single)
    ei = np.zeros((2, 2 * (m - k)), dtype=np.int64)                                                # dataset.py:127-134
    ea = np.zeros((2 * (m - k), 2), dtype=np.int64)
    count = 0
    for b in range(2 * m):
        if b not in mask_edges:
            ei[:, count] = edge_index[:, b]; ea[count, :] = edge_attr[b, :]; count += 1
    return xv, ei, ea


def collate(views):
    """PyG Batch.from_data_list as the reference's DataLoader applies it: concatenate, offset edge_index, emit ``batch``."""
    xs, eis, eas, bs, off = [], [], [], [], 0
    for g, (x, ei, ea) in enumerate(views):
        xs.append(x); eis.append(ei + off); eas.append(ea); bs.append(np.full(len(x), g, dtype=np.int64)); off += len(x)
    return (np.concatenate(xs), np.concatenate(eis, axis=1), np.concatenate(eas), np.concatenate(bs))
