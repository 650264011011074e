"""CPU test of the ALGORITHM the subgraph-augmentation kernel implements (csrc/augment.cu): a Python restatement of its explicit
container emulation -- networkx insertion orders as arrays, CPython's set of small ints as an open-addressing table -- against
oracle/subgraph.py, which runs on Python's own dicts and sets (and is pinned by vectors from the reference's functions).  The GPU
test (tests/test_gpu_augment.py) then checks the kernel itself against the oracle."""
import numpy as np

from molclr_b200.synth import random_molecule
from oracle import subgraph as osg


class SmallIntSet:
    """CPython's set of small non-negative ints (Objects/setobject.c: set_add_entry, set_insert_clean, set_table_resize), as the
    kernel's PySmallIntSet implements it: slot = key & mask; 9 linear probes only when they fit below the end of the table;
    then i = 5 i + 1 + (perturb >>= 5); growth to the next power of two above 4 * used when 5 * fill >= 3 * mask; re-insertion in
    slot order.  Iteration = slot order."""

    def __init__(self):
        self.mask, self.fill, self.slot = 7, 0, [-1] * 8

    @staticmethod
    def _insert_clean(table, mask, key):
        perturb, i = key, key & mask
        while True:
            if table[i] < 0:
                table[i] = key
                return
            if i + 9 <= mask:
                for j in range(1, 10):
                    if table[i + j] < 0:
                        table[i + j] = key
                        return
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & mask

    def add(self, key):
        perturb, i = key, key & self.mask
        while True:
            probes = 9 if i + 9 <= self.mask else 0
            e, found = i, False
            while True:
                if self.slot[e] < 0:
                    found = True
                    break
                if self.slot[e] == key:
                    return
                e += 1
                if probes == 0:
                    break
                probes -= 1
            if found:
                break
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & self.mask
        self.slot[e] = key
        self.fill += 1
        if self.fill * 5 < self.mask * 3:
            return
        newsize = 8
        while newsize <= self.fill * 4:
            newsize <<= 1
        old = [k for k in self.slot if k >= 0]
        self.mask, self.slot = newsize - 1, [-1] * newsize
        for k in old:
            self._insert_clean(self.slot, self.mask, k)

    def items(self):
        return [k for k in self.slot if k >= 0]


def emulate(n, bonds, center, percent, mode):
    """The kernel's walk: (removed flags, per-bond survival) for one view; mode 1 = subgraph (orientation-sensitive), 2 = mixed."""
    rank, nodes, adj = [-1] * n, [], [[] for _ in range(n)]
    for s, e in bonds:                                    # nx.Graph(edges): node / adjacency insertion order
        s, e = int(s), int(e)
        for a in (s, e):
            if rank[a] < 0:
                rank[a] = len(nodes)
                nodes.append(a)
        if e not in adj[s]:
            adj[s].append(e)
            if e != s:
                adj[e].append(s)
    copied = []                                           # Graph.copy(): preceding neighbours in node order, then the others in own order
    for a in range(n):
        first = [nodes[r] for r in range(max(rank[a], 0)) if nodes[r] in adj[a]]
        copied.append(first + [y for y in adj[a] if rank[y] >= rank[a]])
    removed, n_removed = [False] * n, 0
    num = int(np.floor(len(nodes) * percent))
    if num > 0 and rank[center] >= 0:
        temp = [center]
        while n_removed < num and temp:
            level = SmallIntSet()
            for u in temp:
                for v in copied[u]:
                    if not removed[v] and v not in temp:
                        level.add(v)
            for t in temp:
                if n_removed < num:
                    removed[t] = True
                    n_removed += 1
            temp = level.items()
    keep = []
    for s, e in bonds:
        ok = not removed[int(s)] and not removed[int(e)]
        if ok and mode == 1:
            ok = rank[int(s)] < rank[int(e)]
        keep.append(ok)
    return removed, keep


def test_container_emulation_matches_oracle_on_random_molecules():
    rng = np.random.default_rng(0)
    total = 0
    for trial in range(500):
        x, bonds, battr = random_molecule(rng, mean_atoms=rng.choice([8, 25, 45, 70]), std_atoms=8)
        if trial % 5 == 0 and len(bonds) > 3:             # shuffled / flipped bond lists exercise the insertion orders
            perm = rng.permutation(len(bonds))
            bonds, battr = bonds[perm], battr[perm]
            flip = rng.random(len(bonds)) < 0.5
            bonds = np.where(flip[:, None], bonds[:, ::-1], bonds)
        n = len(x)
        for center in rng.choice(n, size=min(3, n), replace=False):
            for mode, percent in ((1, 0.25), (2, float(rng.uniform(0, 0.2)))):
                total += 1
                removed, keep = emulate(n, bonds, int(center), percent, mode)
                g, want_removed = osg.remove_subgraph(osg.build_graph(bonds), int(center), percent, stop_when_exhausted=(mode == 2))
                edges = osg.edge_list(g)
                if mode == 1:
                    want_keep = [(int(s), int(e)) in edges for s, e in bonds]
                else:
                    want_keep = [((int(s), int(e)) in edges) or ((int(e), int(s)) in edges) for s, e in bonds]
                assert sorted(want_removed) == [i for i, r in enumerate(removed) if r], (trial, center, mode, percent)
                assert want_keep == keep, (trial, center, mode, percent)
    assert total > 2000


def test_small_int_set_iterates_like_cpython():
    rng = np.random.default_rng(1)
    for _ in range(2000):
        keys = [int(k) for k in rng.integers(0, 200, size=int(rng.integers(1, 60)))]
        s = SmallIntSet()
        for k in keys:
            s.add(k)
        assert s.items() == list(set(keys)), keys
