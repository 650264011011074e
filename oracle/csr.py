"""Oracle: index structures (numpy, stable sorts).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

These define the bit-exact expectation for the CSR the CUDA path builds once per batch.  The
ordering contract comes from how the reference's CPU path sums messages
(``zeros.scatter_add_`` walks edges in input order; PyG appends the self-loop last,
ginet_molclr.py:31-37), so a destination row lists its in-edges in input order and the
self-loop is implicit and last.
"""
import numpy as np


def pack_attr(edge_attr):
    """bond type t in 0..4, direction d in 0..2 -> one byte t*3+d (row of the fused table)."""
    return (edge_attr[:, 0] * 3 + edge_attr[:, 1]).astype(np.uint8)


def build_csr(edge_index, edge_attr, num_nodes):
    """Returns dict of int32/uint8 arrays:
    rowptr[N+1], col[E] (source of each in-edge), eattr[E]  -- destination-sorted, stable;
    rowptr_t[N+1], col_t[E] (destination of each out-edge)  -- source-sorted, stable;
    cnt[N, 8] uint16: per destination node, number of in-edges per bond type (0..4, self-loop
    counted as type 4) and per direction (5..7, self-loop counted as direction 0).
    """
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    ea = np.asarray(edge_attr, dtype=np.int64).reshape(-1, 2)
    n = int(num_nodes)
    order = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    order_t = np.argsort(src, kind="stable")
    rowptr_t = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr_t, src + 1, 1)
    rowptr_t = np.cumsum(rowptr_t)
    cnt = np.zeros((n, 8), dtype=np.int64)
    np.add.at(cnt, (dst, ea[:, 0]), 1)
    np.add.at(cnt, (dst, 5 + ea[:, 1]), 1)
    cnt[:, 4] += 1
    cnt[:, 5] += 1
    return {
        "rowptr": rowptr.astype(np.int32), "col": src[order].astype(np.int32),
        "eattr": pack_attr(ea[order]),
        "rowptr_t": rowptr_t.astype(np.int32), "col_t": dst[order_t].astype(np.int32),
        "cnt": cnt.astype(np.uint16),
    }


def build_graph_segments(batch, num_graphs):
    """gptr[G+1] and the stable node permutation grouping nodes by graph id."""
    b = np.asarray(batch, dtype=np.int64)
    gptr = np.zeros(num_graphs + 1, dtype=np.int64)
    np.add.at(gptr, b + 1, 1)
    return np.cumsum(gptr).astype(np.int32), np.argsort(b, kind="stable").astype(np.int32)
