// On-device construction of the two augmented views of a batch (dataset/dataset.py:112-145 + the PyG DataLoader collate,
// dataset.py:179-184) from a packed, pre-tokenised molecule store resident in HBM: the per-item RDKit parsing, Python
// random.sample, deepcopy loops and 12 DataLoader workers of the reference become one kernel launch per batch.
//
// Packed store ("molclr-packed v1", molclr_b200/dataset.py):  atom_ptr[M+1], atoms[total atoms] = type | chirality << 8
// (dataset.py:75-76);  bond_ptr[M+1], bonds[total bonds] = begin | end << 12 | type << 24 | dir << 27 (molecule-local atom
// indices < 4096, dataset.py:94-106).
//
// Per molecule and view, independently (dataset.py:112-121): mask max(1, floor(N/4)) atoms -> [118, 0] and delete
// floor(M/4) bonds (both directions); surviving bonds keep their order and every bond is emitted as two consecutive
// directed edges with the same attributes.  The uniformly random k-subset (Python's random.sample in the reference) is
// "the k smallest of n counter-based random keys" -- a pure function of (seed, view, batch slot, item), reproducible and
// independent of launch geometry; ties break by index.  Bit parity with Python's Mersenne Twister is impossible by
// construction (as for dropout): the kernel exports the selection it made, and the oracle (oracle/augment.py) applies the
// SAME selection through the reference's loops.
#include "common.cuh"
#include "molclr_b200.h"

namespace molclr {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
// key of item `i` of kind `kind` (0 atoms, 1 bonds) of batch slot `slot`, view `view`
__device__ __forceinline__ uint32_t aug_key(uint64_t seed, int view, int kind, int slot, int i) {
  uint32_t h = mix32((uint32_t)seed ^ 0x9E3779B9u);
  h = mix32(h ^ (uint32_t)(seed >> 32));
  h = mix32(h + (uint32_t)slot * 0x632BE5ABu + (uint32_t)(view * 2 + kind));
  return mix32(h ^ ((uint32_t)i * 0x9E3779B1u + 0x7F4A7C15u));
}
// rank of item i among n keys (ties by index): selected iff rank < k
__device__ __forceinline__ bool aug_selected(uint64_t seed, int view, int kind, int slot, int i, int n, int k) {
  if (k <= 0) return false;
  const uint32_t ki = aug_key(seed, view, kind, slot, i);
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const uint32_t kj = aug_key(seed, view, kind, slot, j);
    rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    if (rank >= k) return false;
  }
  return true;
}

// One warp per (batch slot, view).  blockIdx.y = view.
__global__ void __launch_bounds__(256) augment_views_kernel(
    const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ atoms, const int32_t* __restrict__ bond_ptr,
    const int32_t* __restrict__ bonds, const int64_t* __restrict__ mol_ids, int B, const int32_t* __restrict__ node_off,
    const int32_t* __restrict__ edge_off, const int32_t* __restrict__ bond_off, uint64_t seed, int64_t n_mols,
    int64_t* __restrict__ x0, int64_t* __restrict__ ei0, int64_t* __restrict__ ea0, int64_t* __restrict__ batch0,
    int64_t* __restrict__ x1, int64_t* __restrict__ ei1, int64_t* __restrict__ ea1, int64_t* __restrict__ batch1,
    int64_t E_total, uint8_t* __restrict__ node_masked /* [2][N] or null */, uint8_t* __restrict__ bond_deleted /* [2][sum M] or null */,
    int64_t N_total, int64_t M_total, int32_t* __restrict__ status) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), view = blockIdx.y;
  if (slot >= B) return;
  const int64_t mol = mol_ids[slot];
  if (mol < 0 || mol >= n_mols) { if (lane == 0) atomicOr(status, 1); return; }
  int64_t* x = view ? x1 : x0; int64_t* ei = view ? ei1 : ei0; int64_t* ea = view ? ea1 : ea0; int64_t* batch = view ? batch1 : batch0;
  const int a0 = atom_ptr[mol], n = atom_ptr[mol + 1] - a0, b0 = bond_ptr[mol], m = bond_ptr[mol + 1] - b0;
  const int no = node_off[slot], eo = edge_off[slot], bo = bond_off[slot];
  const int k_n = n > 0 ? max(1, n / 4) : 0, k_m = m / 4;          // dataset.py:112-113 (floor(0.25 N), floor(0.25 M))
  for (int a = lane; a < n; a += 32) {
    const bool masked = aug_selected(seed, view, 0, slot, a, n, k_n);
    const int v = atoms[a0 + a];
    x[2 * (size_t)(no + a)] = masked ? 118 : (v & 0xff);            // [len(ATOM_LIST), 0], dataset.py:126
    x[2 * (size_t)(no + a) + 1] = masked ? 0 : (v >> 8);
    batch[no + a] = slot;
    if (node_masked) node_masked[(size_t)view * N_total + no + a] = masked ? 1 : 0;
  }
  int kept_before = 0;                                              // surviving bonds before this chunk of 32
  for (int c = 0; c < m; c += 32) {
    const int b = c + lane;
    const bool valid = b < m;
    const bool del = valid && aug_selected(seed, view, 1, slot, b, m, k_m);
    const unsigned keep_mask = __ballot_sync(0xffffffffu, valid && !del);
    if (valid && bond_deleted) bond_deleted[(size_t)view * M_total + bo + b] = del ? 1 : 0;
    if (valid && !del) {
      const int pos = kept_before + __popc(keep_mask & ((1u << lane) - 1u));
      const uint32_t w = (uint32_t)bonds[b0 + b];
      const int64_t s = no + (int)(w & 0xfff), d = no + (int)((w >> 12) & 0xfff), t = (w >> 24) & 7, r = (w >> 27) & 3;
      const size_t e = (size_t)eo + 2 * (size_t)pos;
      ei[e] = s; ei[e + 1] = d;                                     // row += [start, end]   (dataset.py:96)
      ei[E_total + e] = d; ei[E_total + e + 1] = s;                 // col += [end, start]   (dataset.py:97)
      ea[2 * e] = t; ea[2 * e + 1] = r; ea[2 * e + 2] = t; ea[2 * e + 3] = r;
    }
    kept_before += __popc(keep_mask);
  }
}

// ------------------------------------------------------------------------------------------------
// Subgraph-removal (dataset/dataset_subgraph.py:70-88,125-172) and mixed (dataset/dataset_mix.py:45-68,128-215) augmentation.
//
// What has to be reproduced is an ALGORITHM ON ORDERED CONTAINERS: networkx graphs are insertion-ordered dicts of insertion-ordered
// adjacency dicts, and `temp = list(set(neighbors))` iterates a CPython set of small ints in hash-table slot order.  Both orders
// decide which atoms of the last breadth-first level are removed when the budget runs out, and the orientation test
// `(start, end) in list(G.edges)` silently drops surviving bonds whose END atom entered the graph before their BEGIN atom
// (dataset_subgraph.py:152-161; the mixed variant accepts either orientation, dataset_mix.py:156).  oracle/subgraph.py restates
// this on Python's own dicts and sets and is pinned by vectors from the reference's functions; this kernel emulates the same
// containers explicitly:
//   * node order   = order of first appearance in the bond list (nx.Graph(edges));
//   * G.copy()     : adjacency of x = first the neighbours that precede x in node order, in node order, then those that follow x, in
//                    the order of x's own bonds;
//   * set(ints)    : CPython's open addressing -- slot = key & mask, 9 linear probes only when they fit (never in an 8-slot table),
//                    then i = (5 i + 1 + (perturb >>= 5)) & mask, growth to the next power of two above 4 * used when 5 * fill >= 3 * mask,
//                    re-insertion in slot order (Objects/setobject.c, set_add_entry / set_insert_clean / set_table_resize).
// Deviations, both where the reference does not terminate normally: a start atom without bonds (networkx raises) removes
// nothing and sets status bit 2; a component smaller than the budget ends the removal (dataset_subgraph.py spins forever; the
// mixed variant breaks, dataset_mix.py:55-56).  One thread per (molecule, view): the work is a short sequential walk.
// ------------------------------------------------------------------------------------------------
constexpr int SG_MAXA = 128, SG_MAXDEG = 8, SG_TABLE = 512;

struct PySmallIntSet {
  int16_t slot[SG_TABLE];
  int mask, fill;
  __device__ void init() { mask = 7; fill = 0; for (int i = 0; i <= 7; ++i) slot[i] = -1; }
  __device__ static void insert_clean(int16_t* table, int mask, int key) {
    unsigned perturb = (unsigned)key;
    int i = key & mask;
    while (true) {
      if (table[i] < 0) { table[i] = (int16_t)key; return; }
      if (i + 9 <= mask)
        for (int j = 1; j <= 9; ++j)
          if (table[i + j] < 0) { table[i + j] = (int16_t)key; return; }
      perturb >>= 5;
      i = (int)(((unsigned)i * 5u + 1u + perturb) & (unsigned)mask);
    }
  }
  __device__ void add(int key) {
    unsigned perturb = (unsigned)key;
    int i = key & mask;
    int e;
    while (true) {
      int probes = (i + 9 <= mask) ? 9 : 0;
      e = i;
      bool found = false;
      do {
        if (slot[e] < 0) { found = true; break; }
        if (slot[e] == key) return;                       // already present
        ++e;
      } while (probes--);
      if (found) break;
      perturb >>= 5;
      i = (int)(((unsigned)i * 5u + 1u + perturb) & (unsigned)mask);
    }
    slot[e] = (int16_t)key;
    ++fill;
    if (fill * 5 < mask * 3) return;
    int newsize = 8;
    while (newsize <= fill * 4) newsize <<= 1;            // (used == fill: nothing is ever deleted from this set)
    int16_t old[SG_MAXA];
    int n = 0;
    for (int k = 0; k <= mask; ++k) if (slot[k] >= 0) old[n++] = slot[k];      // slot order
    mask = newsize - 1;
    for (int k = 0; k <= mask; ++k) slot[k] = -1;
    for (int k = 0; k < n; ++k) insert_clean(slot, mask, old[k]);
  }
};

// mode 1 = subgraph removal (percent 0.25, orientation-sensitive survival), 2 = mixed (percent ~ U(0, 0.2), then random atom
// masking / bond deletion up to the 25 % budgets).  Phase 1: atoms (x, batch), per-bond keep flags, per-task surviving edge counts.
__global__ void __launch_bounds__(128) subgraph_select_kernel(
    const int32_t* __restrict__ atom_ptr, const int32_t* __restrict__ atoms, const int32_t* __restrict__ bond_ptr,
    const int32_t* __restrict__ bonds, const int64_t* __restrict__ mol_ids, int B, const int32_t* __restrict__ node_off,
    const int32_t* __restrict__ bond_off, uint64_t seed, int64_t n_mols, int mode, int64_t* __restrict__ x0, int64_t* __restrict__ batch0,
    int64_t* __restrict__ x1, int64_t* __restrict__ batch1, int64_t N_total, int64_t M_total, uint8_t* __restrict__ bond_keep /* [2][M] */,
    int32_t* __restrict__ edge_count /* [2][B] */, int32_t* __restrict__ center_out /* [2][B] */, double* __restrict__ percent_out /* [2][B] */,
    uint8_t* __restrict__ removed_out /* [2][N] */, uint8_t* __restrict__ extra_masked_out /* [2][N] */, int32_t* __restrict__ status) {
  pdl_sync();
  const int task = blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= 2 * B) return;
  const int view = task / B, slot = task - view * B;
  const int64_t mol = mol_ids[slot];
  edge_count[task] = 0;
  if (mol < 0 || mol >= n_mols) { atomicOr(status, 1); return; }
  int64_t* x = view ? x1 : x0; int64_t* batch = view ? batch1 : batch0;
  const int a0 = atom_ptr[mol], n = atom_ptr[mol + 1] - a0, b0 = bond_ptr[mol], m = bond_ptr[mol + 1] - b0;
  const int no = node_off[slot], bo = bond_off[slot];
  uint8_t* keep = bond_keep + (size_t)view * M_total + bo;
  uint8_t* rem_out = removed_out + (size_t)view * N_total + no;
  uint8_t* ext_out = extra_masked_out + (size_t)view * N_total + no;
  // the two start atoms: random.sample(range(N), 2) (dataset_subgraph.py:109) = the two smallest keys; view i takes the first
  int c_first = -1, c_second = -1;
  {
    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    for (int a = 0; a < n; ++a) {
      const uint32_t k = aug_key(seed, 0, 2, slot, a);
      if (c_first < 0 || k < k1) { k2 = k1; c_second = c_first; k1 = k; c_first = a; }
      else if (c_second < 0 || k < k2) { k2 = k; c_second = a; }
    }
  }
  const int center = view == 0 ? c_first : c_second;
  const double percent = mode == 1 ? 0.25 : (double)aug_key(seed, view, 3, slot, 0) * (0.2 / 4294967296.0);
  center_out[task] = center;
  percent_out[task] = percent;
  bool removed[SG_MAXA];
  int16_t rank[SG_MAXA];
  bool too_big = n > SG_MAXA;
  int n_removed = 0, n_graph = 0;
  if (!too_big) {
    uint8_t adj[SG_MAXA][SG_MAXDEG], gl[SG_MAXA][SG_MAXDEG], deg[SG_MAXA];
    int16_t nodes[SG_MAXA];
    for (int a = 0; a < n; ++a) { removed[a] = false; rank[a] = -1; deg[a] = 0; }
    for (int b = 0; b < m && !too_big; ++b) {              // nx.Graph(edges): node and adjacency insertion order
      const uint32_t w = (uint32_t)bonds[b0 + b];
      const int s = (int)(w & 0xfff), e = (int)((w >> 12) & 0xfff);
      if (rank[s] < 0) { rank[s] = (int16_t)n_graph; nodes[n_graph++] = (int16_t)s; }
      if (rank[e] < 0) { rank[e] = (int16_t)n_graph; nodes[n_graph++] = (int16_t)e; }
      bool have = false;
      for (int k = 0; k < deg[s]; ++k) have |= adj[s][k] == e;
      if (!have) {
        if (deg[s] >= SG_MAXDEG || deg[e] >= SG_MAXDEG) { too_big = true; break; }
        adj[s][deg[s]++] = (uint8_t)e;
        if (e != s) adj[e][deg[e]++] = (uint8_t)s;
      }
    }
    if (!too_big) {
      // Graph.copy(): neighbours that precede x in node order first (in node order), then the others in x's own order
      for (int a = 0; a < n; ++a) {
        int c = 0;
        for (int r = 0; r < rank[a]; ++r) {                // (rank[a] < 0 for atoms without bonds: both loops are empty)
          const int y = nodes[r];
          for (int k = 0; k < deg[a]; ++k) if (adj[a][k] == y) gl[a][c++] = (uint8_t)y;
        }
        for (int k = 0; k < deg[a]; ++k) if (rank[adj[a][k]] >= rank[a]) gl[a][c++] = adj[a][k];
      }
      const int num = (int)floor((double)n_graph * percent);
      if (num > 0 && (center < 0 || rank[center] < 0)) atomicOr(status, 2);      // the reference raises here (start atom without bonds)
      else if (num > 0) {
        int16_t temp[SG_MAXA];
        int nt = 1;
        temp[0] = (int16_t)center;
        PySmallIntSet set;
        while (n_removed < num && nt > 0) {
          set.init();
          for (int t = 0; t < nt; ++t) {                   // neighbours of the whole level, BEFORE anything of it is removed
            const int u = temp[t];
            for (int k = 0; k < deg[u]; ++k) {
              const int v = gl[u][k];
              if (removed[v]) continue;
              bool in_temp = false;
              for (int q = 0; q < nt; ++q) in_temp |= temp[q] == v;
              if (!in_temp) set.add(v);
            }
          }
          for (int t = 0; t < nt && n_removed < num; ++t) { removed[temp[t]] = true; ++n_removed; }
          nt = 0;
          for (int k = 0; k <= set.mask; ++k) if (set.slot[k] >= 0) temp[nt++] = set.slot[k];      // list(set(neighbors))
        }
      }
    }
  }
  if (too_big) atomicOr(status, 4);
  // surviving bonds
  int survivors = 0;
  for (int b = 0; b < m; ++b) {
    const uint32_t w = (uint32_t)bonds[b0 + b];
    const int s = (int)(w & 0xfff), e = (int)((w >> 12) & 0xfff);
    bool ok = !too_big && !removed[s] && !removed[e];
    if (ok && mode == 1) ok = rank[s] < rank[e];           // (start, end) in list(G.edges): u is the endpoint met first in node order
    keep[b] = ok ? 1 : 0;
    survivors += ok ? 1 : 0;
  }
  // mixed: random masking tops the view up to floor(0.25 N) hidden atoms and down to ceil(0.75 M) bonds (dataset_mix.py:175-181)
  int k_nodes = 0, k_bonds = 0;
  if (mode == 2 && !too_big) {
    k_nodes = max(0, n / 4 - n_removed);
    k_bonds = max(0, survivors - (3 * m + 3) / 4);
  }
  for (int a = 0; a < n; ++a) {
    const bool rem = !too_big && removed[a];
    bool extra = false;
    if (k_nodes > 0 && !rem) {                             // random.sample(atom_remain_indices, k): k smallest keys among the remaining atoms
      const uint32_t ka = aug_key(seed, view, 0, slot, a);
      int r = 0;
      for (int j = 0; j < n && r < k_nodes; ++j) {
        if (removed[j]) continue;
        const uint32_t kj = aug_key(seed, view, 0, slot, j);
        r += (kj < ka || (kj == ka && j < a)) ? 1 : 0;
      }
      extra = r < k_nodes;
    }
    const int v = atoms[a0 + a];
    const bool hide = rem || extra;
    x[2 * (size_t)(no + a)] = hide ? 118 : (v & 0xff);
    x[2 * (size_t)(no + a) + 1] = hide ? 0 : (v >> 8);
    batch[no + a] = slot;
    rem_out[a] = rem ? 1 : 0;
    ext_out[a] = extra ? 1 : 0;
  }
  if (k_bonds > 0) {                                        // random.sample(range(surviving), k): k smallest keys among the surviving bonds
    for (int b = 0; b < m; ++b) {
      if (!keep[b]) continue;
      const uint32_t kb = aug_key(seed, view, 1, slot, b);
      int r = 0;
      for (int j = 0; j < m && r < k_bonds; ++j) {
        if (keep[j] != 1 && keep[j] != 3) continue;       // eligible: survived the removal (3 = already drawn in this loop)
        const uint32_t kj = aug_key(seed, view, 1, slot, j);
        r += (kj < kb || (kj == kb && j < b)) ? 1 : 0;
      }
      if (r < k_bonds) keep[b] = 3;                        // survived the subgraph removal, deleted by the random masking
    }
    for (int b = 0; b < m; ++b) if (keep[b] == 3) { keep[b] = 2; --survivors; }
  }
  edge_count[task] = 2 * survivors;
}

// exclusive scan of the per-task edge counts (one block; B is at most a few 10^4) -> edge_off [2][B], totals [2]
__global__ void __launch_bounds__(1024) subgraph_scan_kernel(const int32_t* __restrict__ edge_count, int B, int32_t* __restrict__ edge_off,
                                                             int32_t* __restrict__ totals) {
  pdl_sync();
  __shared__ int32_t part[1024];
  const int view = blockIdx.x, tid = threadIdx.x;
  const int per = (B + 1023) / 1024, lo = min(B, tid * per), hi = min(B, lo + per);
  int32_t s = 0;
  for (int k = lo; k < hi; ++k) s += edge_count[view * B + k];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) { int32_t run = 0; for (int k = 0; k < 1024; ++k) { const int32_t v = part[k]; part[k] = run; run += v; } totals[view] = run; }
  __syncthreads();
  int32_t run = part[tid];
  for (int k = lo; k < hi; ++k) { edge_off[view * B + k] = run; run += edge_count[view * B + k]; }
}

// Phase 2: one warp per (slot, view) emits the surviving bonds, in bond order, as two consecutive directed edges each.
__global__ void __launch_bounds__(256) subgraph_fill_kernel(const int32_t* __restrict__ bond_ptr, const int32_t* __restrict__ bonds,
                                                            const int64_t* __restrict__ mol_ids, int B, const int32_t* __restrict__ node_off,
                                                            const int32_t* __restrict__ bond_off, const int32_t* __restrict__ edge_off,
                                                            const uint8_t* __restrict__ bond_keep, int64_t M_total, int64_t n_mols,
                                                            int64_t* __restrict__ ei0, int64_t* __restrict__ ea0, int64_t E0,
                                                            int64_t* __restrict__ ei1, int64_t* __restrict__ ea1, int64_t E1) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), view = blockIdx.y;
  if (slot >= B) return;
  const int64_t mol = mol_ids[slot];
  if (mol < 0 || mol >= n_mols) return;
  int64_t* ei = view ? ei1 : ei0; int64_t* ea = view ? ea1 : ea0;
  const int64_t E_total = view ? E1 : E0;
  const int b0 = bond_ptr[mol], m = bond_ptr[mol + 1] - b0, no = node_off[slot], bo = bond_off[slot], eo = edge_off[view * B + slot];
  const uint8_t* keep = bond_keep + (size_t)view * M_total + bo;
  int kept_before = 0;
  for (int c = 0; c < m; c += 32) {
    const int b = c + lane;
    const bool k = b < m && keep[b] == 1;
    const unsigned km = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int pos = kept_before + __popc(km & ((1u << lane) - 1u));
      const uint32_t w = (uint32_t)bonds[b0 + b];
      const int64_t s = no + (int)(w & 0xfff), d = no + (int)((w >> 12) & 0xfff), t = (w >> 24) & 7, r = (w >> 27) & 3;
      const size_t e = (size_t)eo + 2 * (size_t)pos;
      ei[e] = s; ei[e + 1] = d;
      ei[E_total + e] = d; ei[E_total + e + 1] = s;
      ea[2 * e] = t; ea[2 * e + 1] = r; ea[2 * e + 2] = t; ea[2 * e + 3] = r;
    }
    kept_before += __popc(km);
  }
}

}  // namespace molclr

using namespace molclr;

extern "C" int molclr_augment_views(const int32_t* atom_ptr, const int32_t* atoms, const int32_t* bond_ptr, const int32_t* bonds,
                                    int64_t n_mols, const int64_t* mol_ids, int64_t B, const int32_t* node_off,
                                    const int32_t* edge_off, const int32_t* bond_off, uint64_t seed, int64_t N_total,
                                    int64_t E_total, int64_t M_total, int64_t* x_i, int64_t* edge_index_i, int64_t* edge_attr_i,
                                    int64_t* batch_i, int64_t* x_j, int64_t* edge_index_j, int64_t* edge_attr_j, int64_t* batch_j,
                                    uint8_t* node_masked, uint8_t* bond_deleted, int32_t* status, cudaStream_t stream) {
  MOLCLR_REQUIRE(B >= 0 && B < (1ll << 31) && N_total < (1ll << 31) && E_total < (1ll << 31), "augment_views: extents exceed int32");
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return cuda_fail(e, "augment_views memset");
  if (B == 0) return 0;
  const int warps = 8;
  MOLCLR_LAUNCH(augment_views_kernel, dim3((unsigned)((B + warps - 1) / warps), 2), 32 * warps, 0, stream,
                atom_ptr, atoms, bond_ptr, bonds, mol_ids, (int)B, node_off, edge_off, bond_off, seed, n_mols, x_i, edge_index_i, edge_attr_i,
                batch_i, x_j, edge_index_j, edge_attr_j, batch_j, E_total, node_masked, bond_deleted, N_total, M_total, status);
  MOLCLR_CHECK_LAUNCH("augment_views");
  return 0;
}

// Subgraph-removal (mode 1) / mixed (mode 2) augmentation, phase 1: atom features and batch vectors of both views, per-bond keep
// flags (0 removed with the subgraph or dropped by the orientation test, 1 kept, 2 deleted by the mixed variant's random masking),
// per-(view, slot) directed-edge counts, their exclusive scan and the two totals (device) -- the caller reads the totals to size
// edge_index / edge_attr, then calls molclr_subgraph_fill.  Selection outputs (centres, fractions, removed / extra-masked atoms)
// let the oracle replay exactly the same draws.
extern "C" int molclr_subgraph_select(const int32_t* atom_ptr, const int32_t* atoms, const int32_t* bond_ptr, const int32_t* bonds,
                                      int64_t n_mols, const int64_t* mol_ids, int64_t B, const int32_t* node_off, const int32_t* bond_off,
                                      uint64_t seed, int mode, int64_t N_total, int64_t M_total, int64_t* x_i, int64_t* batch_i, int64_t* x_j,
                                      int64_t* batch_j, uint8_t* bond_keep, int32_t* edge_count, int32_t* edge_off, int32_t* totals,
                                      int32_t* centers, double* percents, uint8_t* removed, uint8_t* extra_masked, int32_t* status,
                                      cudaStream_t stream) {
  MOLCLR_REQUIRE(mode == 1 || mode == 2, "subgraph_select: mode must be 1 (subgraph) or 2 (mix)");
  MOLCLR_REQUIRE(B >= 0 && B < (1ll << 30) && N_total < (1ll << 31) && M_total < (1ll << 31), "subgraph_select: extents exceed int32");
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return cuda_fail(e, "subgraph_select memset");
  e = cudaMemsetAsync(totals, 0, 2 * sizeof(int32_t), stream);
  if (e != cudaSuccess) return cuda_fail(e, "subgraph_select memset");
  if (B == 0) return 0;
  MOLCLR_LAUNCH(subgraph_select_kernel, (unsigned)((2 * B + 127) / 128), 128, 0, stream, atom_ptr, atoms, bond_ptr, bonds, mol_ids, (int)B, node_off, bond_off, seed,
                n_mols, mode, x_i, batch_i, x_j, batch_j, N_total, M_total, bond_keep,
                edge_count, centers, percents, removed, extra_masked, status);
  MOLCLR_CHECK_LAUNCH("subgraph_select");
  MOLCLR_LAUNCH(subgraph_scan_kernel, 2, 1024, 0, stream, edge_count, (int)B, edge_off, totals);
  MOLCLR_CHECK_LAUNCH("subgraph_scan");
  return 0;
}

extern "C" int molclr_subgraph_fill(const int32_t* bond_ptr, const int32_t* bonds, int64_t n_mols, const int64_t* mol_ids, int64_t B,
                                    const int32_t* node_off, const int32_t* bond_off, const int32_t* edge_off, const uint8_t* bond_keep,
                                    int64_t M_total, int64_t* edge_index_i, int64_t* edge_attr_i, int64_t E_i, int64_t* edge_index_j,
                                    int64_t* edge_attr_j, int64_t E_j, cudaStream_t stream) {
  if (B == 0) return 0;
  const int warps = 8;
  MOLCLR_LAUNCH(subgraph_fill_kernel, dim3((unsigned)((B + warps - 1) / warps), 2), 32 * warps, 0, stream,
                bond_ptr, bonds, mol_ids, (int)B, node_off, bond_off, edge_off, bond_keep, M_total, n_mols, edge_index_i, edge_attr_i, E_i, edge_index_j,
                edge_attr_j, E_j);
  MOLCLR_CHECK_LAUNCH("subgraph_fill");
  return 0;
}
