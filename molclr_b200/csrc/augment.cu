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
  augment_views_kernel<<<dim3((unsigned)((B + warps - 1) / warps), 2), 32 * warps, 0, stream>>>(
      atom_ptr, atoms, bond_ptr, bonds, mol_ids, (int)B, node_off, edge_off, bond_off, seed, n_mols, x_i, edge_index_i, edge_attr_i,
      batch_i, x_j, edge_index_j, edge_attr_j, batch_j, E_total, node_masked, bond_deleted, N_total, M_total, status);
  MOLCLR_CHECK_LAUNCH("augment_views");
  return 0;
}
