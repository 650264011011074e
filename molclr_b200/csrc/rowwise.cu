// Row-wise (HBM-bound) kernels of the encoder: node embedding, GINE neighbour aggregation
// forward/backward, BatchNorm finalisation / backward apply, segmented pooling, table gradients.
//
// All of them use one mapping: a warp owns a node (or graph) row of D fp32 features and its lanes
// stride over the D/4 float4 chunks, so every global access is a coalesced 128-bit load/store and
// a row is summed in a fixed order without atomics (deterministic, and bit-exact against the CPU
// reference order where that matters).  Grids are persistent: (#SM x resident CTAs) blocks that
// grid-stride over rows.
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm.cuh"
#include "molclr_b200.h"
#include "ptx.cuh"

namespace molclr {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, int64_t rows_per_block_iter, int64_t rows) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t want = (rows + rows_per_block_iter - 1) / rows_per_block_iter;
  int64_t cap = (int64_t)per_sm * sm_count();
  int64_t g = want < cap ? want : cap;
  return (int)(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------------
// Dropout (ginet_molclr.py:108-111: h = dropout(relu(BN(z))), last layer dropout(BN(z))).  The normalised activations are
// never stored, so the Bernoulli mask is a pure function of (seed, node, feature) -- a counter-based hash -- that every
// consumer (the next layer's gather, the pooling, and their backward kernels) recomputes.  One 64-bit draw per float4:
// four 16-bit lanes, keep iff draw >= thr (thr = round(p * 65536)); kept values are scaled by 1/(1-p).
// ------------------------------------------------------------------------------------------------
struct DropCfg { uint32_t seed, thr; float scale; };
static DropCfg make_drop(uint32_t seed, float p) {
  DropCfg d;
  d.seed = seed;
  d.thr = p > 0.f ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
  d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  return d;
}
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float4 drop_mask4(const DropCfg& d, int row, int q, int D4) {
  const uint32_t idx = (uint32_t)row * (uint32_t)D4 + (uint32_t)q;
  const uint32_t h0 = fmix32(idx ^ fmix32(d.seed ^ 0x9E3779B9u)), h1 = fmix32(h0 + 0x6A09E667u + d.seed);
  float4 m;
  m.x = (h0 & 0xffffu) >= d.thr ? d.scale : 0.f; m.y = (h0 >> 16) >= d.thr ? d.scale : 0.f;
  m.z = (h1 & 0xffffu) >= d.thr ? d.scale : 0.f; m.w = (h1 >> 16) >= d.thr ? d.scale : 0.f;
  return m;
}
// (a rounded product, never contracted into an FMA with a following add: dropout(h) is a tensor of its own in the reference)
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) {
  return make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
}

#define NCH_DISPATCH(D4, ...)                                                    \
  do {                                                                           \
    if ((D4) <= 32) { constexpr int NCH = 1; __VA_ARGS__; }                             \
    else if ((D4) <= 64) { constexpr int NCH = 2; __VA_ARGS__; }                        \
    else if ((D4) <= 96) { constexpr int NCH = 3; __VA_ARGS__; }                        \
    else if ((D4) <= 128) { constexpr int NCH = 4; __VA_ARGS__; }                       \
    else { set_error("feature width %d not supported (D <= 512, D %% 4 == 0)", 4 * (D4)); return -2; } \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Node embedding  h0[n] = E1[x[n,0]] + E2[x[n,1]]            (ginet_molclr.py:103, gcn_molclr.py:144)
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kRowThreads) embed_nodes_fwd_kernel(
    const int32_t* __restrict__ xpacked, const float* __restrict__ E1, const float* __restrict__ E2,
    int N, int D, float* __restrict__ out) {
  pdl_sync();
  const int lane = threadIdx.x & 31, D4 = D >> 2;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  for (int n = warp; n < N; n += nwarps) {
    const int xp = xpacked[n];
    const float* r1 = E1 + (size_t)(xp & 0xff) * D;
    const float* r2 = E2 + (size_t)(xp >> 8) * D;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) st_f4(out + (size_t)n * D + 4 * q, f4_add(ldg_f4(r1 + 4 * q), ldg_f4(r2 + 4 * q)));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GINE neighbour aggregation, forward.
//   a[i] = sum_{e in row i, input order} ( f(src[col[e]]) + (B1[t_e] + B2[d_e]) )  +  ( f(src[i]) + (B1[4] + B2[0]) )
// (ginet_molclr.py:29-44 + PyG propagate/scatter-add).  f is the identity for layer 0 and the
// previous layer's BatchNorm + ReLU otherwise (ginet_molclr.py:107-111), applied on the fly from
// per-feature (scale, shift) so the normalised activations never round-trip through HBM.
// Both bond tables are folded into one 15-row table staged in shared memory.
// ------------------------------------------------------------------------------------------------
// Index prefetch for the warp-per-row CSR gather of the backward pass (measured: 56 -> 49 us on the plain transposed
// gather; the forward kernel, whose BatchNorm/ReLU/table terms keep more registers live, is faster without it).  Rows are dealt round-robin to the warps of the grid (so that the rows
// in flight at any time form one contiguous, DRAM-page-friendly window); a row's rowptr pair is fetched two iterations
// ahead and its first 32 column/attribute entries (lane l holds entry beg + l) one iteration ahead, which leaves the feature
// rows as the only dependent global loads of an iteration (instead of a rowptr -> col -> features chain).
struct CsrRowPrefetch {
  const int32_t* rowptr; const int32_t* col; const uint8_t* attr;
  int n_rows, lane;
  int b0, e0, c0, a0;      // current row: range and this lane's entry
  int b1, e1, c1, a1;      // next row
  int b2, e2;              // row after next: range only
  __device__ __forceinline__ void ld_rp(int n, int& b, int& e) const {
    if (n < n_rows) { b = __ldg(rowptr + n); e = __ldg(rowptr + n + 1); } else { b = 0; e = 0; }
  }
  __device__ __forceinline__ void ld_ch(int b, int e, int& c, int& a) const {
    const int idx = b + lane;
    const bool ok = idx < e;
    c = ok ? __ldg(col + idx) : 0;
    a = (ok && attr) ? (int)__ldg(attr + idx) : 0;
  }
  __device__ __forceinline__ void init(const int32_t* rowptr_, const int32_t* col_, const uint8_t* attr_, int n_rows_, int lane_, int i,
                                       int stride) {
    rowptr = rowptr_; col = col_; attr = attr_; n_rows = n_rows_; lane = lane_;
    ld_rp(i, b0, e0); ld_rp(i + stride, b1, e1); ld_rp(i + 2 * stride, b2, e2);
    ld_ch(b0, e0, c0, a0); ld_ch(b1, e1, c1, a1);
  }
  // entry e of the current row (warp-uniform e in [b0, e0))
  __device__ __forceinline__ void get(int e, int& c, int& a) const {
    const int k = e - b0;
    if (k < 32) { c = __shfl_sync(0xffffffffu, c0, k); a = __shfl_sync(0xffffffffu, a0, k); }
    else { c = __ldg(col + e); a = attr ? (int)__ldg(attr + e) : 0; }
  }
  // call at the end of the iteration of row i
  __device__ __forceinline__ void advance(int i, int stride) {
    b0 = b1; e0 = e1; c0 = c1; a0 = a1;
    b1 = b2; e1 = e2;
    ld_ch(b1, e1, c1, a1);
    ld_rp(i + 3 * stride, b2, e2);
  }
};

// SCALAR (GCNConv, gcn_molclr.py:72-88): the bond tables are [5][1] / [3][1] scalars broadcast over the D features and a
// bias row is added after the sum (`out += bias`, gcn_molclr.py:81-82); the staged table then holds splatted scalars.
template <int NCH, bool HAS_BN, bool SCALAR, bool DROP>
__global__ void __launch_bounds__(kRowThreads) gine_aggregate_fwd_kernel(
    const float* __restrict__ src, const float* __restrict__ coef, int relu,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const uint8_t* __restrict__ eattr,
    const float* __restrict__ B1, const float* __restrict__ B2, const float* __restrict__ bias, int N, int D,
    float* __restrict__ out, long long ld_out, int round_out, float* __restrict__ out_lo, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* ee = sm4;                       // [15][D4]
  float4* sc = ee + kNumEdgeClass * D4;   // [D4] scale  (SCALAR: the bias row)
  float4* sh = sc + D4;                   // [D4] shift
  for (int i = threadIdx.x; i < kNumEdgeClass * D4; i += blockDim.x) {
    const int cls = i / D4, q = i - cls * D4;
    if (SCALAR) {
      const float sv = __ldg(B1 + cls / 3) + __ldg(B2 + cls % 3);
      ee[i] = make_float4(sv, sv, sv, sv);
    } else {
      ee[i] = f4_add(ldg_f4(B1 + (size_t)(cls / 3) * D + 4 * q), ldg_f4(B2 + (size_t)(cls % 3) * D + 4 * q));
    }
  }
  if (HAS_BN)
    for (int i = threadIdx.x; i < D4; i += blockDim.x) { sc[i] = ldg_f4(coef + 4 * i); sh[i] = ldg_f4(coef + D + 4 * i); }
  if (SCALAR)
    for (int i = threadIdx.x; i < D4; i += blockDim.x) sc[i] = bias ? ldg_f4(bias + 4 * i) : f4_zero();
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  auto act = [&](float4 v, int q, int row) -> float4 {
    if (HAS_BN) {
      const float4 s = sc[q], b = sh[q];
      v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y); v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (DROP) v = f4_mul(v, drop_mask4(drop, row, q, D4));
    }
    return v;
  };
  const int row_lines = (D * 4 + 127) / 128;                 // 128-byte lines per feature row
  for (int i = warp; i < N; i += nwarps) {
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    // the next row this warp will own: pull its feature row (first touch: HBM latency) and its CSR entries into L2 now
    // (measured: 86 -> 79 us for the BatchNorm-fused variant; the leaner layer-0 variant gets slower, so it is left alone)
    if (HAS_BN && i + nwarps < N && lane < row_lines) prefetch_l2(src + (size_t)(i + nwarps) * D + 32 * lane);
    if (HAS_BN && i + nwarps < N && lane == 31) prefetch_l2(rowptr + i + nwarps);
    float4 acc[NCH], self[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      acc[j] = f4_zero();
      self[j] = (q < D4) ? ldg_f4(src + (size_t)i * D + 4 * q) : f4_zero();
    }
    for (int e = beg; e < end; e += 2) {                 // two neighbour rows in flight per step
      const bool two = (e + 1 < end);
      const int s0 = __ldg(col + e), a0 = __ldg(eattr + e);
      const int s1 = two ? __ldg(col + e + 1) : s0, a1 = two ? __ldg(eattr + e + 1) : a0;
      float4 v0[NCH], v1[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int q = lane + 32 * j;
        v0[j] = (q < D4) ? ldg_f4(src + (size_t)s0 * D + 4 * q) : f4_zero();
        v1[j] = (q < D4 && two) ? ldg_f4(src + (size_t)s1 * D + 4 * q) : f4_zero();
      }
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int q = lane + 32 * j;
        if (q < D4) {
          acc[j] = f4_add(acc[j], f4_add(act(v0[j], q, s0), ee[a0 * D4 + q]));
          if (two) acc[j] = f4_add(acc[j], f4_add(act(v1[j], q, s1), ee[a1 * D4 + q]));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        float4 r = f4_add(acc[j], f4_add(act(self[j], q, i), ee[kSelfLoopAttr * D4 + q]));   // self loop LAST
        if (SCALAR) r = f4_add(r, sc[q]);                                                  // then the bias row
        if (out_lo) st_f4(out_lo + (size_t)i * ld_out + 4 * q, f4_tf32_residual(r));
        if (round_out) r = f4_tf32(r);
        st_f4(out + (size_t)i * ld_out + 4 * q, r);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GINE aggregation forward, shared-memory tile variant (the GINEConv product path; same sums in the same order).
// The warp-per-row kernel above is bound by latency, not bandwidth: a row is one long dependent chain (rowptr -> col ->
// feature rows from L2 -> adds) per warp, a quarter of the lane slots idle (75 float4 chunks on 3 x 32 lanes), and its
// registers allow 24 warps per SM.  Here molecules being small connected blocks of CONSECUTIVE rows is used: a producer
// thread streams tiles [t*T, (t+1)*T) of `src` (one 1D bulk copy of T*D*4 bytes) and of the fixed-width neighbour table
// nbr[N][8] (see molclr_plan_build) through a ring of shared-memory stages; ~1000 consumer threads are laid out as
// R rows x D/4 float4 chunks (thread = one chunk of one row: no idle lanes, ~45 registers, 32 warps per SM) and gather
// from the staged tile.  A neighbour outside the tile (a molecule cut by the tile boundary) or a row with more than 8
// in-edges falls back to global loads.  The previous layer's BatchNorm coefficients and the self-loop table row of the
// thread's chunk live in registers, so the staged tile is read-only and every feature row crosses L2 -> SM once.
// ------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 1024;
constexpr int kTileConsumers = kTileThreads - 32;      // the last warp hosts the producer thread
constexpr int kTileMaxStages = 12;
static int env_int(const char* name, int dflt) { const char* e = debug_env(name); return e ? atoi(e) : dflt; }
// ring depth and rows of a stage handled by one consumer thread (overridable for measurements; sweep on B200 at the bench
// shape, tools/sweep_aggregate.sh: 3 x 3 and 4 x 2 are best, deeper rings are SLOWER -- more reads in flight delay the writes)
static const int g_tile_stages = env_int("MOLCLR_AGG_STAGES", 3), g_tile_rows_per_slot = env_int("MOLCLR_AGG_ROWS", 3);
static const int g_tile_store_cs = env_int("MOLCLR_AGG_STORE_CS", 0), g_tile_blocked = env_int("MOLCLR_AGG_BLOCKED", 0);
constexpr uint32_t kNbrLong = 0xFFFFFFFEu;      // entries >= kNbrLong end a row's list: 0xFFFFFFFF = empty, kNbrLong in [0] = use the CSR

// A/B switch for measurements: MOLCLR_AGG_TILE=0 forces the warp-per-row kernel
static const bool g_aggregate_tile = [] { const char* e = debug_env("MOLCLR_AGG_TILE"); return !(e && e[0] == '0'); }();

static size_t aggregate_tile_smem(int D, int T, int stages) {
  return (size_t)kNumEdgeClass * D * 4 + (size_t)stages * T * (D * 4 + 32) + 2 * kTileMaxStages * sizeof(uint64_t);
}
static int aggregate_tile_rows(int D, int stages) {       // rows per stage: a whole number of row slots per consumer thread that fits
  const int R = kTileConsumers / (D / 4);
  int per_slot = g_tile_rows_per_slot;
  while (per_slot > 1 && aggregate_tile_smem(D, per_slot * R, stages) > 220 * 1024) --per_slot;
  return per_slot * R;
}

template <bool HAS_BN, bool DROP>
__global__ void __launch_bounds__(kTileThreads, 1) gine_aggregate_fwd_tile_kernel(
    const float* __restrict__ src, const float* __restrict__ coef, int relu,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const uint8_t* __restrict__ eattr,
    const uint32_t* __restrict__ nbr, const float* __restrict__ B1, const float* __restrict__ B2, int N, int D, int T,
    int kTileStages, int store_cs, int blocked, float* __restrict__ out, long long ld_out, int round_out, float* __restrict__ out_lo,
    const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* ee = sm4;                                                  // [15][D4]
  float4* feat = ee + kNumEdgeClass * D4;                            // [stages][T][D4]
  uint32_t* nb = reinterpret_cast<uint32_t*>(feat + (size_t)kTileStages * T * D4);   // [stages][T][8]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(nb + (size_t)kTileStages * T * 8);
  uint64_t* empty_bar = full_bar + kTileMaxStages;
  const int R = kTileConsumers / D4;                                 // rows in flight: consumer thread = (row slot, chunk)
  const int n_active = R * D4, n_active_warps = (n_active + 31) >> 5;
  for (int i = threadIdx.x; i < kNumEdgeClass * D4; i += blockDim.x) {
    const int cls = i / D4, q = i - cls * D4;
    ee[i] = f4_add(ldg_f4(B1 + (size_t)(cls / 3) * D + 4 * q), ldg_f4(B2 + (size_t)(cls % 3) * D + 4 * q));
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTileStages; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, n_active_warps); }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  const int tid = threadIdx.x, lane = tid & 31;
  const int ntiles = (N + T - 1) / T;
  // tiles of this CTA: round-robin over the grid, or (blocked) one contiguous range per CTA
  const int per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
  const int t_first = blocked ? blockIdx.x * per_cta : blockIdx.x, t_step = blocked ? 1 : gridDim.x;
  const int t_end = blocked ? min(ntiles, t_first + per_cta) : ntiles;
  if (tid >= kTileConsumers) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      int k = 0;
      for (int t = t_first; t < t_end; t += t_step, ++k) {
        const int s = k % kTileStages;
        ptx::mbar_wait(empty_bar + s, ((k / kTileStages) & 1) ^ 1);
        const int t0 = t * T, rows = min(T, N - t0);
        ptx::mbar_arrive_expect_tx(full_bar + s, (uint32_t)rows * (uint32_t)(D * 4 + 32));
        ptx::bulk_load_1d(feat + (size_t)s * T * D4, src + (size_t)t0 * D, (uint32_t)rows * (uint32_t)(D * 4), full_bar + s);
        ptx::bulk_load_1d(nb + (size_t)s * T * 8, nbr + (size_t)t0 * 8, (uint32_t)rows * 32u, full_bar + s);
      }
    }
    return;
  }
  if ((tid & ~31) >= n_active) return;                               // warps without any (row slot, chunk)
  // ------------------------------------------------------------------ consumers
  const bool active = tid < n_active;
  const int slot = active ? tid / D4 : 0, q = active ? tid - slot * D4 : 0;
  const float4 sc = HAS_BN ? ldg_f4(coef + 4 * q) : f4_zero(), sh = HAS_BN ? ldg_f4(coef + D + 4 * q) : f4_zero();
  const float4* ee_q = ee + q;                                       // this thread's chunk of the 15 table rows
  const float4 ee_self = ee_q[kSelfLoopAttr * D4];
  auto act = [&](float4 v, int row) -> float4 {
    if (HAS_BN) {
      v = f4_fma2(v, sc, sh);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (DROP) v = f4_mul(v, drop_mask4(drop, row, q, D4));
    }
    return v;
  };
  int k = 0;
  for (int t = t_first; t < t_end; t += t_step, ++k) {
    const int s = k % kTileStages;
    const int t0 = t * T, rows = min(T, N - t0);
    const float4* ft_q = feat + (size_t)s * T * D4 + q;              // this thread's chunk of the staged rows
    const uint32_t* nt = nb + (size_t)s * T * 8;
    ptx::mbar_wait(full_bar + s, (k / kTileStages) & 1);
    if (active) {
      for (int r = slot; r < rows; r += R) {
        const int i = t0 + r;
        float4 acc = f4_zero();
        // one in-edge, packed (source << 4 | attr): acc += f(src[source]) + table[attr]; the source row comes from the staged
        // tile when it is one of this tile's rows, else (a molecule cut by the tile boundary) from global memory
        auto edge = [&](uint32_t w) {
          const int sidx = (int)(w >> 4);
          const unsigned rel = (unsigned)(sidx - t0);
          float4 v;
          if (rel < (unsigned)rows) v = ft_q[rel * D4];
          else v = ldg_f4(src + (size_t)sidx * D + 4 * q);
          acc = f4_add2(acc, f4_add2(act(v, sidx), ee_q[(w & 15u) * D4]));
        };
        const uint4 wa = *reinterpret_cast<const uint4*>(nt + r * 8);
        if (wa.x < kNbrLong) {                                       // entries in input order; kNbrEmpty ends the list
          edge(wa.x);
          if (wa.y < kNbrLong) {
            edge(wa.y);
            if (wa.z < kNbrLong) {
              edge(wa.z);
              if (wa.w < kNbrLong) {
                edge(wa.w);
                const uint4 wb = *reinterpret_cast<const uint4*>(nt + r * 8 + 4);
                const uint32_t w4[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  if (w4[kk] >= kNbrLong) break;
                  edge(w4[kk]);
                }
              }
            }
          }
        } else if (wa.x == kNbrLong) {                               // more than 8 in-edges: walk the CSR row
          const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
          for (int e = beg; e < end; ++e) edge(((uint32_t)__ldg(col + e) << 4) | (uint32_t)__ldg(eattr + e));
        }
        float4 rr = f4_add2(acc, f4_add2(act(ft_q[r * D4], i), ee_self));   // self loop LAST
        if (out_lo) st_f4(out_lo + (size_t)i * ld_out + 4 * q, f4_tf32_residual(rr));
        if (round_out) rr = f4_tf32(rr);
        if (store_cs) __stcs(reinterpret_cast<float4*>(out + (size_t)i * ld_out + 4 * q), rr);
        else st_f4(out + (size_t)i * ld_out + 4 * q, rr);
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(empty_bar + s);
  }
}

// Writes per-lane accumulators `acc[NV][NCH]` of every warp to shared memory and reduces them over
// the block's warps in warp order; block partial goes to partials[blockIdx.x][NV][D].
template <int NCH, int NV>
__device__ __forceinline__ void block_reduce_rows(float4 (&acc)[NV][NCH], float4* red /* [kRowWarps][NV][D4] */,
                                                  int D4, float* __restrict__ partials) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) red[(w * NV + v) * D4 + q] = acc[v][j];
    }
  __syncthreads();
  float* dst = partials + (size_t)blockIdx.x * NV * D4 * 4;
  for (int i = threadIdx.x; i < NV * D4; i += blockDim.x) {
    float4 s = red[i];
    for (int ww = 1; ww < kRowWarps; ++ww) s = f4_add(s, red[ww * NV * D4 + i]);
    st_f4(dst + 4 * i, s);
  }
}

// ------------------------------------------------------------------------------------------------
// GINE aggregation, backward (transposed gather over the source-sorted CSR):
//   g_h[j] = sum_{e: src_e = j} g_a[dst_e] + g_a[j]
// fused with the backward of the previous layer's ReLU and the BatchNorm batch statistics:
//   g_y[j] = g_h[j] * [z[j]*scale + shift > 0];  s1 += g_y;  s2 += g_y * xhat,  xhat = (z - mean) * invstd
// MODE 0: plain (layer 0: g_h feeds the node-embedding backward).  MODE 1: fused as above.
// ------------------------------------------------------------------------------------------------
// GATHER = false: no neighbour terms (g_h = g_a): the ReLU / BatchNorm-statistics stage alone, used by the GCN backward
// where the gradient arrives from a GEMM instead of an aggregation.
template <int NCH, int MODE, bool GATHER, bool DROP>
__global__ void __launch_bounds__(kRowThreads) gine_aggregate_bwd_kernel(
    const float* __restrict__ ga, const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ col_t,
    const float* __restrict__ z, const float* __restrict__ coef, int relu, int N, int D,
    float* __restrict__ gy, float* __restrict__ partials, int round_out, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* cf = sm4;                          // [4][D4] scale, shift, mean, invstd
  float4* red = sm4 + 4 * D4;                // [kRowWarps][2][D4]
  if (MODE == 1) {
    for (int i = threadIdx.x; i < 4 * D4; i += blockDim.x) cf[i] = ldg_f4(coef + 4 * i);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  float4 st[2][NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) { st[0][j] = f4_zero(); st[1][j] = f4_zero(); }
  CsrRowPrefetch pf;
  if (GATHER) pf.init(rowptr_t, col_t, nullptr, N, lane, warp, nwarps);
  for (int i = warp; i < N; i += nwarps) {
    const int beg = GATHER ? pf.b0 : 0, end = GATHER ? pf.e0 : 0;
    float4 acc[NCH], zz[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      zz[j] = (MODE == 1 && q < D4) ? ld_stream_f4(z + (size_t)i * D + 4 * q) : f4_zero();
      acc[j] = (q < D4) ? ldg_f4(ga + (size_t)i * D + 4 * q) : f4_zero();       // self loop
    }
    for (int e = beg; e < end; e += 2) {
      const bool two = (e + 1 < end);
      int d0, d1, unused;
      pf.get(e, d0, unused);
      if (two) pf.get(e + 1, d1, unused); else d1 = d0;
      float4 v0[NCH], v1[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int q = lane + 32 * j;
        v0[j] = (q < D4) ? ldg_f4(ga + (size_t)d0 * D + 4 * q) : f4_zero();
        v1[j] = (q < D4 && two) ? ldg_f4(ga + (size_t)d1 * D + 4 * q) : f4_zero();
      }
#pragma unroll
      for (int j = 0; j < NCH; ++j) acc[j] = f4_add(acc[j], f4_add(v0[j], v1[j]));
    }
    if (GATHER) pf.advance(i, nwarps);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        float4 r = acc[j];
        if (MODE == 1) {
          const float4 s = cf[q], b = cf[D4 + q], m = cf[2 * D4 + q], is = cf[3 * D4 + q], zv = zz[j];
          if (relu) {
            if (!(fmaf(zv.x, s.x, b.x) > 0.f)) r.x = 0.f;
            if (!(fmaf(zv.y, s.y, b.y) > 0.f)) r.y = 0.f;
            if (!(fmaf(zv.z, s.z, b.z) > 0.f)) r.z = 0.f;
            if (!(fmaf(zv.w, s.w, b.w) > 0.f)) r.w = 0.f;
          }
          if (DROP) r = f4_mul(r, drop_mask4(drop, i, q, D4));
          st[0][j] = f4_add(st[0][j], r);
          st[1][j].x = fmaf(r.x, (zv.x - m.x) * is.x, st[1][j].x);
          st[1][j].y = fmaf(r.y, (zv.y - m.y) * is.y, st[1][j].y);
          st[1][j].z = fmaf(r.z, (zv.z - m.z) * is.z, st[1][j].z);
          st[1][j].w = fmaf(r.w, (zv.w - m.w) * is.w, st[1][j].w);
        }
        st_f4(gy + (size_t)i * D + 4 * q, round_out ? f4_tf32(r) : r);
      }
    }
  }
  if (MODE == 1) block_reduce_rows<NCH, 2>(st, red, D4, partials);
}

// ------------------------------------------------------------------------------------------------
// GINE aggregation backward, shared-memory tile variant: the transposed gather over the same ring of staged tiles as the
// forward kernel (tiles of g_a rows + the fixed-width OUT-edge table nbr_t of the plan), thread = (row slot, float4 chunk).
// Sums run in the order of the warp-per-row kernel (self first, then out-edges in pairs), so gy is bitwise identical to it;
// the BatchNorm statistics partials are reduced over the row slots of a CTA in slot order (deterministic).
// ------------------------------------------------------------------------------------------------
template <int MODE, bool DROP>
__global__ void __launch_bounds__(kTileThreads, 1) gine_aggregate_bwd_tile_kernel(
    const float* __restrict__ ga, const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ col_t,
    const uint32_t* __restrict__ nbr_t, const float* __restrict__ z, const float* __restrict__ coef, int relu, int N, int D, int T,
    int n_stages, float* __restrict__ gy, float* __restrict__ partials, int round_out, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* feat = sm4;                                                // [stages][T][D4]   (reused for the final reduction)
  uint32_t* nb = reinterpret_cast<uint32_t*>(feat + (size_t)n_stages * T * D4);   // [stages][T][8]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(nb + (size_t)n_stages * T * 8);
  uint64_t* empty_bar = full_bar + kTileMaxStages;
  const int R = kTileConsumers / D4;
  const int n_active = R * D4, n_active_warps = (n_active + 31) >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, n_active_warps); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31;
  const int ntiles = (N + T - 1) / T;
  const bool active = tid < n_active;
  const int slot = active ? tid / D4 : 0, q = active ? tid - slot * D4 : 0;
  float4 s1 = f4_zero(), s2 = f4_zero();
  if (tid >= kTileConsumers) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      int k = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
        const int s = k % n_stages;
        ptx::mbar_wait(empty_bar + s, ((k / n_stages) & 1) ^ 1);
        const int t0 = t * T, rows = min(T, N - t0);
        ptx::mbar_arrive_expect_tx(full_bar + s, (uint32_t)rows * (uint32_t)(D * 4 + 32));
        ptx::bulk_load_1d(feat + (size_t)s * T * D4, ga + (size_t)t0 * D, (uint32_t)rows * (uint32_t)(D * 4), full_bar + s);
        ptx::bulk_load_1d(nb + (size_t)s * T * 8, nbr_t + (size_t)t0 * 8, (uint32_t)rows * 32u, full_bar + s);
      }
    }
  } else if ((tid & ~31) < n_active) {
    // ------------------------------------------------------------------ consumers
    float4 sc = f4_zero(), sh = f4_zero(), mean = f4_zero(), istd = f4_zero();
    if (MODE == 1) { sc = ldg_f4(coef + 4 * q); sh = ldg_f4(coef + D + 4 * q); mean = ldg_f4(coef + 2 * D + 4 * q); istd = ldg_f4(coef + 3 * D + 4 * q); }
    int k = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
      const int s = k % n_stages;
      const int t0 = t * T, rows = min(T, N - t0);
      const float4* ft_q = feat + (size_t)s * T * D4 + q;
      const uint32_t* nt = nb + (size_t)s * T * 8;
      ptx::mbar_wait(full_bar + s, (k / n_stages) & 1);
      if (active) {
        for (int r = slot; r < rows; r += R) {
          const int i = t0 + r;
          float4 zv = f4_zero();
          if (MODE == 1) zv = ld_stream_f4(z + (size_t)i * D + 4 * q);
          auto fetch = [&](uint32_t w) -> float4 {
            const int didx = (int)(w >> 4);
            const unsigned rel = (unsigned)(didx - t0);
            return rel < (unsigned)rows ? ft_q[rel * D4] : ldg_f4(ga + (size_t)didx * D + 4 * q);
          };
          float4 acc = ft_q[r * D4];                                     // self loop first (as the warp-per-row kernel)
          const uint4 wa = *reinterpret_cast<const uint4*>(nt + r * 8);
          if (wa.x < kNbrLong) {
            acc = f4_add2(acc, f4_add2(fetch(wa.x), wa.y < kNbrLong ? fetch(wa.y) : f4_zero()));
            if (wa.y < kNbrLong && wa.z < kNbrLong) {
              acc = f4_add2(acc, f4_add2(fetch(wa.z), wa.w < kNbrLong ? fetch(wa.w) : f4_zero()));
              if (wa.w < kNbrLong) {
                const uint4 wb = *reinterpret_cast<const uint4*>(nt + r * 8 + 4);
                if (wb.x < kNbrLong) {
                  acc = f4_add2(acc, f4_add2(fetch(wb.x), wb.y < kNbrLong ? fetch(wb.y) : f4_zero()));
                  if (wb.y < kNbrLong && wb.z < kNbrLong)
                    acc = f4_add2(acc, f4_add2(fetch(wb.z), wb.w < kNbrLong ? fetch(wb.w) : f4_zero()));
                }
              }
            }
          } else if (wa.x == kNbrLong) {                                 // more than 8 out-edges: walk the CSR row, in pairs
            const int beg = __ldg(rowptr_t + i), end = __ldg(rowptr_t + i + 1);
            for (int e = beg; e < end; e += 2) {
              const float4 v0 = fetch((uint32_t)__ldg(col_t + e) << 4);
              const float4 v1 = (e + 1 < end) ? fetch((uint32_t)__ldg(col_t + e + 1) << 4) : f4_zero();
              acc = f4_add2(acc, f4_add2(v0, v1));
            }
          }
          float4 rr = acc;
          if (MODE == 1) {
            if (relu) {
              if (!(fmaf(zv.x, sc.x, sh.x) > 0.f)) rr.x = 0.f;
              if (!(fmaf(zv.y, sc.y, sh.y) > 0.f)) rr.y = 0.f;
              if (!(fmaf(zv.z, sc.z, sh.z) > 0.f)) rr.z = 0.f;
              if (!(fmaf(zv.w, sc.w, sh.w) > 0.f)) rr.w = 0.f;
            }
            if (DROP) rr = f4_mul(rr, drop_mask4(drop, i, q, D4));
            s1 = f4_add(s1, rr);
            s2.x = fmaf(rr.x, (zv.x - mean.x) * istd.x, s2.x); s2.y = fmaf(rr.y, (zv.y - mean.y) * istd.y, s2.y);
            s2.z = fmaf(rr.z, (zv.z - mean.z) * istd.z, s2.z); s2.w = fmaf(rr.w, (zv.w - mean.w) * istd.w, s2.w);
          }
          st_f4(gy + (size_t)i * D + 4 * q, round_out ? f4_tf32(rr) : rr);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(empty_bar + s);
    }
  }
  if (MODE == 1) {
    // block partial of the BatchNorm statistics: row slots summed in slot order, through the (now idle) ring memory
    __syncthreads();
    float4* red = feat;                                               // [R][2][D4]
    if (active) { red[(slot * 2 + 0) * D4 + q] = s1; red[(slot * 2 + 1) * D4 + q] = s2; }
    __syncthreads();
    for (int e = tid; e < 2 * D4; e += blockDim.x) {
      float4 a = red[e];
      for (int sl = 1; sl < R; ++sl) a = f4_add(a, red[sl * 2 * D4 + e]);
      st_f4(partials + (size_t)blockIdx.x * 2 * D + 4 * e, a);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Deterministic reduction of partial rows:  out[c] (+)= scale * sum_p partials[p][c], p in order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partials, int P, int len,
                                                               float scale, int accumulate, float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < len)
    for (int p = threadIdx.y; p < P; p += 32) s += partials[(size_t)p * len + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < len) {
    float t = red[0][threadIdx.x];
    for (int k = 1; k < 32; ++k) t += red[k][threadIdx.x];
    t *= scale;
    out[c] = accumulate ? out[c] + t : t;
  }
}

// The same for many partial rows and few columns: 19 CTAs of the kernel above cannot pull 7.7 MB fast enough.  Fixed order:
// lane y sums p = y, y + 128, ...; the 128 lane sums are added in lane order.
__global__ void __launch_bounds__(1024) reduce_partials_tall_kernel(const float* __restrict__ partials, int P, int len,
                                                                    float scale, int accumulate, float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[128][9];
  const int c = blockIdx.x * 8 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  if (c < len) {
    int p = threadIdx.y;
    for (; p + 128 < P; p += 256) { s0 += partials[(size_t)p * len + c]; s1 += partials[(size_t)(p + 128) * len + c]; }
    if (p < P) s0 += partials[(size_t)p * len + c];
  }
  red[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < len) {
    float t = red[0][threadIdx.x];
    for (int k = 1; k < 128; ++k) t += red[k][threadIdx.x];
    t *= scale;
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d, training mode (ginet_molclr.py:107; torch.nn.BatchNorm1d defaults eps=1e-5,
// momentum=0.1).  The producing GEMM epilogue leaves per-128-row-tile column (mean, M2); this
// merges them (Chan et al.) into batch mean / biased variance, writes the per-feature
// coefficients the consumers use, and updates the running statistics exactly once.
//   coef[0]=scale=gamma*invstd  coef[1]=shift=beta-mean*scale  coef[2]=mean  coef[3]=invstd
// ------------------------------------------------------------------------------------------------
constexpr int kBnSplits = 32;   // first-level partial merges (one block row each) == lanes of the second-level warp

// Chan et al. merge of (n, mean, M2) pairs.  Counts are small exact integers, so the ratio nb/tot is formed in fp32 (the
// double-precision divide is by far the slowest instruction here); everything else stays in double.
struct BnAcc { double n, mean, m2; };
__device__ __forceinline__ void bn_merge(BnAcc& a, double nb, double mb, double qb) {
  if (nb == 0.0) return;
  const double tot = a.n + nb, delta = mb - a.mean;
  const double w = (double)((float)nb / (float)tot);
  a.mean += delta * w;
  a.m2 += qb + delta * delta * (a.n * w);
  a.n = tot;
}

// Level 1: block (x = 32 columns, y = 16) of split s merges the tiles [s*per, (s+1)*per) into one (n, mean, M2)
// per column, kept in double: ws[s][3][D].
__global__ void __launch_bounds__(512) bn_merge_tiles_kernel(
    const float* __restrict__ tile_stats /* [T][2][D] */, int T, int tile_rows, int N, int D, int per,
    double* __restrict__ ws) {
  pdl_sync();
  __shared__ double s_n[16][33], s_mean[16][33], s_m2[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int t0 = blockIdx.y * per, t1 = min(T, t0 + per);
  BnAcc a = {0.0, 0.0, 0.0};
  if (c < D) {
    for (int t = t0 + threadIdx.y; t < t1; t += 16) {
      const int rows = min(tile_rows, N - t * tile_rows);
      if (rows <= 0) continue;                       // padding groups of the last 128-row GEMM tile
      bn_merge(a, rows, tile_stats[((size_t)t * 2) * D + c], tile_stats[((size_t)t * 2 + 1) * D + c]);
    }
  }
  s_n[threadIdx.y][threadIdx.x] = a.n; s_mean[threadIdx.y][threadIdx.x] = a.mean; s_m2[threadIdx.y][threadIdx.x] = a.m2;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    for (int k = 1; k < 16; ++k) bn_merge(a, s_n[k][threadIdx.x], s_mean[k][threadIdx.x], s_m2[k][threadIdx.x]);
    double* w = ws + (size_t)blockIdx.y * 3 * D;
    w[c] = a.n; w[D + c] = a.mean; w[2 * D + c] = a.m2;
  }
}

// Level 2: one warp per column; lane s holds partial s, merged by a fixed shuffle tree (deterministic); lane 0 writes the
// coefficients and updates the running statistics.
__global__ void __launch_bounds__(256) bn_fwd_finalize_kernel(
    const double* __restrict__ ws /* [S][3][D] */, int S, int D,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
    long long* num_batches_tracked, float momentum, float eps, float* __restrict__ coef) {
  pdl_sync();
  const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c < D) {
    BnAcc a = {0.0, 0.0, 0.0};
    if (lane < S) { a.n = ws[(size_t)lane * 3 * D + c]; a.mean = ws[(size_t)lane * 3 * D + D + c]; a.m2 = ws[(size_t)lane * 3 * D + 2 * D + c]; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double nb = __shfl_down_sync(0xffffffffu, a.n, o), mb = __shfl_down_sync(0xffffffffu, a.mean, o),
                   qb = __shfl_down_sync(0xffffffffu, a.m2, o);
      if ((lane & (2 * o - 1)) == 0) bn_merge(a, nb, mb, qb);
    }
    if (lane == 0) {
      const double var = (a.n > 0.0) ? a.m2 / a.n : 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      const float scale = gamma[c] * invstd;
      coef[c] = scale;
      coef[D + c] = beta[c] - (float)a.mean * scale;
      coef[2 * D + c] = (float)a.mean;
      coef[3 * D + c] = invstd;
      if (running_mean) {
        const double unbiased = (a.n > 1.0) ? a.m2 / (a.n - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)a.mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
}

// Eval mode (molclr.py:163): coefficients from the running statistics.
__global__ void bn_eval_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                    float eps, int D, float* __restrict__ coef) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const float is = 1.0f / sqrtf(running_var[c] + eps);
  const float scale = gamma[c] * is;
  coef[c] = scale;
  coef[D + c] = beta[c] - running_mean[c] * scale;
  coef[2 * D + c] = running_mean[c];
  coef[3 * D + c] = is;
}

// BatchNorm backward, step 1: reduce the (s1, s2) partials and emit
//   dgamma = s2, dbeta = s1, and the apply coefficients so that  g_z = k1*g_y + A + B*z :
//   k1 = gamma*invstd,  A = -k1*s1/N + k1*(s2/N)*invstd*mean,  B = -k1*(s2/N)*invstd
// In eval mode (use_batch_stats=0) BN is a fixed affine map: g_z = k1*g_y.
__global__ void __launch_bounds__(512) bn_bwd_finalize_kernel(
    const float* __restrict__ partials /* [P][2][D] */, int P, int N, int D, const float* __restrict__ gamma,
    const float* __restrict__ coef, int use_batch_stats, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ bcoef /* [3][D] */) {
  pdl_sync();
  __shared__ float r1[16][33], r2[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  if (c < D)
    for (int p = threadIdx.y; p < P; p += 16) { s1 += partials[((size_t)p * 2) * D + c]; s2 += partials[((size_t)p * 2 + 1) * D + c]; }
  r1[threadIdx.y][threadIdx.x] = s1; r2[threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    for (int k = 1; k < 16; ++k) { s1 += r1[k][threadIdx.x]; s2 += r2[k][threadIdx.x]; }
    dgamma[c] = s2; dbeta[c] = s1;
    const float mean = coef[2 * D + c], invstd = coef[3 * D + c];
    const float k1 = gamma[c] * invstd;
    if (use_batch_stats) {
      const float c1 = s1 / (float)N, c2 = s2 / (float)N;
      bcoef[c] = k1;
      bcoef[D + c] = -k1 * c1 + k1 * c2 * invstd * mean;
      bcoef[2 * D + c] = -k1 * c2 * invstd;
    } else {
      bcoef[c] = k1; bcoef[D + c] = 0.f; bcoef[2 * D + c] = 0.f;
    }
  }
}

// BatchNorm backward, step 2 (elementwise): g_z = k1*g_y + A + B*z, written tf32-rounded because g_z
// is only ever a tensor-core operand; also leaves per-block column sums of g_z (the bias gradient
// of the Linear in front of the BN).  SRC 0: g_y from memory.  SRC 1: g_y[n] = g_p[graph(n)] * w[graph(n)]
// (the backward of global_mean/add_pool feeding the last layer's BN, ginet_molclr.py:113).
template <int NCH, int SRC>
__global__ void __launch_bounds__(kRowThreads) bn_bwd_apply_kernel(
    const float* __restrict__ gy, const float* __restrict__ gp, const int32_t* __restrict__ node2graph,
    const int32_t* __restrict__ gptr, int pool_mode, const int32_t* __restrict__ arg, const float* __restrict__ z,
    const float* __restrict__ bcoef, int N, int D, float* __restrict__ gz, long long ld_gz, float* __restrict__ partials,
    int round_out, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* cf = sm4;                 // [3][D4]
  float4* red = sm4 + 3 * D4;       // [kRowWarps][1][D4]
  for (int i = threadIdx.x; i < 3 * D4; i += blockDim.x) cf[i] = ldg_f4(bcoef + 4 * i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  float4 st[1][NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) st[0][j] = f4_zero();
  for (int i = warp; i < N; i += nwarps) {
    float w = 1.f;
    const float* grow;
    const int32_t* arow = nullptr;
    if (SRC == 1) {
      const int g = __ldg(node2graph + i);
      if (pool_mode == 0) w = 1.f / (float)max(__ldg(gptr + g + 1) - __ldg(gptr + g), 1);
      grow = gp + (size_t)g * D;
      if (pool_mode == 2) arow = arg + (size_t)g * D;
    } else {
      grow = gy + (size_t)i * D;
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        float4 g = (SRC == 1) ? ldg_f4(grow + 4 * q) : ld_stream_f4(grow + 4 * q);
        const float4 zv = ld_stream_f4(z + (size_t)i * D + 4 * q);
        const float4 k1 = cf[q], A = cf[D4 + q], B = cf[2 * D4 + q];
        if (SRC == 1) {
          g.x *= w; g.y *= w; g.z *= w; g.w *= w;
          if (arow) {                                  // max pooling: the gradient goes to the arg-max node only
            const int4 am = __ldg(reinterpret_cast<const int4*>(arow + 4 * q));
            g.x = am.x == i ? g.x : 0.f; g.y = am.y == i ? g.y : 0.f; g.z = am.z == i ? g.z : 0.f; g.w = am.w == i ? g.w : 0.f;
          }
          if (drop.thr) g = f4_mul(g, drop_mask4(drop, i, q, D4));
        }
        float4 r;
        r.x = fmaf(k1.x, g.x, fmaf(B.x, zv.x, A.x)); r.y = fmaf(k1.y, g.y, fmaf(B.y, zv.y, A.y));
        r.z = fmaf(k1.z, g.z, fmaf(B.z, zv.z, A.z)); r.w = fmaf(k1.w, g.w, fmaf(B.w, zv.w, A.w));
        st[0][j] = f4_add(st[0][j], r);
        st_f4(gz + (size_t)i * ld_gz + 4 * q, round_out ? f4_tf32(r) : r);
      }
    }
  }
  block_reduce_rows<NCH, 1>(st, red, D4, partials);
}

// ------------------------------------------------------------------------------------------------
// global_mean_pool / global_add_pool (ginet_molclr.py:83-88,113) as a segmented reduction fused with
// the last layer's BatchNorm apply:  p[g] = w_g * sum_{n in graph g, node order} (z[n]*scale + shift)
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kRowThreads) pool_fwd_kernel(
    const float* __restrict__ z, const float* __restrict__ coef, int relu, const int32_t* __restrict__ gptr,
    const int32_t* __restrict__ gperm, int pool_mode, int G, int D, float* __restrict__ out, long long ld_out, int round_out,
    float* __restrict__ out_lo, int32_t* __restrict__ arg, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* sc = sm4; float4* sh = sm4 + D4;
  for (int i = threadIdx.x; i < D4; i += blockDim.x) { sc[i] = ldg_f4(coef + 4 * i); sh[i] = ldg_f4(coef + D + 4 * i); }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  for (int g = warp; g < G; g += nwarps) {
    const int beg = __ldg(gptr + g), end = __ldg(gptr + g + 1);
    float4 acc[NCH];
    int4 am[NCH];
    const float init = (pool_mode == 2 && end > beg) ? -INFINITY : 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) { acc[j] = make_float4(init, init, init, init); am[j] = make_int4(-1, -1, -1, -1); }
    auto consume = [&](float4 v, int n, int q, int j) {
      const float4 s = sc[q], b = sh[q];
      v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y); v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (drop.thr) v = f4_mul(v, drop_mask4(drop, n, q, D4));
      if (pool_mode == 2) {                             // global_max_pool: first maximum wins
        if (v.x > acc[j].x) { acc[j].x = v.x; am[j].x = n; }
        if (v.y > acc[j].y) { acc[j].y = v.y; am[j].y = n; }
        if (v.z > acc[j].z) { acc[j].z = v.z; am[j].z = n; }
        if (v.w > acc[j].w) { acc[j].w = v.w; am[j].w = n; }
      } else {
        acc[j] = f4_add(acc[j], v);
      }
    };
    for (int p = beg; p < end; p += 4) {                  // four node rows in flight; consumed in node order
      int n[4];
      float4 v[4][NCH];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        n[u] = (p + u < end) ? __ldg(gperm + p + u) : -1;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int q = lane + 32 * j;
          v[u][j] = (q < D4 && n[u] >= 0) ? ld_stream_f4(z + (size_t)n[u] * D + 4 * q) : f4_zero();
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (n[u] >= 0) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            const int q = lane + 32 * j;
            if (q < D4) consume(v[u][j], n[u], q, j);
          }
        }
    }
    const bool pool_mean = pool_mode == 0;
    const float cntf = (float)max(end - beg, 1);             // count.clamp(min=1)
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        float4 r = acc[j];
        if (pool_mean) { r.x /= cntf; r.y /= cntf; r.z /= cntf; r.w /= cntf; }
        if (pool_mode == 2 && arg) *reinterpret_cast<int4*>(arg + (size_t)g * D + 4 * q) = am[j];
        if (out_lo) st_f4(out_lo + (size_t)g * ld_out + 4 * q, f4_tf32_residual(r));
        if (round_out) r = f4_tf32(r);
        st_f4(out + (size_t)g * ld_out + 4 * q, r);
      }
    }
  }
}

// Backward of the pool into the last layer's BatchNorm statistics:
//   g_y[n] = g_p[graph(n)] * w,  s1 += g_y,  s2 += g_y * xhat[n]      (g_y is never materialised)
template <int NCH>
__global__ void __launch_bounds__(kRowThreads) pool_bwd_stats_kernel(
    const float* __restrict__ gp, const int32_t* __restrict__ node2graph, const int32_t* __restrict__ gptr,
    int pool_mode, const int32_t* __restrict__ arg, const float* __restrict__ z, const float* __restrict__ coef, int N, int D,
    float* __restrict__ partials, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* cf = sm4;                 // mean, invstd
  float4* red = sm4 + 2 * D4;
  for (int i = threadIdx.x; i < 2 * D4; i += blockDim.x) cf[i] = ldg_f4(coef + 2 * D + 4 * i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  float4 st[2][NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) { st[0][j] = f4_zero(); st[1][j] = f4_zero(); }
  for (int i = warp; i < N; i += nwarps) {
    const int g = __ldg(node2graph + i);
    const float w = pool_mode == 0 ? 1.f / (float)max(__ldg(gptr + g + 1) - __ldg(gptr + g), 1) : 1.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        float4 r = ldg_f4(gp + (size_t)g * D + 4 * q);
        const float4 zv = ld_stream_f4(z + (size_t)i * D + 4 * q), m = cf[q], is = cf[D4 + q];
        r.x *= w; r.y *= w; r.z *= w; r.w *= w;
        if (pool_mode == 2) {
          const int4 am = __ldg(reinterpret_cast<const int4*>(arg + (size_t)g * D + 4 * q));
          r.x = am.x == i ? r.x : 0.f; r.y = am.y == i ? r.y : 0.f; r.z = am.z == i ? r.z : 0.f; r.w = am.w == i ? r.w : 0.f;
        }
        if (drop.thr) r = f4_mul(r, drop_mask4(drop, i, q, D4));
        st[0][j] = f4_add(st[0][j], r);
        st[1][j].x = fmaf(r.x, (zv.x - m.x) * is.x, st[1][j].x); st[1][j].y = fmaf(r.y, (zv.y - m.y) * is.y, st[1][j].y);
        st[1][j].z = fmaf(r.z, (zv.z - m.z) * is.z, st[1][j].z); st[1][j].w = fmaf(r.w, (zv.w - m.w) * is.w, st[1][j].w);
      }
    }
  }
  block_reduce_rows<NCH, 2>(st, red, D4, partials);
}

// ------------------------------------------------------------------------------------------------
// GCN path helpers (gcn_molclr.py:62-84: the GEMM comes BEFORE the aggregation, so its A operand -- the previous
// layer's BatchNorm + ReLU output, gcn_molclr.py:146-152 -- has to exist in memory, and the BatchNorm statistics of
// the aggregated output cannot come from a GEMM epilogue).
// ------------------------------------------------------------------------------------------------
// x = [relu](z*scale + shift) (coef == NULL: x = z), written as a tensor-core operand: hi = tf32(x), lo = tf32(x - hi).
template <int NCH>
__global__ void __launch_bounds__(kRowThreads) bn_apply_fwd_kernel(
    const float* __restrict__ z, const float* __restrict__ coef, int relu, int N, int D, float* __restrict__ hi,
    float* __restrict__ lo, long long ld, int round_hi, const DropCfg drop) {
  pdl_sync();
  extern __shared__ float4 sm4[];
  const int D4 = D >> 2;
  float4* sc = sm4; float4* sh = sm4 + D4;
  if (coef) {
    for (int i = threadIdx.x; i < D4; i += blockDim.x) { sc[i] = ldg_f4(coef + 4 * i); sh[i] = ldg_f4(coef + D + 4 * i); }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  for (int i = warp; i < N; i += nwarps) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        float4 v = ld_stream_f4(z + (size_t)i * D + 4 * q);
        if (coef) {
          const float4 s = sc[q], b = sh[q];
          v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y); v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if (drop.thr) v = f4_mul(v, drop_mask4(drop, i, q, D4));
        }
        st_f4(hi + (size_t)i * ld + 4 * q, round_hi ? f4_tf32(v) : v);
        if (lo) st_f4(lo + (size_t)i * ld + 4 * q, f4_tf32_residual(v));
      }
    }
  }
}

// Column (mean, M2) of every 32-row group of z: the same [T][2][D] tile statistics the GEMM epilogue emits, so that
// molclr_bn_fwd_finalize serves both encoders.  One warp per group; the second pass re-reads the group from L1/L2.
template <int NCH>
__global__ void __launch_bounds__(kRowThreads) bn_tile_stats_kernel(const float* __restrict__ z, int N, int D, int T,
                                                                    float* __restrict__ tile_stats) {
  pdl_sync();
  const int D4 = D >> 2, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kRowWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kRowWarps;
  for (int t = warp; t < T; t += nwarps) {
    const int r0 = t * 32, rows = max(0, min(32, N - r0));
    float4 sum[NCH], m2[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) { sum[j] = f4_zero(); m2[j] = f4_zero(); }
    for (int r = 0; r < rows; ++r)
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int q = lane + 32 * j;
        if (q < D4) sum[j] = f4_add(sum[j], ldg_f4(z + (size_t)(r0 + r) * D + 4 * q));
      }
    const float inv = rows > 0 ? 1.f / (float)rows : 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) { sum[j].x *= inv; sum[j].y *= inv; sum[j].z *= inv; sum[j].w *= inv; }
    for (int r = 0; r < rows; ++r)
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int q = lane + 32 * j;
        if (q < D4) {
          const float4 v = ldg_f4(z + (size_t)(r0 + r) * D + 4 * q);
          const float dx = v.x - sum[j].x, dy = v.y - sum[j].y, dz = v.z - sum[j].z, dw = v.w - sum[j].w;
          m2[j].x = fmaf(dx, dx, m2[j].x); m2[j].y = fmaf(dy, dy, m2[j].y); m2[j].z = fmaf(dz, dz, m2[j].z); m2[j].w = fmaf(dw, dw, m2[j].w);
        }
      }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int q = lane + 32 * j;
      if (q < D4) {
        st_f4(tile_stats + ((size_t)t * 2) * D + 4 * q, sum[j]);
        st_f4(tile_stats + ((size_t)t * 2 + 1) * D + 4 * q, m2[j]);
      }
    }
  }
}

// out[r] = sum_c in[r][c] (one warp per row, fixed order): collapses the [8][D] bond-table gradient to the GCN's [8][1].
__global__ void row_sum_kernel(const float* __restrict__ in, int R, int C, float* __restrict__ out) {
  pdl_sync();
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += in[(size_t)r * C + c];
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

// ------------------------------------------------------------------------------------------------
// Activations of the fine-tune prediction head (ginet_finetune.py:96-127): mode 0 = Softplus (torch defaults beta 1,
// threshold 20), mode 1 = ReLU.  Forward writes the result as a tensor-core operand pair; backward writes the rounded
// gradient (it only ever feeds GEMMs).
// ------------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float* __restrict__ x, int mode, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  pdl_sync();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float y = mode == 0 ? (v > 20.f ? v : log1pf(expf(v))) : fmaxf(v, 0.f);
    const float h = round_tf32(y);
    hi[i] = h;
    if (lo) lo[i] = round_tf32(y - h);
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, int mode, int64_t n, float* __restrict__ gx) {
  pdl_sync();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float d = mode == 0 ? (v > 20.f ? 1.f : 1.f / (1.f + expf(-v))) : (v > 0.f ? 1.f : 0.f);
    gx[i] = round_tf32(gy[i] * d);
  }
}

// The dropout mask itself (0 or 1/(1-p)) as a matrix: what the fused consumers apply.  Test / diagnostics hook.
__global__ void dropout_mask_kernel(int N, int D, float* __restrict__ out, const DropCfg drop) {
  pdl_sync();
  const int D4 = D >> 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)N * D4; t += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(t / D4), q = (int)(t - (long long)row * D4);
    st_f4(out + (size_t)row * D + 4 * q, drop.thr ? drop_mask4(drop, row, q, D4) : make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

// ------------------------------------------------------------------------------------------------
// Small elementwise helpers
// ------------------------------------------------------------------------------------------------
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n4, int64_t n) {
  pdl_sync();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = i; k < n4; k += stride) st_f4(y + 4 * k, f4_add(*reinterpret_cast<const float4*>(y + 4 * k), ldg_f4(x + 4 * k)));
  for (int64_t k = 4 * n4 + i; k < n; k += stride) y[k] += x[k];
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  pdl_sync();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float v = src[i], h = round_tf32(v);
    hi[i] = h;
    if (lo) lo[i] = round_tf32(v - h);
  }
}

__global__ void round_tf32_2d_kernel(const float* __restrict__ src, long long ld_src, float* __restrict__ hi, float* __restrict__ lo,
                                     long long ld_dst, int rows, int cols) {
  pdl_sync();
  const long long n = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float v = src[r * ld_src + c], h = round_tf32(v);
    hi[r * ld_dst + c] = h;
    if (lo) lo[r * ld_dst + c] = round_tf32(v - h);
  }
}

// F.normalize(z, dim=1) (molclr.py:63-64; eps = 1e-12): y = z / max(||z||, eps).  One warp per row.
__global__ void l2_normalize_fwd_kernel(const float* __restrict__ z, int R, int C, float eps, float* __restrict__ y,
                                        float* __restrict__ inv_norm) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = z[(size_t)row * C + c]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  const float inv = 1.f / fmaxf(sqrtf(s), eps);
  for (int c = lane; c < C; c += 32) y[(size_t)row * C + c] = z[(size_t)row * C + c] * inv;
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// The row preparation of NTXentLoss.forward in one pass (nt_xent.py:48 + :40-45): rep = cat([zjs, zis]) -- rows [0, RA) from zA,
// [RA, RA + RB) from zB -- optionally divided by max(||row||, eps) (torch.nn.CosineSimilarity's normalisation), written
// unrounded (y, for the backward) and tf32-rounded (y_r, the tensor-core operand).
__global__ void l2_normalize_cat_fwd_kernel(const float* __restrict__ zA, const float* __restrict__ zB, int RA, int RB, int C, float eps,
                                            int normalise, float* __restrict__ y, float* __restrict__ y_r, float* __restrict__ inv_norm,
                                            __half* __restrict__ y16, int ld16) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= RA + RB) return;
  const float* z = row < RA ? zA + (size_t)row * C : zB + (size_t)(row - RA) * C;
  float inv = 1.f;
  if (normalise) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = z[c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    inv = 1.f / fmaxf(sqrtf(s), eps);
  }
  for (int c = lane; c < C; c += 32) {
    const float v = z[c] * inv;
    if (y) y[(size_t)row * C + c] = v;
    if (y_r) y_r[(size_t)row * C + c] = round_tf32(v);
    if (y16) y16[(size_t)row * ld16 + c] = __float2half_rn(v);
  }
  if (y16) for (int c = C + lane; c < ld16; c += 32) y16[(size_t)row * ld16 + c] = __float2half_rn(0.f);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// backward: g_z = (g_y - y * <g_y, y>) * inv_norm          (rows with ||z|| < eps: g_z = g_y * inv_norm)
// gscale (optional, device scalar): g_y is multiplied by *gscale first (the incoming gradient of a scalar loss).
__global__ void l2_normalize_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                        const float* __restrict__ inv_norm, int R, int C, float eps, const float* __restrict__ gscale,
                                        float* __restrict__ gz) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float gs = gscale ? __ldg(gscale) : 1.f;
  float d = 0.f;
  for (int c = lane; c < C; c += 32) d = fmaf(gy[(size_t)row * C + c] * gs, y[(size_t)row * C + c], d);
  d = warp_sum(d);
  const float inv = inv_norm[row];
  if (inv >= 1.f / eps) d = 0.f;          // clamped norm: y = z / eps is linear in z
  for (int c = lane; c < C; c += 32) gz[(size_t)row * C + c] = (gy[(size_t)row * C + c] * gs - y[(size_t)row * C + c] * d) * inv;
}

}  // namespace molclr

// ================================================================================================
// C ABI
// ================================================================================================
using namespace molclr;

#define REQUIRE_D(D) MOLCLR_REQUIRE((D) > 0 && (D) % 4 == 0 && (D) <= 512, "feature width D=%d must be a multiple of 4, <= 512", (int)(D))

extern "C" int molclr_embed_nodes_fwd(const int32_t* xpacked, const float* E1, const float* E2, int64_t N, int D,
                                      float* out, cudaStream_t stream) {
  REQUIRE_D(D);
  if (N == 0) return 0;
  NCH_DISPATCH(D / 4, {
    auto k = embed_nodes_fwd_kernel<NCH>;
    MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, 0, kRowWarps, N), kRowThreads, 0, stream, xpacked, E1, E2, (int)N, D, out);
  });
  MOLCLR_CHECK_LAUNCH("embed_nodes_fwd");
  return 0;
}

extern "C" int molclr_reduce_partials(const float* partials, int P, int len, float scale, int accumulate, float* out,
                                      cudaStream_t stream) {
  if (len == 0) return 0;
  if (P >= 1024 && len <= 2048) {      // many partial rows, few columns (bias gradients from per-32-row column sums): 8 columns x 128 lanes per CTA
    MOLCLR_LAUNCH(reduce_partials_tall_kernel, (len + 7) / 8, dim3(8, 128), 0, stream, partials, P, len, scale, accumulate, out);
    MOLCLR_CHECK_LAUNCH("reduce_partials");
    return 0;
  }
  MOLCLR_LAUNCH(reduce_partials_kernel, (len + 31) / 32, dim3(32, 32), 0, stream, partials, P, len, scale, accumulate, out);
  MOLCLR_CHECK_LAUNCH("reduce_partials");
  return 0;
}

template <bool HAS_BN, bool DROP>
static void launch_fwd_tile(const float* src, const float* bn_coef, int relu, const int32_t* rowptr, const int32_t* col,
                            const uint8_t* eattr, const uint32_t* nbr, const float* B1, const float* B2, int64_t N, int D, int T,
                            float* out, int64_t ld_out, int round_out, float* out_lo, const DropCfg& drop, cudaStream_t stream) {
  auto k = gine_aggregate_fwd_tile_kernel<HAS_BN, DROP>;
  const int stages = g_tile_stages < 2 ? 2 : (g_tile_stages > kTileMaxStages ? kTileMaxStages : g_tile_stages);
  const size_t smem = aggregate_tile_smem(D, T, stages);
  static bool attr_set = false;                      // per instantiation
  if (!attr_set) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = true; }
  const int64_t ntiles = (N + T - 1) / T;
  const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  MOLCLR_LAUNCH(k, grid, kTileThreads, smem, stream, src, bn_coef, relu, rowptr, col, eattr, nbr, B1, B2, (int)N, D, T, stages, g_tile_store_cs, g_tile_blocked, out,
                ld_out, round_out, out_lo, drop);
}

static int aggregate_fwd_launch(const float* src, const float* bn_coef, int relu, const int32_t* rowptr, const int32_t* col,
                                const uint8_t* eattr, const uint32_t* nbr, const float* B1, const float* B2, const float* bias, bool scalar,
                                int64_t N, int D, float* out, int64_t ld_out, int round_tf32_out, float* out_lo, const DropCfg drop,
                                cudaStream_t stream) {
  REQUIRE_D(D);
  MOLCLR_REQUIRE(ld_out >= D && ld_out % 4 == 0, "aggregate_fwd: ld_out must be >= D and a multiple of 4");
  if (N == 0) return 0;
  if (nbr && !scalar && g_aggregate_tile) {
    // shared-memory tile path (see gine_aggregate_fwd_tile_kernel)
    MOLCLR_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(nbr) & 15) == 0,
                   "aggregate_fwd: src and nbr must be 16-byte aligned");
    const int stages = g_tile_stages < 2 ? 2 : (g_tile_stages > kTileMaxStages ? kTileMaxStages : g_tile_stages);
    const int T = aggregate_tile_rows(D, stages);
    MOLCLR_REQUIRE(aggregate_tile_smem(D, T, stages) <= 227 * 1024, "aggregate_fwd: tile of %d rows x %d features does not fit shared memory", T, D);
    if (bn_coef && drop.thr)
      launch_fwd_tile<true, true>(src, bn_coef, relu, rowptr, col, eattr, nbr, B1, B2, N, D, T, out, ld_out, round_tf32_out, out_lo, drop, stream);
    else if (bn_coef)
      launch_fwd_tile<true, false>(src, bn_coef, relu, rowptr, col, eattr, nbr, B1, B2, N, D, T, out, ld_out, round_tf32_out, out_lo, drop, stream);
    else
      launch_fwd_tile<false, false>(src, nullptr, 0, rowptr, col, eattr, nbr, B1, B2, N, D, T, out, ld_out, round_tf32_out, out_lo, drop, stream);
    MOLCLR_CHECK_LAUNCH("aggregate_fwd_tile");
    return 0;
  }
  const size_t smem = (size_t)(kNumEdgeClass + 2) * D * sizeof(float);
  NCH_DISPATCH(D / 4, {
    if (scalar) {
      auto k = gine_aggregate_fwd_kernel<NCH, false, true, false>;
      MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, smem, kRowWarps, N), kRowThreads, smem, stream,
                    src, nullptr, 0, rowptr, col, eattr, B1, B2, bias, (int)N, D, out, ld_out, round_tf32_out, out_lo, drop);
    } else if (bn_coef && drop.thr) {
      auto k = gine_aggregate_fwd_kernel<NCH, true, false, true>;
      MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, smem, kRowWarps, N), kRowThreads, smem, stream,
                    src, bn_coef, relu, rowptr, col, eattr, B1, B2, nullptr, (int)N, D, out, ld_out, round_tf32_out, out_lo, drop);
    } else if (bn_coef) {
      auto k = gine_aggregate_fwd_kernel<NCH, true, false, false>;
      MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, smem, kRowWarps, N), kRowThreads, smem, stream,
                    src, bn_coef, relu, rowptr, col, eattr, B1, B2, nullptr, (int)N, D, out, ld_out, round_tf32_out, out_lo, drop);
    } else {
      auto k = gine_aggregate_fwd_kernel<NCH, false, false, false>;
      MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, smem, kRowWarps, N), kRowThreads, smem, stream,
                    src, nullptr, 0, rowptr, col, eattr, B1, B2, nullptr, (int)N, D, out, ld_out, round_tf32_out, out_lo, drop);
    }
  });
  MOLCLR_CHECK_LAUNCH("aggregate_fwd");
  return 0;
}

extern "C" int molclr_gine_aggregate_fwd(const float* src, const float* bn_coef, int relu, const int32_t* rowptr,
                                         const int32_t* col, const uint8_t* eattr, const uint32_t* nbr, const float* B1,
                                         const float* B2, int64_t N, int D, float* out, int64_t ld_out, int round_tf32_out,
                                         float* out_lo, uint32_t drop_seed, float drop_p, cudaStream_t stream) {
  MOLCLR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability %f not in [0, 1)", drop_p);
  return aggregate_fwd_launch(src, bn_coef, relu, rowptr, col, eattr, nbr, B1, B2, nullptr, false, N, D, out, ld_out, round_tf32_out,
                              out_lo, make_drop(drop_seed, bn_coef ? drop_p : 0.f), stream);
}

extern "C" int molclr_gcn_aggregate_fwd(const float* src, const int32_t* rowptr, const int32_t* col, const uint8_t* eattr,
                                        const float* b1, const float* b2, const float* bias, int64_t N, int D, float* out,
                                        int64_t ld_out, cudaStream_t stream) {
  return aggregate_fwd_launch(src, nullptr, 0, rowptr, col, eattr, nullptr, b1, b2, bias, true, N, D, out, ld_out, 0, nullptr, make_drop(0, 0.f), stream);
}

// Number of partial rows the persistent row-wise kernels with block partials emit (upper bound on
// their grid): callers size `partials` as [molclr_rowwise_max_blocks()][NV][D].
extern "C" int molclr_rowwise_max_blocks(void) { return 8 * sm_count(); }

template <int NCH, bool GATHER, bool DROP>
static int launch_bwd_fused(const float* ga, const int32_t* rowptr_t, const int32_t* col_t, const float* z_prev, const float* bn_coef,
                            int relu, int64_t N, int D, float* gy, float* partials, int round_out, const DropCfg& drop, size_t smem,
                            cudaStream_t stream) {
  auto k = gine_aggregate_bwd_kernel<NCH, 1, GATHER, DROP>;
  const int grid = persistent_grid(k, kRowThreads, smem, kRowWarps, N);
  MOLCLR_LAUNCH(k, grid, kRowThreads, smem, stream, ga, rowptr_t, col_t, z_prev, bn_coef, relu, (int)N, D, gy, partials, round_out, drop);
  return grid;
}

template <int MODE, bool DROP>
static int launch_bwd_tile(const float* ga, const int32_t* rowptr_t, const int32_t* col_t, const uint32_t* nbr_t, const float* z_prev,
                           const float* bn_coef, int relu, int64_t N, int D, float* gy, float* partials, int round_out,
                           const DropCfg& drop, cudaStream_t stream) {
  auto k = gine_aggregate_bwd_tile_kernel<MODE, DROP>;
  const int stages = g_tile_stages < 2 ? 2 : (g_tile_stages > kTileMaxStages ? kTileMaxStages : g_tile_stages);
  const int T = aggregate_tile_rows(D, stages);
  const size_t smem = aggregate_tile_smem(D, T, stages);        // (sized with the forward kernel's table; only more than needed)
  static bool attr_set = false;                      // per instantiation
  if (!attr_set) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = true; }
  const int64_t ntiles = (N + T - 1) / T;
  const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  MOLCLR_LAUNCH(k, grid, kTileThreads, smem, stream, ga, rowptr_t, col_t, nbr_t, z_prev, bn_coef, relu, (int)N, D, T, stages, gy, partials, round_out, drop);
  return grid;
}

static int aggregate_bwd_launch(const float* ga, const int32_t* rowptr_t, const int32_t* col_t, const uint32_t* nbr_t, bool gather,
                                const float* z_prev, const float* bn_coef, int relu, int64_t N, int D, float* gy, int round_out,
                                float* partials, int* num_partials, const DropCfg drop, cudaStream_t stream) {
  REQUIRE_D(D);
  if (num_partials) *num_partials = 0;
  if (N == 0) return 0;
  const int D4 = D / 4;
  if (gather && nbr_t && g_aggregate_tile) {
    MOLCLR_REQUIRE((reinterpret_cast<uintptr_t>(ga) & 15) == 0 && (reinterpret_cast<uintptr_t>(nbr_t) & 15) == 0,
                   "aggregate_bwd: ga and nbr_t must be 16-byte aligned");
    int grid;
    if (z_prev && drop.thr) grid = launch_bwd_tile<1, true>(ga, rowptr_t, col_t, nbr_t, z_prev, bn_coef, relu, N, D, gy, partials, round_out, drop, stream);
    else if (z_prev) grid = launch_bwd_tile<1, false>(ga, rowptr_t, col_t, nbr_t, z_prev, bn_coef, relu, N, D, gy, partials, round_out, drop, stream);
    else grid = launch_bwd_tile<0, false>(ga, rowptr_t, col_t, nbr_t, nullptr, nullptr, 0, N, D, gy, nullptr, round_out, drop, stream);
    if (num_partials && z_prev) *num_partials = grid;
    MOLCLR_CHECK_LAUNCH("aggregate_bwd_tile");
    return 0;
  }
  NCH_DISPATCH(D4, {
    if (z_prev) {
      const size_t smem = (size_t)(4 + 2 * kRowWarps) * D * sizeof(float);
      int grid;
if (gather && drop.thr) grid = launch_bwd_fused<NCH, true, true>(ga, rowptr_t, col_t, z_prev, bn_coef, relu, N, D, gy, partials, round_out, drop, smem, stream);
      else if (gather) grid = launch_bwd_fused<NCH, true, false>(ga, rowptr_t, col_t, z_prev, bn_coef, relu, N, D, gy, partials, round_out, drop, smem, stream);
      else if (drop.thr) grid = launch_bwd_fused<NCH, false, true>(ga, nullptr, nullptr, z_prev, bn_coef, relu, N, D, gy, partials, round_out, drop, smem, stream);
      else grid = launch_bwd_fused<NCH, false, false>(ga, nullptr, nullptr, z_prev, bn_coef, relu, N, D, gy, partials, round_out, drop, smem, stream);
      if (num_partials) *num_partials = grid;
    } else {
      MOLCLR_REQUIRE(gather, "relu_bn_bwd_stats: z_prev is required");
      auto k = gine_aggregate_bwd_kernel<NCH, 0, true, false>;
      MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, 0, kRowWarps, N), kRowThreads, 0, stream, ga, rowptr_t, col_t, nullptr, nullptr, 0,
                    (int)N, D, gy, nullptr, round_out, drop);
    }
  });
  MOLCLR_CHECK_LAUNCH("aggregate_bwd");
  return 0;
}

extern "C" int molclr_gine_aggregate_bwd(const float* ga, const int32_t* rowptr_t, const int32_t* col_t, const uint32_t* nbr_t,
                                         const float* z_prev, const float* bn_coef, int relu, int64_t N, int D, float* gy,
                                         int round_tf32_out, float* partials, int* num_partials, uint32_t drop_seed, float drop_p,
                                         cudaStream_t stream) {
  MOLCLR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability %f not in [0, 1)", drop_p);
  return aggregate_bwd_launch(ga, rowptr_t, col_t, nbr_t, true, z_prev, bn_coef, relu, N, D, gy, round_tf32_out, partials, num_partials,
                              make_drop(drop_seed, z_prev ? drop_p : 0.f), stream);
}

extern "C" int molclr_relu_bn_bwd_stats(const float* g, const float* z_prev, const float* bn_coef, int relu, int64_t N, int D,
                                        float* gy, float* partials, int* num_partials, uint32_t drop_seed, float drop_p,
                                        cudaStream_t stream) {
  MOLCLR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability %f not in [0, 1)", drop_p);
  return aggregate_bwd_launch(g, nullptr, nullptr, nullptr, false, z_prev, bn_coef, relu, N, D, gy, 0, partials, num_partials,
                              make_drop(drop_seed, drop_p), stream);
}

// Bond-embedding table gradients: dB1[t] = sum_i cnt[i][t] * g_a[i], dB2[d] = sum_i cnt[i][5+d] * g_a[i], where cnt[i][.]
// counts node i's in-edges per bond type / direction (self loop included) -- the E' x D embedding_dense_backward of the
// reference collapsed to dB [8][D] = cnt^T [8][N] . g_a [N][D], which is a skinny split-K contraction: it runs
// on the tensor-core GEMM (both operands consumed MN-major in place; counts <= 2048 are exact in TF32).
extern "C" size_t molclr_bn_finalize_workspace_bytes(int D) { return (size_t)kBnSplits * 3 * D * sizeof(double); }

extern "C" int molclr_bn_fwd_finalize(const float* tile_stats, int T, int tile_rows, int64_t N, int D, const float* gamma,
                                      const float* beta, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                      float momentum, float eps, float* coef, void* workspace, cudaStream_t stream) {
  MOLCLR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "bn_fwd_finalize: workspace must be 8-byte aligned");
  MOLCLR_REQUIRE(T > 0 && D > 0, "bn_fwd_finalize: empty statistics");
  int S = (T + 15) / 16;
  if (S > kBnSplits) S = kBnSplits;
  const int per = (T + S - 1) / S;
  S = (T + per - 1) / per;
  double* ws = reinterpret_cast<double*>(workspace);
  MOLCLR_LAUNCH(bn_merge_tiles_kernel, dim3((D + 31) / 32, S), dim3(32, 16), 0, stream, tile_stats, T, tile_rows, (int)N, D, per, ws);
  MOLCLR_CHECK_LAUNCH("bn_merge_tiles");
  MOLCLR_LAUNCH(bn_fwd_finalize_kernel, (D + 7) / 8, 256, 0, stream, ws, S, D, gamma, beta, running_mean, running_var,
                reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, coef);
  MOLCLR_CHECK_LAUNCH("bn_fwd_finalize");
  return 0;
}

extern "C" int molclr_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                   float eps, int D, float* coef, cudaStream_t stream) {
  MOLCLR_LAUNCH(bn_eval_coef_kernel, (D + 127) / 128, 128, 0, stream, gamma, beta, running_mean, running_var, eps, D, coef);
  MOLCLR_CHECK_LAUNCH("bn_eval_coef");
  return 0;
}

extern "C" int molclr_bn_bwd_finalize(const float* partials, int P, int64_t N, int D, const float* gamma, const float* coef,
                                      int use_batch_stats, float* dgamma, float* dbeta, float* bcoef, cudaStream_t stream) {
  MOLCLR_LAUNCH(bn_bwd_finalize_kernel, (D + 31) / 32, dim3(32, 16), 0, stream, partials, P, (int)N, D, gamma, coef, use_batch_stats, dgamma,
                dbeta, bcoef);
  MOLCLR_CHECK_LAUNCH("bn_bwd_finalize");
  return 0;
}

#define REQUIRE_POOL(mode, arg) \
  MOLCLR_REQUIRE((mode) >= 0 && (mode) <= 2 && ((mode) != 2 || (arg) != nullptr), "pool mode %d (0 = mean, 1 = add, 2 = max; max needs the arg-max buffer)", (int)(mode))
#define REQUIRE_DROP(p) MOLCLR_REQUIRE((p) >= 0.f && (p) < 1.f, "dropout probability %f not in [0, 1)", (double)(p))

extern "C" int molclr_bn_bwd_apply(const float* gy, const float* gp, const int32_t* node2graph, const int32_t* gptr, int pool_mode,
                                   const int32_t* argmax, const float* z, const float* bcoef, int64_t N, int D, float* gz, int64_t ld_gz,
                                   int round_tf32_out, float* dbias, float* partials, uint32_t drop_seed, float drop_p,
                                   cudaStream_t stream) {
  REQUIRE_D(D);
  REQUIRE_DROP(drop_p);
  if (gp) REQUIRE_POOL(pool_mode, argmax);
  const DropCfg drop = make_drop(drop_seed, gp ? drop_p : 0.f);
  MOLCLR_REQUIRE(ld_gz >= D && ld_gz % 4 == 0, "bn_bwd_apply: ld_gz must be >= D and a multiple of 4");
  if (N == 0) return 0;
  const size_t smem = (size_t)(3 + kRowWarps) * D * sizeof(float);
  int grid = 1;
  NCH_DISPATCH(D / 4, {
    if (gp) {
      auto k = bn_bwd_apply_kernel<NCH, 1>;
      grid = persistent_grid(k, kRowThreads, smem, kRowWarps, N);
      MOLCLR_LAUNCH(k, grid, kRowThreads, smem, stream, nullptr, gp, node2graph, gptr, pool_mode, argmax, z, bcoef, (int)N, D, gz, ld_gz, partials,
                    round_tf32_out, drop);
    } else {
      auto k = bn_bwd_apply_kernel<NCH, 0>;
      grid = persistent_grid(k, kRowThreads, smem, kRowWarps, N);
      MOLCLR_LAUNCH(k, grid, kRowThreads, smem, stream, gy, nullptr, nullptr, nullptr, 0, nullptr, z, bcoef, (int)N, D, gz, ld_gz, partials,
                    round_tf32_out, drop);
    }
  });
  MOLCLR_CHECK_LAUNCH("bn_bwd_apply");
  if (dbias) return molclr_reduce_partials(partials, grid, D, 1.f, 0, dbias, stream);
  return 0;
}

extern "C" int molclr_pool_fwd(const float* z, const float* bn_coef, int relu, const int32_t* gptr, const int32_t* gperm,
                               int pool_mode, int64_t G, int D, float* out, int64_t ld_out, int round_tf32_out, float* out_lo,
                               int32_t* argmax, uint32_t drop_seed, float drop_p, cudaStream_t stream) {
  REQUIRE_D(D);
  REQUIRE_DROP(drop_p);
  REQUIRE_POOL(pool_mode, argmax);
  if (G == 0) return 0;
  const size_t smem = (size_t)2 * D * sizeof(float);
  NCH_DISPATCH(D / 4, {
    auto k = pool_fwd_kernel<NCH>;
    MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, smem, kRowWarps, G), kRowThreads, smem, stream, z, bn_coef, relu, gptr, gperm,
                  pool_mode, (int)G, D, out, ld_out, round_tf32_out, out_lo, argmax,
                  make_drop(drop_seed, drop_p));
  });
  MOLCLR_CHECK_LAUNCH("pool_fwd");
  return 0;
}

extern "C" int molclr_pool_bwd_stats(const float* gp, const int32_t* node2graph, const int32_t* gptr, int pool_mode,
                                     const int32_t* argmax, const float* z, const float* bn_coef, int64_t N, int D, float* partials,
                                     int* num_partials, uint32_t drop_seed, float drop_p, cudaStream_t stream) {
  REQUIRE_D(D);
  REQUIRE_DROP(drop_p);
  REQUIRE_POOL(pool_mode, argmax);
  if (num_partials) *num_partials = 0;
  if (N == 0) return 0;
  const size_t smem = (size_t)(2 + 2 * kRowWarps) * D * sizeof(float);
  NCH_DISPATCH(D / 4, {
    auto k = pool_bwd_stats_kernel<NCH>;
    const int grid = persistent_grid(k, kRowThreads, smem, kRowWarps, N);
    MOLCLR_LAUNCH(k, grid, kRowThreads, smem, stream, gp, node2graph, gptr, pool_mode, argmax, z, bn_coef, (int)N, D, partials, make_drop(drop_seed, drop_p));
    if (num_partials) *num_partials = grid;
  });
  MOLCLR_CHECK_LAUNCH("pool_bwd_stats");
  return 0;
}

extern "C" int molclr_bn_apply_fwd(const float* z, const float* bn_coef, int relu, int64_t N, int D, float* hi, float* lo,
                                   int64_t ld, int round_hi, uint32_t drop_seed, float drop_p, cudaStream_t stream) {
  REQUIRE_D(D);
  REQUIRE_DROP(drop_p);
  MOLCLR_REQUIRE(ld >= D && ld % 4 == 0, "bn_apply_fwd: ld must be >= D and a multiple of 4");
  if (N == 0) return 0;
  const size_t smem = (size_t)2 * D * sizeof(float);
  NCH_DISPATCH(D / 4, {
    auto k = bn_apply_fwd_kernel<NCH>;
    MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, smem, kRowWarps, N), kRowThreads, smem, stream, z, bn_coef, relu, (int)N, D, hi, lo, ld, round_hi,
                  make_drop(drop_seed, bn_coef ? drop_p : 0.f));
  });
  MOLCLR_CHECK_LAUNCH("bn_apply_fwd");
  return 0;
}

extern "C" int molclr_bn_tile_stats(const float* z, int64_t N, int D, int T, float* tile_stats, cudaStream_t stream) {
  REQUIRE_D(D);
  MOLCLR_REQUIRE((int64_t)T * 32 >= N, "bn_tile_stats: T=%d groups of 32 rows do not cover N=%lld", T, (long long)N);
  if (T == 0) return 0;
  NCH_DISPATCH(D / 4, {
    auto k = bn_tile_stats_kernel<NCH>;
    MOLCLR_LAUNCH(k, persistent_grid(k, kRowThreads, 0, kRowWarps, T), kRowThreads, 0, stream, z, (int)N, D, T, tile_stats);
  });
  MOLCLR_CHECK_LAUNCH("bn_tile_stats");
  return 0;
}

extern "C" int molclr_dropout_mask(uint32_t drop_seed, float drop_p, int64_t N, int D, float* out, cudaStream_t stream) {
  REQUIRE_D(D);
  REQUIRE_DROP(drop_p);
  if (N == 0) return 0;
  int64_t blocks = (N * (D / 4) + 255) / 256;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  MOLCLR_LAUNCH(dropout_mask_kernel, (int)blocks, 256, 0, stream, (int)N, D, out, make_drop(drop_seed, drop_p));
  MOLCLR_CHECK_LAUNCH("dropout_mask");
  return 0;
}

extern "C" int molclr_row_sum(const float* in, int R, int C, float* out, cudaStream_t stream) {
  if (R == 0) return 0;
  MOLCLR_LAUNCH(row_sum_kernel, (R + 7) / 8, 256, 0, stream, in, R, C, out);
  MOLCLR_CHECK_LAUNCH("row_sum");
  return 0;
}

extern "C" int molclr_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height,
                              cudaStream_t stream) {
  if (width_bytes == 0 || height == 0) return 0;
  cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "copy_2d");
  return 0;
}

extern "C" int molclr_act_fwd(const float* x, int mode, int64_t n, float* hi, float* lo, cudaStream_t stream) {
  MOLCLR_REQUIRE(mode == 0 || mode == 1, "act_fwd: mode %d (0 = softplus, 1 = relu)", mode);
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  MOLCLR_LAUNCH(act_fwd_kernel, (int)blocks, 256, 0, stream, x, mode, n, hi, lo);
  MOLCLR_CHECK_LAUNCH("act_fwd");
  return 0;
}

extern "C" int molclr_act_bwd(const float* gy, const float* x, int mode, int64_t n, float* gx, cudaStream_t stream) {
  MOLCLR_REQUIRE(mode == 0 || mode == 1, "act_bwd: mode %d (0 = softplus, 1 = relu)", mode);
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  MOLCLR_LAUNCH(act_bwd_kernel, (int)blocks, 256, 0, stream, gy, x, mode, n, gx);
  MOLCLR_CHECK_LAUNCH("act_bwd");
  return 0;
}

extern "C" int molclr_round_tf32(const float* src, float* dst, float* lo, int64_t n, cudaStream_t stream) {
  if (n == 0) return 0;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  MOLCLR_LAUNCH(round_tf32_kernel, (int)blocks, 256, 0, stream, src, dst, lo, n);
  MOLCLR_CHECK_LAUNCH("round_tf32");
  return 0;
}

extern "C" int molclr_round_tf32_2d(const float* src, int64_t ld_src, float* hi, float* lo, int64_t ld_dst, int64_t rows, int64_t cols,
                                    cudaStream_t stream) {
  if (rows * cols == 0) return 0;
  int64_t blocks = (rows * cols + 1023) / 1024;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  MOLCLR_LAUNCH(round_tf32_2d_kernel, (int)blocks, 256, 0, stream, src, ld_src, hi, lo, ld_dst, (int)rows, (int)cols);
  MOLCLR_CHECK_LAUNCH("round_tf32_2d");
  return 0;
}

extern "C" int molclr_l2_normalize_fwd(const float* z, int64_t R, int C, float eps, float* y, float* inv_norm, cudaStream_t stream) {
  if (R == 0) return 0;
  MOLCLR_LAUNCH(l2_normalize_fwd_kernel, (int)((R + 7) / 8), 256, 0, stream, z, (int)R, C, eps, y, inv_norm);
  MOLCLR_CHECK_LAUNCH("l2_normalize_fwd");
  return 0;
}

extern "C" int molclr_l2_normalize_bwd(const float* gy, const float* y, const float* inv_norm, int64_t R, int C, float eps, float* gz,
                                       cudaStream_t stream) {
  if (R == 0) return 0;
  MOLCLR_LAUNCH(l2_normalize_bwd_kernel, (int)((R + 7) / 8), 256, 0, stream, gy, y, inv_norm, (int)R, C, eps, nullptr, gz);
  MOLCLR_CHECK_LAUNCH("l2_normalize_bwd");
  return 0;
}

extern "C" int molclr_l2_normalize_bwd_scaled(const float* gy, const float* y, const float* inv_norm, int64_t R, int C, float eps,
                                              const float* gscale, float* gz, cudaStream_t stream) {
  if (R == 0) return 0;
  MOLCLR_LAUNCH(l2_normalize_bwd_kernel, (int)((R + 7) / 8), 256, 0, stream, gy, y, inv_norm, (int)R, C, eps, gscale, gz);
  MOLCLR_CHECK_LAUNCH("l2_normalize_bwd_scaled");
  return 0;
}

extern "C" int molclr_ntxent_rows_fwd(const float* zA, const float* zB, int64_t RA, int64_t RB, int C, float eps, int normalise,
                                      float* y, float* y_r, float* inv_norm, void* y16, int64_t ld16, cudaStream_t stream) {
  MOLCLR_REQUIRE((y_r != nullptr || y16 != nullptr) && RA >= 0 && RB >= 0 && RA + RB < (1ll << 31), "ntxent_rows_fwd: bad arguments");
  MOLCLR_REQUIRE(y16 == nullptr || (ld16 >= C && ld16 % 8 == 0), "ntxent_rows_fwd: ld16 must be >= C and a multiple of 8 halves");
  if (RA + RB == 0) return 0;
  MOLCLR_LAUNCH(l2_normalize_cat_fwd_kernel, (int)((RA + RB + 7) / 8), 256, 0, stream, zA, zB, (int)RA, (int)RB, C, eps, normalise, y, y_r, inv_norm,
                reinterpret_cast<__half*>(y16), (int)ld16);
  MOLCLR_CHECK_LAUNCH("ntxent_rows_fwd");
  return 0;
}

extern "C" int molclr_add_inplace(float* y, const float* x, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  MOLCLR_REQUIRE(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, "add_inplace: 16-byte aligned buffers");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  if (blocks < 1) blocks = 1;
  MOLCLR_LAUNCH(add_inplace_kernel, (int)blocks, 256, 0, stream, y, x, n / 4, n);
  MOLCLR_CHECK_LAUNCH("add_inplace");
  return 0;
}
