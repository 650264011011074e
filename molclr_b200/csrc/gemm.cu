// TF32 tensor-core GEMM for the dense contractions of the hot path (GIN MLP 300->600->300 forward
// and backward, GCN 300x300, projection head), written directly against tcgen05 / TMEM / TMA:
//
//   * operands are the fp32 tensors themselves (pre-rounded to TF32 by their producers), staged by
//     TMA into 128B-swizzled shared-memory tiles; out-of-range rows/columns are zero-filled by TMA so
//     the awkward extents (300, 600, ragged node counts) need no padding in HBM;
//   * one elected thread issues tcgen05.mma.kind::tf32 (K=8 per instruction) into a TMEM accumulator: M=128 on one CTA, or
//     M=256 on a CTA pair (cta_group::2: each CTA stages its own 128 rows of A and half of the B tile);
//   * both operands may be K-major ([rows][K]) or MN-major ([K][rows]) so that Y = X W^T, dX = dY W
//     and dW = dY^T X all read the row-major activations/weights in place (no transposed copies);
//   * the epilogue warps (two per TMEM lane quadrant) drain TMEM through per-warp staging chunks and fuse bias, ReLU,
//     ReLU bit masks, addend, TF32 rounding / hi+lo residual outputs, BatchNorm tile statistics / column sums, the NT-Xent
//     log-sum-exp and softmax-weight transforms, and coalesced stores.
//
// A scalar-FMA kernel with the same argument struct (MOLCLR_GEMM_IMPL=simt) exists for debugging
// the tensor-core path on the GPU; it is never selected implicitly.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "molclr_b200.h"
#include "ptx.cuh"
#include "gemm.cuh"

namespace molclr {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 32;          // single-pass K block: 32 tf32 = one 128-byte swizzle row
constexpr int GEMM_BK4 = 32;         // compensated (4-tile) K block.  16 (64-byte swizzle rows, 5 stages, 8 epilogue warps) also works
                                     // but measured 20% slower: the 64-byte TMA boxes double the L2 request count of an L2-bound loop
constexpr int GEMM_TMEM_COLS = 512;  // two accumulator buffers of up to 256 columns
constexpr int GEMM_STAT_ROWS = 32;   // column statistics are emitted per 32-row group (one epilogue warp)
constexpr int GEMM_SMEM_LIMIT = 232448;

static int gemm_workers(bool pair) { return pair ? sm_count() / 2 : sm_count(); }

// FOUR = compensated product with all four operand tiles (A_hi, A_lo, B_hi, B_lo) in one stage and three
// MMAs per K-slice; otherwise one (A, B) pair per stage.
// TWO = CTA pair (cluster of 2, tcgen05 cta_group::2): the pair computes a 256 x BN tile; each CTA stages its own 128 rows
// of A and HALF of the B tile, which cuts the L2 -> shared-memory traffic that bounds these loops (the weights are
// re-read per row tile) by 28..45%.
// H3 (with FOUR) = the fp16 three-product form of the compensated product: stage = [A fp32][A_h16 | A_l16][B_h16 | B_l16], no fp32 B tile
// (see K_PLAIN_H3 below).
template <int BN, bool FOUR, bool TWO, bool H3 = false>
struct GemmCfg {
  static constexpr int BK = FOUR ? GEMM_BK4 : GEMM_BK;
  // BN = 320 ("wide", split-K weight gradients on CTA pairs only): the tile is covered by two MMAs per K slice, N1 = 256
  // and N2 = 64 columns, into adjacent TMEM columns (one accumulator buffer: such a launch gives every worker one tile).
  static constexpr int N1 = BN > 256 ? 256 : BN, N2 = BN - N1;
  static constexpr int BN_CTA = TWO ? BN / 2 : BN;        // B rows staged by one CTA
  static constexpr int A_BYTES = GEMM_BM * BK * 4;
  static constexpr int B_BYTES = BN_CTA * BK * 4;
  static constexpr int MN_BLOCK_BYTES = 32 * BK * 4;      // one [BK k][32 mn] block of an MN-major operand
  static constexpr int STAGE_BYTES = H3 ? 2 * A_BYTES + B_BYTES : (FOUR ? 2 : 1) * (A_BYTES + B_BYTES);
  static_assert(!H3 || FOUR, "H3 is a form of the compensated (FOUR) product");
  // epilogue column chunk staged per warp, and its pitch in floats (conflict-free float4 rows); the 72 KB stages of
  // the compensated product leave room for four warps with 16-column chunks only
  static constexpr bool SMALL_EPI = FOUR && (GEMM_SMEM_LIMIT - 8 * 32 * 32 * 4 - 512 - 2048) / STAGE_BYTES < 3;
  static constexpr int CHUNK = SMALL_EPI ? 16 : 32;
  // 32-column chunks are staged unpadded with an XOR swizzle of the 16-byte column groups (stg_off); 16-column ones padded
  static constexpr int CHUNK_LD = CHUNK == 32 ? 32 : CHUNK + 4;
  // epilogue warps: two per TMEM lane quadrant (splitting the column chunks) where shared memory allows
  // (sixteen epilogue warps -- four per scheduler -- were tried in round 2: no gain, the single-pass row GEMMs are bound by shared-memory
  // bandwidth, not by the epilogue's instruction latency)
  static constexpr int EPI_WARPS = SMALL_EPI ? 4 : 8;
  // converter warps (compensated product with derive_lo): compute the A_lo tile from the fp32 A tile in shared memory
  static constexpr int CONV_WARPS = FOUR ? 8 : 0;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32 * CONV_WARPS;
  static constexpr int STAGING_BYTES = EPI_WARPS * 32 * CHUNK_LD * 4;     // per epilogue warp: 32 rows x chunk
  static constexpr bool BIAS_SMEM = !(FOUR && !TWO);       // (the single-CTA compensated config has no room left)
  static constexpr int BAR_BYTES = 512 + (BIAS_SMEM ? 2 * 256 * 4 : 0);   // mbarriers + TMEM slot, then two bias tiles (double-buffered)
  static constexpr int STAGES_RAW = (GEMM_SMEM_LIMIT - STAGING_BYTES - BAR_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = PIPE_BYTES + STAGING_BYTES + BAR_BYTES;
  static constexpr int TX_BYTES = (TWO ? 2 : 1) * STAGE_BYTES;   // bytes arriving on the (leader's) full barrier per stage
  static_assert(BN % 32 == 0 && (BN <= 256 || (BN == 320 && TWO && !FOUR)), "BN: multiple of 32, <= 256 (or the wide 320 on pairs)");
  static_assert(!TWO || BN_CTA % 8 == 0, "a CTA pair splits B in halves of whole 8-row groups");
  static_assert(STAGE_BYTES % 1024 == 0, "stage bases must stay 1024B aligned for the 128B swizzles");
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

// Global candidate index of the POSITIVE of local row `grow`: rows [0, row_split) are this rank's zjs block (candidates row_offset + r),
// the rest its zis block (candidates row_offset2 + r - row_split); partners sit at the same position of the other block
// (nt_xent.py:53-55: the +-N diagonals).  Independent of how the blocks of different ranks are laid out among the candidates.
__device__ __forceinline__ long long ntx_pos(const GemmParams& p, long long grow) {
  return grow < p.row_split ? grow + p.row_offset2 : grow - p.row_split + p.row_offset;
}

__device__ __forceinline__ float ntx_w_elem(float acc, const GemmParams& p, float lse_r, long long grow_g, long long pos, long long gcol) {
  // W[r][k] = P[r][k] + P[k][r] - 2*[k == pos(r)],  P[i][k] = exp(l - lse_i) for k != i   (nt_xent.py:53-65 differentiated)
  if (gcol == grow_g) return 0.f;
  const float l = acc * p.inv_tau;
  float w = __expf(l - lse_r) + __expf(l - __ldg(p.col_lse + gcol));
  if (gcol == pos) w -= 2.f;
  return w;
}

__device__ __forceinline__ float4 epilogue_apply(float4 v, const GemmParams& p, int grow, int col) {
  if (p.epi == EPI_NTX_W) {
    const float lse_r = __ldg(p.row_lse + grow);
    const long long gr = grow < p.row_split ? grow + p.row_offset : grow - p.row_split + p.row_offset2, gc = col + p.col_offset;
    const long long pos = ntx_pos(p, grow);
    v.x = ntx_w_elem(v.x, p, lse_r, gr, pos, gc); v.y = ntx_w_elem(v.y, p, lse_r, gr, pos, gc + 1);
    v.z = ntx_w_elem(v.z, p, lse_r, gr, pos, gc + 2); v.w = ntx_w_elem(v.w, p, lse_r, gr, pos, gc + 3);
    return v;
  }
  if (p.alpha != 1.f) { v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha; }
  if (p.bias) { const float4 b = ldg_f4(p.bias + col); v = f4_add(v, b); }
  if (p.addend) { const float4 a = *reinterpret_cast<const float4*>(p.addend + (size_t)grow * p.ldadd + col); v = f4_add(v, a); }
  if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  if (p.mask) {
    const float4 m = ldg_f4(p.mask + (size_t)grow * p.ldmask + col);
    if (!(m.x > 0.f)) v.x = 0.f;
    if (!(m.y > 0.f)) v.y = 0.f;
    if (!(m.z > 0.f)) v.z = 0.f;
    if (!(m.w > 0.f)) v.w = 0.f;
  }
  return v;
}

// Column statistics over rows [0, rows) of a staged sub-tile (`ncols` columns, pitch `lds`): mode 1 = sums,
// mode 2 = (mean, M2) for the BatchNorm merge.  Written as partial row `group` of the colstat buffer.
__device__ __forceinline__ void colstat_group(const float* stage, int lds, int ncols, int rows, int col0, int group,
                                              const GemmParams& p, int tid, int nthreads) {
  for (int c = tid; c < ncols; c += nthreads) {
    const int col = col0 + c;
    if (col >= p.N) continue;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += stage[r * lds + c];
    if (p.colstat_mode == 1) {
      p.colstat[(size_t)group * p.N + col] = s;
    } else {
      const float mean = rows > 0 ? s / (float)rows : 0.f;
      float m2 = 0.f;
      for (int r = 0; r < rows; ++r) { const float d = stage[r * lds + c] - mean; m2 = fmaf(d, d, m2); }
      p.colstat[((size_t)group * 2) * p.N + col] = mean;
      p.colstat[((size_t)group * 2 + 1) * p.N + col] = m2;
    }
  }
}

// Float offset of element (row r, column c) of a staged epilogue chunk with pitch LD: LD == 32 -> rows unpadded, the eight
// 16-byte column groups of a row XOR-swizzled with r & 7 (conflict-free for row-wise float4 writes, row-segment float4 reads
// and column reads alike); otherwise plain padded rows.
template <int LD>
__device__ __forceinline__ int stg_off(int r, int c) { return LD == 32 ? r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)) : r * LD + c; }

// Warp version for the tensor-core epilogue: lane c owns column c of a 32-row staged chunk; the 32 row reads are
// independent (fully unrolled), so their latency overlaps.
template <int NCOLS, int LD>
__device__ __forceinline__ void colstat_warp(const float* stage, int rows, int col0, int group, const GemmParams& p, int lane) {
  const int c = lane, col = col0 + c;
  if ((NCOLS < 32 && c >= NCOLS) || group >= p.stat_groups) return;
  float x[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) x[r] = stage[stg_off<LD>(r, c)];
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 32; ++r) s += (r < rows) ? x[r] : 0.f;
  if (col >= p.N) return;
  if (p.colstat_mode == 1) {
    p.colstat[(size_t)group * p.N + col] = s;
  } else {
    const float mean = rows > 0 ? s / (float)rows : 0.f;
    float m2 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) { const float d = x[r] - mean; m2 = (r < rows) ? fmaf(d, d, m2) : m2; }
    p.colstat[((size_t)group * 2) * p.N + col] = mean;
    p.colstat[((size_t)group * 2 + 1) * p.N + col] = m2;
  }
}

// Register version for the TMA-store epilogue (thread = row, v = its 32 columns of the chunk; the staging area then holds the values AS
// STORED, possibly rounded, so the statistics are taken from the exact registers instead): a butterfly transpose-reduction over the
// warp -- 31 shuffles leave the sum of column `lane` in lane `lane` -- which also takes the 32 column reads per thread off the
// shared-memory pipe that bounds these kernels.  Fixed order (pairwise), deterministic.
__device__ __forceinline__ float warp_colsum32(float (&s)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float keep = up ? s[j + off] : s[j];
      const float give = up ? s[j] : s[j + off];
      s[j] = keep + __shfl_xor_sync(0xffffffffu, give, off);
    }
  }
  return s[0];
}

__device__ __forceinline__ void colstat_regs(const float (&v)[32], int rows, int col0, int group, const GemmParams& p, int lane) {
  if (group >= p.stat_groups) return;          // (warp-uniform)
  const bool valid = lane < rows;
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = valid ? v[j] : 0.f;
  const float colsum = warp_colsum32(s, lane);
  const int col = col0 + lane;
  if (p.colstat_mode == 1) {
    if (col < p.N) p.colstat[(size_t)group * p.N + col] = colsum;
    return;
  }
  const float mean = rows > 0 ? colsum / (float)rows : 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = v[j] - __shfl_sync(0xffffffffu, mean, j);
    s[j] = valid ? d * d : 0.f;
  }
  const float m2 = warp_colsum32(s, lane);
  if (col < p.N) {
    p.colstat[((size_t)group * 2) * p.N + col] = mean;
    p.colstat[((size_t)group * 2 + 1) * p.N + col] = m2;
  }
}

// Persistent, warp-specialised tcgen05 GEMM.  Each CTA (one per SM) walks tiles t = blockIdx.x + i*gridDim.x of the
// (n_tile fastest, m_tile, k_split) grid.  The TMA producer and the MMA issuer run ahead across tiles through a
// STAGES-deep smem ring; accumulators are double-buffered in TMEM so the epilogue warps drain tile i while the
// tensor core works on tile i+1.
// KIND selects the epilogue at compile time so that each variant is a short, branch-free loop:
// K_PLAIN_H3 (compensate = 2): the plain epilogue on the fp16 three-product compensated main loop -- A is split on chip into
// fp16(x) and fp16(x - fp16(x)) (22 significand bits), B arrives pre-split the same way (molclr_prepare_weights, scaled by 2^6 so
// that the low halves of typical weights stay normal fp16 numbers; the epilogue's alpha undoes it), and the product is
// A_h B_h + A_l B_h + A_h B_l as three kind::f16 MMAs per K = 16: 3/4 of the tensor time and 2/3 of the L2 -> shared-memory
// bytes of the TF32 + bf16 form (no fp32 B tile), one more ring stage, and a smaller rounding error (hi halves rounded to
// nearest instead of truncated, 11-bit instead of 8-bit corrections).  Price: fp16's range -- see molclr_gemm_args.compensate.
enum : int { K_PLAIN = 0, K_LATE = 1, K_NTX_W = 2, K_NTX_FWD = 3, K_ATOMIC = 4, K_PLAIN16 = 5, K_NTX_W16 = 6, K_NTX_FWD16 = 7, K_PLAIN_H3 = 8 };

template <int BN, bool FOUR, int KIND, bool TWO>
__global__ void __launch_bounds__(GemmCfg<BN, FOUR, TWO, KIND == K_PLAIN_H3>::THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmO,
                 const GemmParams p) {
  constexpr bool H3 = KIND == K_PLAIN_H3;
  using Cfg = GemmCfg<BN, FOUR, TWO, H3>;
  // fp16-operand instances (K-major tiles of 64 halves per 128-byte row, kind::f16 MMAs) share the epilogue of their TF32 kind
  constexpr bool H16 = KIND >= K_PLAIN16 && !H3;
  constexpr int EK = (KIND == K_PLAIN16 || H3) ? K_PLAIN : KIND == K_NTX_W16 ? K_NTX_W : KIND == K_NTX_FWD16 ? K_NTX_FWD : KIND;
  static_assert(!(H16 && FOUR), "fp16 operands: single-pass instances only");
  static_assert(!H3 || FOUR, "the fp16 three-product form is a compensated (FOUR) instance");
  extern __shared__ __align__(1024) uint8_t smem[];
  float* staging = reinterpret_cast<float*>(smem + Cfg::PIPE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;              // [2] accumulator drained
  uint64_t* afull_bar = tempty_bar + 2;              // [STAGES] derive_lo: this CTA's fp32 A tile has landed
  uint64_t* conv_bar = afull_bar + Cfg::STAGES;      // [STAGES] derive_lo: A_lo tiles of the worker converted (leader's copy is used)
  uint64_t* raw_bar = conv_bar + Cfg::STAGES;        // [STAGES] mixed: the raw fp32 tiles of the worker have landed (leader's copy is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_bar + Cfg::STAGES);
  static_assert((5 * Cfg::STAGES + 4) * 8 + 4 <= 512, "barrier area");
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::PIPE_BYTES + Cfg::STAGING_BYTES + 512);   // [2][256]
  const bool derive = FOUR && p.derive_lo != 0;
  // derive_lo == 2 ("mixed"): A and B are both UNROUNDED fp32, K-major; stage = [A][A_hi16|A_lo16][B][B_hi16|B_lo16]: the tensor core
  // truncates the raw tiles for the TF32 pass, the converter warps form bf16(x) and bf16(x - trunc_tf32(x)) of both tiles for the
  // two correction passes, which run as kind::f16 MMAs -- no rounded or split copy of either operand is ever read from HBM / L2
  const bool mixed = H3 || (FOUR && p.derive_lo == 2);     // (H3: as mixed, without an fp32 B tile; the B16 tiles are mandatory)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_tiles, m_tiles = p.m_tiles;
  const int total = n_tiles * m_tiles * p.splits;
  // a CTA pair is one worker: both CTAs walk the same tile sequence, rank r owns rows [128 r, 128 r + 128) of the 256-row
  // tile and stages rows [r BN/2, (r+1) BN/2) of the B tile; only the leader (rank 0) issues MMAs.
  const uint32_t rank = TWO ? ptx::cluster_ctarank() : 0u;
  const int worker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, num_workers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = TWO ? 2 * GEMM_BM : GEMM_BM;

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) { printf("molclr gemm: dynamic smem base not 1024B aligned\n"); __trap(); }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1);
      ptx::mbar_init(afull_bar + s, 1); ptx::mbar_init(conv_bar + s, (TWO ? 2 : 1) * (Cfg::CONV_WARPS > 0 ? Cfg::CONV_WARPS : 1));
      ptx::mbar_init(raw_bar + s, TWO ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(tfull_bar + b, 1); ptx::mbar_init(tempty_bar + b, (TWO ? 2 : 1) * Cfg::EPI_WARPS); }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (p.segments > 1) { ptx::prefetch_tensormap(&tmA2); ptx::prefetch_tensormap(&tmB2); }
    if (p.tma_store) ptx::prefetch_tensormap(&tmO);
  }
  if (warp == 1) {
    if (TWO) { ptx::tmem_alloc_2cta(tmem_slot, GEMM_TMEM_COLS); ptx::tmem_relinquish_2cta(); }
    else { ptx::tmem_alloc(tmem_slot, GEMM_TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (TWO) ptx::cluster_sync(); else __syncthreads();     // barriers initialised and TMEM allocated in BOTH CTAs
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();          // everything above overlaps the tail of the previous kernel of the stream; global memory is touched only from here

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp walks the loop with warp-uniform
    // coordinates and addresses; one elected lane issues the copies)
    {
      uint32_t it = 0;                                  // global k-block counter -> ring slot / phase
      // L2 prefetch cursor (p.pf > 0, an experiment that did not pay -- see gemm_run): runs p.pf k-blocks ahead of the loads through the
      // same tile sequence.
      int pf_t = worker, pf_i = 0, pf_ahead = 0;
      for (int t = worker; t < total; t += num_workers) {
        const int nb0 = (t % n_tiles) * BN;                                  // first column of the tile
        const int n0 = nb0 + (int)rank * Cfg::BN_CTA, m0 = ((t / n_tiles) % m_tiles) * TILE_M + (int)rank * GEMM_BM;
        const int kb0 = (t / (n_tiles * m_tiles)) * p.kb_per_split;
        const int nkb_seg = min(p.num_kb, kb0 + p.kb_per_split) - kb0;
        const int nkb = FOUR ? nkb_seg : nkb_seg * p.segments;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % Cfg::STAGES;
          ptx::mbar_wait(empty_bar + s, ((it / Cfg::STAGES) & 1) ^ 1);
          if (ptx::elect_one()) {
            if (p.pf > 0) {
              for (; pf_ahead < p.pf && pf_t < total; ++pf_ahead) {
                const int pm0 = ((pf_t / n_tiles) % m_tiles) * TILE_M + (int)rank * GEMM_BM;
                if (pm0 < p.M) ptx::tma_prefetch_2d(&tmA, pf_i * (H16 ? 2 * Cfg::BK : Cfg::BK), pm0);
                if (++pf_i == p.num_kb) { pf_i = 0; pf_t += num_workers; }
              }
              --pf_ahead;
            }
            // derive_lo: the fp32 A tile goes to a CTA-local barrier (the converter warps of THIS CTA consume it) and no A_lo is loaded
            // mixed: both raw tiles go to the CTA-local barrier and the MMA issuer waits for the converters only
            if (rank == 0 && !mixed) ptx::mbar_arrive_expect_tx(full_bar + s, derive ? (TWO ? 2 : 1) * 2 * Cfg::B_BYTES : Cfg::TX_BYTES);
            if (derive)
              ptx::mbar_arrive_expect_tx(afull_bar + s, H3 ? Cfg::A_BYTES + Cfg::B_BYTES
                                                           : mixed ? Cfg::A_BYTES + Cfg::B_BYTES + (p.b_presplit ? Cfg::B_BYTES : 0) : Cfg::A_BYTES);
            const uint32_t fb = TWO ? ptx::mapa(ptx::smem_u32(full_bar + s), 0u) : 0u;     // the leader's barrier
            auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
              if (TWO) ptx::tma_load_2d_2cta(dst, m, fb, c0, c1); else ptx::tma_load_2d(dst, m, full_bar + s, c0, c1);
            };
            auto load3 = [&](void* dst, const CUtensorMap* m, int k0, int blk) {      // MN-major: several [k][32 mn] blocks at once
              if (TWO) ptx::tma_load_3d_2cta(dst, m, fb, 0, k0, blk); else ptx::tma_load_3d(dst, m, full_bar + s, 0, k0, blk);
            };
            uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
            uint8_t* b_dst = a_dst + (FOUR ? 2 : 1) * Cfg::A_BYTES;
            const int seg = FOUR ? 0 : i / nkb_seg;
            const int kc = (kb0 + (i - seg * nkb_seg)) * (H16 ? 2 * Cfg::BK : Cfg::BK);   // fp16: 64 elements per 128-byte tile row
  #pragma unroll
            for (int h = 0; h < (FOUR ? 2 : 1); ++h) {
              const CUtensorMap* ma = (FOUR ? h == 1 : seg == 1) ? &tmA2 : &tmA;
              const CUtensorMap* mb = (FOUR ? h == 1 : seg == 2) ? &tmB2 : &tmB;
              uint8_t* ad = a_dst + h * Cfg::A_BYTES;
              uint8_t* bd = b_dst + h * Cfg::B_BYTES;
              if (derive) {
                if (h == 0) {
                  if (!p.a_mn) ptx::tma_load_2d(ad, ma, afull_bar + s, kc, m0);
                  else
                    for (int j = 0; j < GEMM_BM / 32; ++j) ptx::tma_load_2d(ad + j * Cfg::MN_BLOCK_BYTES, ma, afull_bar + s, m0 + 32 * j, kc);
                }
              } else if (!p.a_mn) load(ad, ma, kc, m0);
              else if (!FOUR && p.a_mn3d) load3(ad, &tmA2, kc, m0 >> 5);               // 4 blocks in one box
              else
                for (int j = 0; j < GEMM_BM / 32; ++j) load(ad + j * Cfg::MN_BLOCK_BYTES, ma, m0 + 32 * j, kc);
              if (mixed) {
                if (h == 0) {
                  if (!H3) ptx::tma_load_2d(bd, &tmB, afull_bar + s, kc, n0);
                  if (H3 || p.b_presplit) {
                    // the two 16-bit tiles of B (rows of 32 elements = 64 bytes, 64-byte swizzle), pre-split once per step: bf16(B) and
                    // bf16(B - trunc B) behind the fp32 tile, or (H3) fp16(s B) and fp16(s B - hi) in its place
                    uint8_t* b16d = H3 ? bd : bd + Cfg::B_BYTES;
                    ptx::tma_load_2d(b16d, &tmB2, afull_bar + s, kc, n0);
                    ptx::tma_load_2d(b16d + Cfg::B_BYTES / 2, &tmB2, afull_bar + s, kc, p.rows16 + n0);
                  }
                }
              } else if (!p.b_mn) load(bd, mb, kc, n0);
              else if (!FOUR && p.b_mn3d) {
                // this CTA's blocks of the first MMA (N1) in one box; the (shorter) N2 group of a wide tile keeps per-block copies
                constexpr int J1 = (TWO ? Cfg::N1 / 2 : Cfg::N1) / 32;
                load3(bd, &tmB2, kc, (nb0 + (int)rank * (TWO ? Cfg::N1 / 2 : 0)) >> 5);
                for (int j = J1; j < Cfg::BN_CTA / 32; ++j)
                  load(bd + j * Cfg::MN_BLOCK_BYTES, mb, nb0 + Cfg::N1 + (int)rank * (Cfg::N2 / 2) + 32 * (j - J1), kc);
              } else
                for (int j = 0; j < Cfg::BN_CTA / 32; ++j) {
                  // blocks of the first MMA (N1) first, then those of the second (N2); each CTA of a pair stages its half of both
                  constexpr int J1 = (TWO ? Cfg::N1 / 2 : Cfg::N1) / 32;
                  const int colb = j < J1 ? nb0 + (int)rank * (TWO ? Cfg::N1 / 2 : 0) + 32 * j
                                          : nb0 + Cfg::N1 + (int)rank * (Cfg::N2 / 2) + 32 * (j - J1);
                  load(bd + j * Cfg::MN_BLOCK_BYTES, mb, colb, kc);
                }
            }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop, so that descriptors and barrier addresses are warp-uniform values (uniform registers, no
    // per-lane "waterfall" moves into them); one elected lane issues the tcgen05 instructions.  This thread sits between "stage
    // full" and "stage free again": every instruction here is on the ring's critical path.
    if (rank == 0) {
      constexpr bool h16 = H16;
      // fp16 operands (K-major): the same 128-byte-row tiles and descriptors, K = 16 (32 bytes) per kind::f16 instruction
      const uint32_t idesc = h16 ? ptx::make_idesc_f16(Cfg::N1, TILE_M) : ptx::make_idesc_tf32(Cfg::N1, p.a_mn != 0, p.b_mn != 0, TILE_M);
      const uint32_t idesc2 = ptx::make_idesc_tf32(Cfg::N2 > 0 ? Cfg::N2 : 16, p.a_mn != 0, p.b_mn != 0, TILE_M);
      const uint32_t idesc16 = ptx::make_idesc_bf16(Cfg::N1, TILE_M);
      const uint32_t idesc_h3 = ptx::make_idesc_f16(Cfg::N1, TILE_M);
      auto mma_i = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (h16) { if (TWO) ptx::mma_f16_ss_2cta(d, a, b, id, acc); else ptx::mma_f16_ss(d, a, b, id, acc); }
        else if (TWO) ptx::mma_tf32_ss_2cta(d, a, b, id, acc); else ptx::mma_tf32_ss(d, a, b, id, acc);
      };
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) { mma_i(d, a, b, idesc, acc); };
      auto commit = [&](uint64_t* bar) { if (TWO) ptx::mma_commit_2cta(bar); else ptx::mma_commit(bar); };
      // K-major tiles: rows of BK*4 bytes (128B or 64B swizzle), 8-row groups SBO apart.  MN-major: [BK k][32 mn] blocks
      // LBO apart, 4-k-row groups 512 B apart.
      constexpr uint32_t kmaj_sbo = 8u * Cfg::BK * 4u, kmaj_lay = Cfg::BK == 16 ? ptx::kLayoutSw64 : ptx::kLayoutSw128;
      const uint32_t a_lbo = p.a_mn ? (uint32_t)Cfg::MN_BLOCK_BYTES : 16u, b_lbo = p.b_mn ? (uint32_t)Cfg::MN_BLOCK_BYTES : 16u;
      const uint32_t a_sbo = p.a_mn ? 512u : kmaj_sbo, b_sbo = p.b_mn ? 512u : kmaj_sbo;
      const uint32_t a_lay = p.a_mn ? ptx::kLayoutSw128Base32 : kmaj_lay, b_lay = p.b_mn ? ptx::kLayoutSw128Base32 : kmaj_lay;
      const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;   // bytes per K=8 slice
      // descriptors without the start address: its 14-bit field (bytes / 16; shared memory is < 256 KB) is added per MMA
      const uint64_t a_d0 = ptx::make_smem_desc(0u, a_lbo, a_sbo, a_lay), b_d0 = ptx::make_smem_desc(0u, b_lbo, b_sbo, b_lay);
      const uint64_t h_d0 = ptx::make_smem_desc(0u, 16u, 512u, ptx::kLayoutSw64);
      const uint32_t smem_base = ptx::smem_u32(smem);
      uint32_t it = 0, tl = 0;                          // tl = local tile counter -> accumulator buffer / phase
      for (int t = worker; t < total; t += num_workers, ++tl) {
        const int kb0 = (t / (n_tiles * m_tiles)) * p.kb_per_split;
        const int nkb_seg = min(p.num_kb, kb0 + p.kb_per_split) - kb0;
        const int nkb = FOUR ? nkb_seg : nkb_seg * p.segments;
        const uint32_t buf = tl & 1;
        ptx::mbar_wait(tempty_bar + buf, ((tl >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256u;
        // (both run on the elected lane only)  mixed: correction passes of a stage on its bf16 tiles: K-major rows of 32 bf16 = 64 bytes
        // (64B swizzle, 8-row groups 512 B apart), K = 16 (32 bytes) per instruction
        auto corrections = [&](uint32_t a_base, uint32_t b_base) {
          if constexpr (FOUR) {
            if (mixed && !(p.debug & 4)) {
#pragma unroll
              for (int k = 0; k < Cfg::BK / 16; ++k) {
                const uint64_t ah = h_d0 + ((a_base + Cfg::A_BYTES + k * 32u) >> 4), al = ah + (Cfg::A_BYTES / 2 >> 4);
                const uint64_t bh = h_d0 + ((b_base + Cfg::B_BYTES + k * 32u) >> 4), bl = bh + (Cfg::B_BYTES / 2 >> 4);
                if (TWO) { ptx::mma_f16_ss_2cta(d_tmem, al, bh, idesc16, 1u); ptx::mma_f16_ss_2cta(d_tmem, ah, bl, idesc16, 1u); }
                else { ptx::mma_f16_ss(d_tmem, al, bh, idesc16, 1u); ptx::mma_f16_ss(d_tmem, ah, bl, idesc16, 1u); }
              }
            }
          }
        };
        auto finish = [&](int s, bool last) {
          commit(empty_bar + s);                    // frees the smem slot (of both CTAs) once these MMAs have read it
          if (last) commit(tfull_bar + buf);        // accumulator complete
        };
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % Cfg::STAGES;
          // mixed: the TF32 pass reads only the raw tiles, so it is issued as soon as they have LANDED in both CTAs (raw_bar) and runs on
          // the tensor pipe while the converter warps are still forming the bf16 tiles of the same stage; the correction passes wait
          // for conv_bar.  Same instruction order as issuing all eight after the conversion: bit-identical results.
          if constexpr (H3) {
            // fp16 three-product form: every MMA reads converted tiles, K = 16 (32 bytes of the 64-byte-swizzle rows) per instruction
            ptx::mbar_wait(conv_bar + s, (it / Cfg::STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t a16 = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES, b16 = a16 + Cfg::A_BYTES;
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < Cfg::BK / 16; ++k) {
                const uint64_t ah = h_d0 + ((a16 + k * 32u) >> 4), al = ah + (Cfg::A_BYTES / 2 >> 4);
                const uint64_t bh = h_d0 + ((b16 + k * 32u) >> 4), bl = bh + (Cfg::B_BYTES / 2 >> 4);
                const uint32_t acc0 = (i | k) != 0 ? 1u : 0u;
                if (TWO) {
                  ptx::mma_f16_ss_2cta(d_tmem, ah, bh, idesc_h3, acc0);
                  if (!(p.debug & 4)) { ptx::mma_f16_ss_2cta(d_tmem, al, bh, idesc_h3, 1u); ptx::mma_f16_ss_2cta(d_tmem, ah, bl, idesc_h3, 1u); }
                } else {
                  ptx::mma_f16_ss(d_tmem, ah, bh, idesc_h3, acc0);
                  if (!(p.debug & 4)) { ptx::mma_f16_ss(d_tmem, al, bh, idesc_h3, 1u); ptx::mma_f16_ss(d_tmem, ah, bl, idesc_h3, 1u); }
                }
              }
              finish(s, i == nkb - 1);
            }
            __syncwarp();
            continue;
          }
          const bool early = mixed && !(p.debug & 16);
          if (!mixed) ptx::mbar_wait(full_bar + s, (it / Cfg::STAGES) & 1);
          if (early) ptx::mbar_wait(raw_bar + s, (it / Cfg::STAGES) & 1);
          else if (derive) ptx::mbar_wait(conv_bar + s, (it / Cfg::STAGES) & 1);
          ptx::tc_fence_after();
          const uint32_t a_base = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_base = a_base + (FOUR ? 2 : 1) * Cfg::A_BYTES;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < Cfg::BK / 8; ++k) {
              const uint64_t ad = a_d0 + ((a_base + k * a_kstep) >> 4);
              const uint64_t bd = b_d0 + ((b_base + k * b_kstep) >> 4);
              mma(d_tmem, ad, bd, (i | k) != 0 ? 1u : 0u);
              if (Cfg::N2 > 0) {     // wide tile: the remaining N2 columns (MN-major B: its blocks follow those of the first MMA)
                constexpr uint32_t b2_off = (uint32_t)((TWO ? Cfg::N1 / 2 : Cfg::N1) / 32) * Cfg::MN_BLOCK_BYTES;
                mma_i(d_tmem + Cfg::N1, ad, bd + (b2_off >> 4), idesc2, (i | k) != 0 ? 1u : 0u);
              }
              if (FOUR && !mixed) {
                const uint64_t ad2 = ad + (Cfg::A_BYTES >> 4), bd2 = bd + (Cfg::B_BYTES >> 4);
                mma(d_tmem, ad2, bd, 1u);          // A_lo * B_hi
                mma(d_tmem, ad, bd2, 1u);          // A_hi * B_lo
              }
            }
            if (!(FOUR && early)) { corrections(a_base, b_base); finish(s, i == nkb - 1); }
          }
          if (FOUR && early) {
            __syncwarp();
            ptx::mbar_wait(conv_bar + s, (it / Cfg::STAGES) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) { corrections(a_base, b_base); finish(s, i == nkb - 1); }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (Cfg::CONV_WARPS > 0 && warp >= 2 + Cfg::EPI_WARPS) {
    // ------------------------------------------------------------ converter (derive_lo): A_lo = tf32(A - trunc_tf32(A))
    // The tensor core reads only the top 19 bits of an fp32 operand, so the raw tile IS the "hi" operand (truncated); the
    // residual is formed element-wise at the same (swizzled) offset of the A_lo slot, whatever the tile layout.
    if (derive) {
      const int ct = threadIdx.x - 32 * (2 + Cfg::EPI_WARPS);
      const uint32_t conv_leader = TWO ? ptx::mapa(ptx::smem_u32(conv_bar), 0u) : 0u;
      float amax = 0.f;                // H3: largest |A element| this thread has converted
      uint32_t it = 0;
      for (int t = worker; t < total; t += num_workers) {
        const int kb0 = (t / (n_tiles * m_tiles)) * p.kb_per_split;
        const int nkb = min(p.num_kb, kb0 + p.kb_per_split) - kb0;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % Cfg::STAGES;
          ptx::mbar_wait(afull_bar + s, (it / Cfg::STAGES) & 1);
          if (mixed && !H3 && ct == 0) {      // this CTA's raw tiles have landed: the leader's MMA warp may start the TF32 pass of the stage
            if (TWO) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(raw_bar + s), 0u)); else ptx::mbar_arrive(raw_bar + s);
          }
          const float4* hi = reinterpret_cast<const float4*>(smem + s * Cfg::STAGE_BYTES);
          float4* lo = reinterpret_cast<float4*>(smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES);
          if (mixed) {
            // bf16 tiles for the correction passes: hi16 = bf16(x), lo16 = bf16(x - trunc_tf32(x)), of the A tile and of this CTA's
            // B tile.  Source: K-major fp32 rows of 128 B, 16-byte chunk c of row r stored at chunk c ^ (r & 7); destination: rows
            // of 64 B, chunk j at j ^ ((r >> 1) & 3).
            // Thread ct owns chunks ct + NT*i (NT = converter threads, a multiple of 64 = 8 rows): the row's swizzle phases are the
            // same for all of them, so source and destination offsets are per-thread constants plus a fixed stride.
            constexpr int NT = Cfg::CONV_WARPS > 0 ? 32 * Cfg::CONV_WARPS : 32;   // (no converter warps: dead code in that instance)
            const int r0 = ct >> 3, c = (ct & 7) ^ (r0 & 7);                 // logical 16-byte chunk of the row: k = 4c .. 4c + 3
            const int off0 = r0 * 64 + (((c >> 1) ^ ((r0 >> 1) & 3)) << 4) + ((c & 1) << 3);
            auto convert = [&](const uint8_t* src, uint8_t* hi16, int chunks) {
              uint8_t* lo16 = hi16 + chunks * 8;                             // the lo tile follows the hi tile (half the bytes each)
              constexpr int U = 4;
              for (int e0 = ct; e0 < chunks; e0 += U * NT) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = (e0 + u * NT < chunks) ? reinterpret_cast<const float4*>(src)[e0 + u * NT] : f4_zero();
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  if (e0 + u * NT >= chunks) break;
                  const int off = off0 + ((e0 - ct) / NT + u) * (NT / 8) * 64;
                  const float4 x = v[u];
                  const float lx = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), ly = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                  const float lz = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), lw = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                  __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
                  __nv_bfloat162 l0 = __floats2bfloat162_rn(lx, ly), l1 = __floats2bfloat162_rn(lz, lw);
                  *reinterpret_cast<uint2*>(hi16 + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                  *reinterpret_cast<uint2*>(lo16 + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
                }
              }
            };
            // H3: h = fp16(x) (round to nearest; saturating, so that an out-of-range value stays finite and is reported through
            // p.status instead of poisoning the tile), l = fp16(x - h): 22 significand bits in two fp16 tiles of the same layout
            auto convert_h3 = [&](const uint8_t* src, uint8_t* hi16, int chunks) {
              uint8_t* lo16 = hi16 + chunks * 8;
              constexpr int U = 4;
              for (int e0 = ct; e0 < chunks; e0 += U * NT) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = (e0 + u * NT < chunks) ? reinterpret_cast<const float4*>(src)[e0 + u * NT] : f4_zero();
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  if (e0 + u * NT >= chunks) break;
                  const int off = off0 + ((e0 - ct) / NT + u) * (NT / 8) * 64;
                  const float4 x = v[u];
                  const uint32_t h0 = ptx::cvt_f16x2_sat(x.x, x.y), h1 = ptx::cvt_f16x2_sat(x.z, x.w);
                  const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
                  const uint32_t l0 = ptx::cvt_f16x2_sat(x.x - f0.x, x.y - f0.y), l1 = ptx::cvt_f16x2_sat(x.z - f1.x, x.w - f1.y);
                  *reinterpret_cast<uint2*>(hi16 + off) = make_uint2(h0, h1);
                  *reinterpret_cast<uint2*>(lo16 + off) = make_uint2(l0, l1);
                  amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
                }
              }
            };
            uint8_t* st = smem + s * Cfg::STAGE_BYTES;
            if (H3) {
              if (!(p.debug & 2)) convert_h3(st, st + Cfg::A_BYTES, Cfg::A_BYTES / 16);
            } else if (!(p.debug & 2)) {      // (timing experiments: skip the conversion)
              convert(st, st + Cfg::A_BYTES, Cfg::A_BYTES / 16);
              if (!p.b_presplit) convert(st + 2 * Cfg::A_BYTES, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, Cfg::B_BYTES / 16);
            }
          } else
#pragma unroll 4
          for (int e = ct; e < Cfg::A_BYTES / 16; e += 32 * Cfg::CONV_WARPS) {
            const float4 v = hi[e];
            float4 r;
            r.x = round_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
            r.y = round_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
            r.z = round_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
            r.w = round_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
            lo[e] = r;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (TWO) ptx::mbar_arrive_cluster(conv_leader + s * 8u); else ptx::mbar_arrive(conv_bar + s);
          }
        }
      }
      // H3: an A element beyond fp16's largest finite value was clamped -- the product is wrong; tell the caller (sticky status word)
      if (H3 && amax > 65504.f && p.status) atomicOr(p.status, MOLCLR_STATUS_FP16_RANGE);
    }
  } else {
    // ------------------------------------------------------------ epilogue: warp q owns accumulator rows 32q..32q+31
    const int q = warp & 3;                          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                // 0/1: which share of the column chunks (8-warp configs)
    constexpr int NSHARE = Cfg::EPI_WARPS / 4;
    float* stg = staging + (warp - 2) * 32 * Cfg::CHUNK_LD;
    uint32_t tl = 0;
    const uint32_t tempty_leader = TWO ? ptx::mapa(ptx::smem_u32(tempty_bar), 0u) : 0u;
    for (int t = worker; t < total; t += num_workers, ++tl) {
      const int n_tile = t % n_tiles, m_tile = (t / n_tiles) % m_tiles;
      const int n0 = n_tile * BN, m0 = m_tile * TILE_M + (int)rank * GEMM_BM;
      const int row = q * 32 + lane, grow = m0 + row;
      const int rows_w = max(0, min(32, p.M - m0 - q * 32));          // valid rows of this warp's group
      const uint32_t buf = tl & 1;
      constexpr int CH = Cfg::CHUNK, LD = Cfg::CHUNK_LD;
      constexpr int STEP = CH * NSHARE;                  // column distance between consecutive chunks of this warp
      constexpr int NCHUNK = (BN + STEP - 1) / STEP;
      // ReLU-mask word of this thread's row for its first chunk (K_PLAIN backward GEMMs): fetched BEFORE waiting for
      // the accumulator, the following ones one chunk ahead, so that their latency hides behind other work.
      auto mask_word = [&](int k) -> uint32_t {
        const int c0 = CH * (half + k * NSHARE);
        return (c0 < BN && n0 + c0 < p.N && grow < p.M) ? __ldg(p.bits_in + (size_t)grow * p.ld_bits + ((n0 + c0) >> 5)) : 0u;
      };
      uint32_t mw_next = 0u;
      if (EK == K_PLAIN && !FOUR && p.bits_in) mw_next = mask_word(0);
      // this tile's bias slice goes to shared memory once (double-buffered across tiles; one named barrier per tile
      // among the epilogue warps), so that the register stage reads it with broadcast LDS instead of dependent LDGs
      const float* bias_t = bias_s + (tl & 1) * 256;
      if (EK == K_PLAIN && Cfg::BIAS_SMEM && p.bias) {
        for (int e = threadIdx.x - 64; e < BN; e += Cfg::EPI_WARPS * 32)
          bias_s[(tl & 1) * 256 + e] = (n0 + e < p.N) ? __ldg(p.bias + n0 + e) : 0.f;
        ptx::named_bar_sync(1, Cfg::EPI_WARPS * 32);
      }
      float nlr = 0.f;                  // K_NTX_W, fp16 output: 10 - log2e * (log-sum-exp of this thread's row)
      if (EK == K_NTX_W) {            // the candidates' log-sum-exps of this column tile, pre-scaled for ex2
        // (fp16 output: negated and shifted by 10, so that ex2(fma(s, k2, .)) is the softmax weight times 2^10)
        constexpr bool w16 = H16;
        for (int e = threadIdx.x - 64; e < BN; e += Cfg::EPI_WARPS * 32) {
          const float l2 = (n0 + e < p.N) ? __ldg(p.col_lse + p.col_offset + n0 + e) * 1.4426950408889634f : 0.f;
          // (bounded logits: the column factor 2^(10 + bound - lse_k) itself)
          bias_s[(tl & 1) * 256 + e] = !w16 ? l2 : p.ntx_bound2 > 0.f ? ptx::ex2_approx(10.f + p.ntx_bound2 - l2) : 10.f - l2;
        }
        if (w16) nlr = 10.f - (grow < p.M ? __ldg(p.row_lse + grow) : 0.f) * 1.4426950408889634f;
        if (w16 && p.ntx_bound2 > 0.f) nlr = ptx::ex2_approx(nlr + p.ntx_bound2);
        ptx::named_bar_sync(1, Cfg::EPI_WARPS * 32);
      }
      ptx::mbar_wait(tfull_bar + buf, (tl >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
      if (EK == K_NTX_FWD) {
        // per-row (max, sum exp) of this warp's share of the column tile, own column masked; thread <-> row straight
        // from TMEM, one pass with a running maximum.  Partials are indexed [n_tile * NSHARE + half][row].
        // 32 columns per step, the next step's TMEM loads in flight while this one is reduced; per element the loop costs
        // FMNMX + FFMA + EX2 + FADD: the self column and the positive are looked for only in the blocks that hold them.
        const long long gr = grow < p.row_split ? grow + p.row_offset : grow - p.row_split + p.row_offset2;
        const long long pos = ntx_pos(p, grow);
        const float k2 = p.inv_tau * 1.4426950408889634f;            // logits in base-2 units
        float mx = (H16 && p.ntx_bound2 > 0.f) ? p.ntx_bound2 : -INFINITY, sum = 0.f;
        constexpr int NBLK = (BN / 32 + NSHARE - 1) / NSHARE;        // 32-column blocks of this warp: c0 = 32 (half + b NSHARE)
        auto blk_c0 = [&](int b) { return 32 * (half + b * NSHARE); };
        auto blk_ok = [&](int b) { return blk_c0(b) < BN && n0 + blk_c0(b) < p.N; };
        auto blk_load = [&](int b, float (&v)[32]) {
          ptx::tmem_ld_x16_nowait(taddr + blk_c0(b), v);
          ptx::tmem_ld_x16_nowait(taddr + blk_c0(b) + 16, v + 16);
        };
        auto blk_reduce = [&](int b, float (&v)[32]) {
          const int c0 = blk_c0(b), nvalid = p.N - n0 - c0;           // > 0
          const long long gc0 = (long long)n0 + c0 + p.col_offset;
          const unsigned long long d_self = (unsigned long long)(gr - gc0), d_pos = (unsigned long long)(pos - gc0);
          if (nvalid < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j < nvalid ? v[j] : -INFINITY;
          }
          if (d_pos < 32ull && grow < p.M) {
            const int dp = (int)d_pos;
            float pv = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) pv = j == dp ? v[j] : pv;
            p.row_pos[grow] = pv * p.inv_tau;
          }
          if (d_self < 32ull) {
            const int ds = (int)d_self;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j == ds ? -INFINITY : v[j];
          }
          if (H16 && p.ntx_bound2 > 0.f) {       // logits bounded a priori: no maximum pass (mx stays at the bound)
            float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) { cs0 += ptx::ex2_approx(fmaf(v[j], k2, -mx)); cs1 += ptx::ex2_approx(fmaf(v[j + 1], k2, -mx)); }
            sum += cs0 + cs1;
            return;
          }
          float cm0 = fmaxf(v[0], v[1]), cm1 = fmaxf(v[2], v[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 2) { cm0 = fmaxf(cm0, v[j]); cm1 = fmaxf(cm1, v[j + 1]); }
          const float nm = fmaxf(mx, fmaxf(cm0, cm1) * k2);
          if (nm > -INFINITY) {
            float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) { cs0 += ptx::ex2_approx(fmaf(v[j], k2, -nm)); cs1 += ptx::ex2_approx(fmaf(v[j + 1], k2, -nm)); }
            sum = fmaf(sum, ptx::ex2_approx(mx - nm), cs0 + cs1);
            mx = nm;
          }
        };
        float va[32], vb[32];
        if (blk_ok(0)) blk_load(0, va);
#pragma unroll
        for (int b = 0; b < NBLK; b += 2) {
          if (!blk_ok(b)) break;
          ptx::tmem_ld_wait_dep(va);
          if (b + 1 < NBLK && blk_ok(b + 1)) blk_load(b + 1, vb);
          blk_reduce(b, va);
          if (b + 1 >= NBLK || !blk_ok(b + 1)) break;
          ptx::tmem_ld_wait_dep(vb);
          if (b + 2 < NBLK && blk_ok(b + 2)) blk_load(b + 2, va);
          blk_reduce(b + 1, vb);
        }
        if (grow < p.M) {                                           // natural-log units for the merge kernel
          p.part_max[((size_t)n_tile * NSHARE + half) * p.M + grow] = mx * 0.6931471805599453f;
          p.part_sum[((size_t)n_tile * NSHARE + half) * p.M + grow] = sum;
        }
      } else if (EK == K_NTX_W && H16) {
        // Softmax-weight tile W[r][k] = P[r][k] + P[k][r] - 2 [k == pos(r)]  (nt_xent.py:53-65 differentiated) written as fp16,
        // scaled by 2^10 (weights are <= 2; the scale keeps the ~1/Rc entries out of the fp16 subnormals), for the fp16 dZ GEMM.
        // 32 columns per step (next step's TMEM loads in flight), thread = row; per element 2 FFMA + 2 EX2 + FADD + half a CVT.
        const long long gr = grow < p.row_split ? grow + p.row_offset : grow - p.row_split + p.row_offset2;
        const long long pos = ntx_pos(p, grow);
        const float k2 = p.inv_tau * 1.4426950408889634f;
        constexpr int NBLK = (BN / 32 + NSHARE - 1) / NSHARE;
        uint8_t* stg8 = reinterpret_cast<uint8_t*>(stg);             // 32 rows x 64 bytes, 16-byte chunk j of row r at j ^ ((r >> 1) & 3)
        __half* out16 = reinterpret_cast<__half*>(p.out16);
        auto blk_c0 = [&](int b) { return 32 * (half + b * NSHARE); };
        auto blk_ok = [&](int b) { return blk_c0(b) < BN && n0 + blk_c0(b) < p.N; };
        auto blk_load = [&](int b, float (&v)[32]) {
          ptx::tmem_ld_x16_nowait(taddr + blk_c0(b), v);
          ptx::tmem_ld_x16_nowait(taddr + blk_c0(b) + 16, v + 16);
        };
        auto blk_emit = [&](int b, float (&v)[32]) {
          const int c0 = blk_c0(b);
          const long long gc0 = (long long)n0 + c0 + p.col_offset;
          const unsigned long long d_self = (unsigned long long)(gr - gc0), d_pos = (unsigned long long)(pos - gc0);
          if (p.ntx_bound2 > 0.f) {
            // bounded logits: W 2^10 = 2^(t - bound) (2^(10 + bound - lse_r) + 2^(10 + bound - lse_k)): one EX2 per element
            const float nb = -p.ntx_bound2;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 ec = *reinterpret_cast<const float4*>(bias_t + c0 + j);     // broadcast
              v[j] = ptx::ex2_approx(fmaf(v[j], k2, nb)) * (nlr + ec.x);
              v[j + 1] = ptx::ex2_approx(fmaf(v[j + 1], k2, nb)) * (nlr + ec.y);
              v[j + 2] = ptx::ex2_approx(fmaf(v[j + 2], k2, nb)) * (nlr + ec.z);
              v[j + 3] = ptx::ex2_approx(fmaf(v[j + 3], k2, nb)) * (nlr + ec.w);
            }
          } else
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 nlc = *reinterpret_cast<const float4*>(bias_t + c0 + j);     // broadcast
            v[j] = ptx::ex2_approx(fmaf(v[j], k2, nlr)) + ptx::ex2_approx(fmaf(v[j], k2, nlc.x));
            v[j + 1] = ptx::ex2_approx(fmaf(v[j + 1], k2, nlr)) + ptx::ex2_approx(fmaf(v[j + 1], k2, nlc.y));
            v[j + 2] = ptx::ex2_approx(fmaf(v[j + 2], k2, nlr)) + ptx::ex2_approx(fmaf(v[j + 2], k2, nlc.z));
            v[j + 3] = ptx::ex2_approx(fmaf(v[j + 3], k2, nlr)) + ptx::ex2_approx(fmaf(v[j + 3], k2, nlc.w));
          }
          if (d_self < 32ull) {
            const int ds = (int)d_self;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j == ds ? 0.f : v[j];
          }
          if (d_pos < 32ull) {
            const int dp = (int)d_pos;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j == dp ? v[j] - 2048.f : v[j];
          }
          uint32_t h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            h[j] = *reinterpret_cast<const uint32_t*>(&hh);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<uint4*>(stg8 + lane * 64 + ((jj ^ ((lane >> 1) & 3)) << 4)) = make_uint4(h[4 * jj], h[4 * jj + 1], h[4 * jj + 2], h[4 * jj + 3]);
          __syncwarp();
          const int r_in = lane >> 2, ch = lane & 3, col = n0 + c0 + 8 * ch;
#pragma unroll
          for (int pass = 0; pass < 4; ++pass) {
            const int r = pass * 8 + r_in;
            const uint4 x = *reinterpret_cast<const uint4*>(stg8 + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4));
            if (r < rows_w && col < p.N) *reinterpret_cast<uint4*>(out16 + (size_t)(m0 + q * 32 + r) * p.ldo16 + col) = x;
          }
          __syncwarp();
        };
        float va[32], vb[32];
        if (blk_ok(0)) blk_load(0, va);
#pragma unroll
        for (int b = 0; b < NBLK; b += 2) {
          if (!blk_ok(b)) break;
          ptx::tmem_ld_wait_dep(va);
          if (b + 1 < NBLK && blk_ok(b + 1)) blk_load(b + 1, vb);
          blk_emit(b, va);
          if (b + 1 >= NBLK || !blk_ok(b + 1)) break;
          ptx::tmem_ld_wait_dep(vb);
          if (b + 2 < NBLK && blk_ok(b + 2)) blk_load(b + 2, va);
          blk_emit(b + 1, vb);
        }
      } else if (EK == K_ATOMIC) {
        for (int c0 = 16 * half; c0 < BN; c0 += 16 * NSHARE) {
          float v[16];
          ptx::tmem_ld_x16(taddr + c0, v);
          if (p.debug & 8) continue;                  // timing experiment: no atomics
          if (p.atomic_out == 3) {                    // ordered split-K: this split's partial tile, plain vector stores
            if (grow < p.M) {
              float* dst = p.out + ((size_t)(t / (n_tiles * m_tiles)) * p.M + grow) * p.ldo;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const int col = n0 + c0 + j;
                if (col < p.N) st_f4(dst + col, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
              }
            }
            continue;
          }
          if (p.atomic_out == 2) {                    // row-major output with 16-byte aligned rows: vector reductions (4x fewer L2 requests)
            if (grow < p.M)
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int col = n0 + c0 + j;
              if (col < p.N)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out + (size_t)grow * p.ldo + col), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            }
            continue;
          }
          if (grow < p.M) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = n0 + c0 + j;
              if (col < p.N)
                atomicAdd(p.transpose_out ? p.out + (size_t)col * p.ldo + grow : p.out + (size_t)grow * p.ldo + col, v[j]);
            }
          }
        }
      } else {
        // TMEM -> registers (thread = row) -> per-warp staging chunk -> coalesced row segments.  K_PLAIN applies
        // alpha / bias / ReLU / the ReLU bit mask in the register stage (and emits the ReLU bits of the result);
        // the other kinds apply their terms in the copy-out pass, where a lane owns one float4 column group.
        const int group = (m0 >> 5) + q;
        constexpr int LPR = CH / 4;            // lanes covering one row of the chunk (float4 each)
        constexpr int RPP = 32 / LPR;          // rows per copy-out pass
        const float alpha = p.alpha, floor_v = p.relu ? 0.f : -INFINITY;
        const bool has_bias = p.bias != nullptr, do_round = p.round_out != 0, has_out = p.out != nullptr,
                   has_out2 = p.out2 != nullptr, has_lo = p.out_lo != nullptr, has_stat = p.colstat != nullptr,
                   has_bits_in = p.bits_in != nullptr, has_bits_out = p.bits_out != nullptr;
        const int cq = 4 * (lane % LPR), r_in = lane / LPR;
        // One output, 32-column chunks: the staged chunk (32 rows x 128 bytes, XOR-swizzled = the 128-byte TMA swizzle; every warp's
        // staging area is 4 KB aligned) leaves through ONE bulk tensor store per chunk -- rows >= M and columns >= N are clipped by the
        // tensor map -- instead of 8 x (ld.shared + predicated st.global) per thread.
        const bool tma_out = EK == K_PLAIN && CH == 32 && p.tma_store != 0;
#pragma unroll 1
        for (int k = 0; k < NCHUNK; ++k) {
          const int c0 = CH * (half + k * NSHARE);
          if (c0 >= BN || n0 + c0 >= p.N) break;
          if (p.debug & 1) break;                     // timing experiment: no epilogue work at all
          float v[CH];
          ptx::tmem_ld_x16_nowait(taddr + c0, v);
          if (CH == 32) ptx::tmem_ld_x16_nowait(taddr + c0 + 16, v + (CH == 32 ? 16 : 0));
          uint32_t mw = mw_next;
          if (EK == K_PLAIN && !FOUR && has_bits_in && k + 1 < NCHUNK) mw_next = mask_word(k + 1);
          ptx::tmem_ld_wait();
          if (EK == K_NTX_W) {
            // W[r][k] = P[r][k] + P[k][r] - 2 [k == pos(r)],  P[i][k] = exp(l - lse_i) for k != i (nt_xent.py:53-65 differentiated)
            const long long gr = grow < p.row_split ? grow + p.row_offset : grow - p.row_split + p.row_offset2;
            const long long pos = ntx_pos(p, grow);
            const float k2 = p.inv_tau * 1.4426950408889634f;
            const float lr2 = (grow < p.M ? __ldg(p.row_lse + grow) : 0.f) * 1.4426950408889634f;
            const long long gc0 = n0 + c0 + p.col_offset;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              const float l2 = v[j] * k2;
              float w = exp2f(l2 - lr2) + exp2f(l2 - bias_t[c0 + j]);
              w = (gc0 + j == gr) ? 0.f : w;
              v[j] = (gc0 + j == pos) ? w - 2.f : w;
            }
          }
          if (EK == K_PLAIN) {
            if (has_bias || alpha != 1.f || p.relu) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                float4 b4 = f4_zero();
                if (has_bias) {
                  if (Cfg::BIAS_SMEM) b4 = *reinterpret_cast<const float4*>(bias_t + c0 + j);     // broadcast
                  else if (n0 + c0 + j < p.N) b4 = ldg_f4(p.bias + n0 + c0 + j);
                }
                v[j] = fmaxf(fmaf(v[j], alpha, b4.x), floor_v); v[j + 1] = fmaxf(fmaf(v[j + 1], alpha, b4.y), floor_v);
                v[j + 2] = fmaxf(fmaf(v[j + 2], alpha, b4.z), floor_v); v[j + 3] = fmaxf(fmaf(v[j + 3], alpha, b4.w), floor_v);
              }
            }
            if (!FOUR && has_bits_in) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = ((mw >> j) & 1u) ? v[j] : 0.f;
            }
            if (has_bits_out) {
              uint32_t ob = 0;
#pragma unroll
              for (int j = 0; j < CH; ++j) ob |= (v[j] > 0.f ? 1u : 0u) << j;
              if (grow < p.M) {
                const size_t w = (size_t)grow * p.ld_bits + ((n0 + c0) >> 5);
                if (CH == 32) p.bits_out[w] = ob;
                else reinterpret_cast<uint16_t*>(p.bits_out)[2 * w + (((n0 + c0) >> 4) & 1)] = (uint16_t)ob;
              }
            }
          }
          if (tma_out) {
            if (lane == 0) ptx::bulk_wait_group_read0();      // the previous chunk's store has finished reading this staging area
            __syncwarp();
            if (CH == 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                st_f4(stg + stg_off<LD>(lane, j), do_round ? f4_tf32(make_float4(v[j], v[j + 1], v[j + 2], v[j + 3])) : make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
              ptx::fence_proxy_async_smem();                  // generic-proxy writes -> visible to the async proxy (TMA)
              __syncwarp();
              if (lane == 0 && rows_w > 0) { ptx::tma_store_2d(&tmO, stg, n0 + c0, m0 + q * 32); ptx::bulk_commit_group(); }
              if (has_stat) colstat_regs(reinterpret_cast<const float (&)[32]>(v), rows_w, n0 + c0, group, p, lane);      // of the exact values
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < CH; j += 4) st_f4(stg + stg_off<LD>(lane, j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          __syncwarp();
          const int col = n0 + c0 + cq;
          if ((EK == K_PLAIN || EK == K_NTX_W) && rows_w == 32 && n0 + c0 + CH <= p.N) {
            // full chunk: straight-line copy-out (predicated stores only)
            const size_t gr0 = (size_t)(m0 + q * 32 + r_in);
            float* po = p.out + gr0 * p.ldo + col;
            float* po2 = p.out2 + gr0 * p.ldo2 + col;
            float* pl = p.out_lo + gr0 * p.ldo_lo + col;
#pragma unroll
            for (int pass = 0; pass < 32 / RPP; ++pass) {
              const float4 x = *reinterpret_cast<const float4*>(stg + stg_off<LD>(pass * RPP + r_in, cq));
              const float4 xr = f4_tf32(x);
              if (has_out) st_f4(po + (size_t)pass * RPP * p.ldo, do_round ? xr : x);
              if (has_out2) st_f4(po2 + (size_t)pass * RPP * p.ldo2, xr);
              if (has_lo) st_f4(pl + (size_t)pass * RPP * p.ldo_lo, f4_tf32_residual(x));
            }
          } else {
            const bool col_ok = col < p.N;
#pragma unroll 1
            for (int pass = 0; pass < 32 / RPP; ++pass) {
              const int r = pass * RPP + r_in;
              if (r < rows_w && col_ok) {
                const int gr = m0 + q * 32 + r;
                float4 x = *reinterpret_cast<const float4*>(stg + stg_off<LD>(r, cq));
                if (EK != K_PLAIN && EK != K_NTX_W) {
                  x = epilogue_apply(x, p, gr, col);
                  if (has_stat) st_f4(stg + stg_off<LD>(r, cq), x);
                }
                const float4 xr = f4_tf32(x);
                if (has_out) st_f4(p.out + (size_t)gr * p.ldo + col, do_round ? xr : x);
                if (has_out2) st_f4(p.out2 + (size_t)gr * p.ldo2 + col, xr);
                if (has_lo) st_f4(p.out_lo + (size_t)gr * p.ldo_lo + col, f4_tf32_residual(x));
              }
            }
          }
          if (has_stat) {
            __syncwarp();
            colstat_warp<CH, LD>(stg, rows_w, n0 + c0, group, p, lane);
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO) ptx::mbar_arrive_cluster(tempty_leader + buf * 8u); else ptx::mbar_arrive(tempty_bar + buf);
      }
    }
    if (p.tma_store && lane == 0) ptx::bulk_wait_group0();      // this thread's bulk stores are complete before the CTA exits
  }
  ptx::tc_fence_before();
  if (TWO) ptx::cluster_sync(); else __syncthreads();     // nobody leaves (or frees TMEM) while the peer may still signal / be read
  if (warp == 1) {
    ptx::tc_fence_after();
    if (TWO) ptx::tmem_dealloc_2cta(tmem_base, GEMM_TMEM_COLS); else ptx::tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Debug implementation: plain fp32 FMA, one thread per output row, 32 columns per block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gemm_simt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                        long long ldb, const float* __restrict__ A_lo, const float* __restrict__ B_lo,
                                                        const GemmParams p) {
  pdl_sync();
  __shared__ float stage[GEMM_BM * 33];
  const int n0 = blockIdx.x * 32, m_tile = blockIdx.y, m0 = m_tile * GEMM_BM;
  const int row = threadIdx.x, grow = m0 + row;
  const int rows_valid = min(GEMM_BM, p.M - m0);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  if (grow < p.M) {
    for (int k = 0; k < p.K; ++k) {
      const size_t ai = p.a_mn ? (size_t)k * lda + grow : (size_t)grow * lda + k;
      const float a = A[ai] + ((p.segments > 1 && A_lo) ? A_lo[ai] : 0.f);   // (compensate = 1: A and B are the full values already)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = n0 + j;
        if (col < p.N) {
          const size_t bi = p.b_mn ? (size_t)k * ldb + col : (size_t)col * ldb + k;
          acc[j] = fmaf(a, B[bi] + ((p.segments > 1 && B_lo) ? B_lo[bi] : 0.f), acc[j]);
        }
      }
    }
  }
  if (p.epi == EPI_NTX_FWD) {
    if (grow < p.M) {
      const long long gr = grow < p.row_split ? grow + p.row_offset : grow - p.row_split + p.row_offset2;
      const long long pos = ntx_pos(p, grow);
      float mx = -INFINITY, sum = 0.f;
      for (int j = 0; j < 32; ++j) {
        const long long gc = n0 + j + p.col_offset;
        if (n0 + j < p.N && gc != gr) mx = fmaxf(mx, acc[j] * p.inv_tau);
        if (n0 + j < p.N && gc == pos) p.row_pos[grow] = acc[j] * p.inv_tau;
      }
      for (int j = 0; j < 32; ++j) {
        const long long gc = n0 + j + p.col_offset;
        if (n0 + j < p.N && gc != gr) sum += __expf(acc[j] * p.inv_tau - mx);
      }
      p.part_max[(size_t)blockIdx.x * p.M + grow] = mx;
      p.part_sum[(size_t)blockIdx.x * p.M + grow] = sum;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const int col = n0 + j;
    float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    if (grow < p.M && col < p.N) {
      if (p.atomic_out) {
        const float vv[4] = {v.x, v.y, v.z, v.w};
        for (int t = 0; t < 4; ++t)
          atomicAdd(p.transpose_out ? p.out + (size_t)(col + t) * p.ldo + grow : p.out + (size_t)grow * p.ldo + col + t, vv[t]);
      } else {
        v = epilogue_apply(v, p, grow, col);
        if (p.out) st_f4(p.out + (size_t)grow * p.ldo + col, p.round_out ? f4_tf32(v) : v);
        if (p.out2) st_f4(p.out2 + (size_t)grow * p.ldo2 + col, f4_tf32(v));
        if (p.out_lo) st_f4(p.out_lo + (size_t)grow * p.ldo_lo + col, f4_tf32_residual(v));
      }
    }
    stage[row * 33 + j] = v.x; stage[row * 33 + j + 1] = v.y; stage[row * 33 + j + 2] = v.z; stage[row * 33 + j + 3] = v.w;
  }
  if (p.colstat && !p.atomic_out) {
    __syncthreads();
    for (int g = 0; g < 4; ++g)
      colstat_group(stage + g * 32 * 33, 33, 32, max(0, min(32, rows_valid - 32 * g)), n0, m_tile * 4 + g, p, threadIdx.x, 128);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Tensor map over a row-major matrix [outer][inner] of fp32 (or fp16: `half16`, K-major only) with leading dimension ld
// (elements), box = [box_outer][bk inner elements] = rows of 128 bytes (bk = 32 fp32 / 64 fp16), 128-byte swizzle, zero fill
// out of bounds.
static int make_tmap(CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_outer, bool mn_major, int bk,
                     bool half16 = false) {
  // cuTensorMapEncodeTiled is a driver-API call: make sure this thread (e.g. the autograd engine's) has the
  // primary context bound before the first one.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  EncodeTiledFn enc = encode_tiled_fn();
  MOLCLR_REQUIRE(enc != nullptr, "gemm: cuTensorMapEncodeTiled not available from the driver");
  const int esize = half16 ? 2 : 4;
  MOLCLR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "gemm: operand base pointer must be 16-byte aligned");
  MOLCLR_REQUIRE(ld % (16 / esize) == 0, "gemm: leading dimension %lld must be a multiple of %d elements (TMA 16-byte stride)", (long long)ld, 16 / esize);
  MOLCLR_REQUIRE(!half16 || !mn_major, "gemm: fp16 operands are K-major only");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  // K-major: box = [box_outer rows][bk k];  MN-major: box = [bk k rows][32 mn]
  cuuint32_t box[2] = {mn_major ? 32u : (cuuint32_t)bk, (cuuint32_t)(mn_major ? bk : box_outer)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, half16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (bk * esize == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MOLCLR_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld outer=%lld ld=%lld)", (int)r,
                 (long long)inner, (long long)outer, (long long)ld);
  return 0;
}

static bool gemm_no_mn3d();

// 3-D tensor map over an MN-major operand [K][MN] (row pitch ld floats): dims {32 mn of a block, K, blocks of 32 mn}, box
// {32, bk, nblk}: one copy lands nblk consecutive [bk k][32 mn] blocks in shared memory, each in the 32B-atom 128-byte swizzle
// the tensor core wants.  Blocks beyond ceil(MN / 32) are zero-filled; the last block may read up to 31 floats of the row's
// own padding (the caller checks ld >= MN rounded up to 32).
static int make_tmap_mn3d(CUtensorMap* m, const float* base, int64_t mn, int64_t k, int64_t ld, int bk, int nblk) {
  EncodeTiledFn enc = encode_tiled_fn();
  MOLCLR_REQUIRE(enc != nullptr, "gemm: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {32u, (cuuint64_t)k, (cuuint64_t)((mn + 31) / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128u};
  cuuint32_t box[3] = {32u, (cuuint32_t)bk, (cuuint32_t)nblk};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MOLCLR_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled (3-D) failed with CUresult %d (mn=%lld k=%lld ld=%lld nblk=%d)", (int)r,
                 (long long)mn, (long long)k, (long long)ld, nblk);
  return 0;
}

// Tensor map over a row-major fp32 OUTPUT [rows][cols] (row pitch ld floats): boxes of 32 rows x 32 columns, 128-byte swizzle -- the layout
// of an epilogue staging chunk.  Stores clip at the extents, so ragged M / N need no masking.
static int make_tmap_out(CUtensorMap* m, float* base, int64_t cols, int64_t rows, int64_t ld) {
  EncodeTiledFn enc = encode_tiled_fn();
  MOLCLR_REQUIRE(enc != nullptr, "gemm: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MOLCLR_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled (output) failed with CUresult %d (cols=%lld rows=%lld ld=%lld)", (int)r,
                 (long long)cols, (long long)rows, (long long)ld);
  return 0;
}

static bool gemm_tma_store();

template <int BN, bool FOUR, int KIND, bool TWO>
static int launch_tc(const GemmJob& j, GemmParams p, int n_tiles, int m_tiles, int splits, cudaStream_t stream) {
  constexpr bool H3 = KIND == K_PLAIN_H3;
  using Cfg = GemmCfg<BN, FOUR, TWO, H3>;
  p.n_tiles = n_tiles; p.m_tiles = m_tiles; p.splits = splits;
  MOLCLR_REQUIRE(!p.b_mn || Cfg::BN_CTA % 32 == 0, "gemm: internal: MN-major B needs whole 32-column blocks per CTA (BN=%d)", BN);
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  // K-major operand [rows][K]: inner = K, outer = rows, box = rows-per-CTA x BK.  MN-major [K][rows]: inner = rows, outer = K, box BK x 32.
  const bool h16 = p.half16 != 0;
  const int bk_el = h16 ? 2 * Cfg::BK : Cfg::BK;          // elements per 128-byte tile row
  rc = p.a_mn ? make_tmap(&tmA, j.A, p.M, p.K, j.lda, 32, true, Cfg::BK) : make_tmap(&tmA, j.A, p.K, p.M, j.lda, GEMM_BM, false, bk_el, h16);
  if (rc) return rc;
  if (H3) tmB = tmA;            // (no fp32 B tile: the slot is unused)
  else rc = p.b_mn ? make_tmap(&tmB, j.B, p.N, p.K, j.ldb, 32, true, Cfg::BK) : make_tmap(&tmB, j.B, p.K, p.N, j.ldb, Cfg::BN_CTA, false, bk_el, h16);
  if (rc) return rc;
  tmA2 = tmA; tmB2 = tmB;
  p.a_mn3d = p.b_mn3d = 0;
  if (p.segments == 1 && !gemm_no_mn3d()) {
    if (p.a_mn && j.lda >= (p.M + 31) / 32 * 32) {
      rc = make_tmap_mn3d(&tmA2, j.A, p.M, p.K, j.lda, Cfg::BK, GEMM_BM / 32);
      if (rc) return rc;
      p.a_mn3d = 1;
    }
    constexpr int J1 = (TWO ? Cfg::N1 / 2 : Cfg::N1) / 32;
    if (p.b_mn && J1 > 0 && j.ldb >= (p.N + 31) / 32 * 32) {
      rc = make_tmap_mn3d(&tmB2, j.B, p.N, p.K, j.ldb, Cfg::BK, J1);
      if (rc) return rc;
      p.b_mn3d = 1;
    }
  }
  if (p.segments > 1) {
    if (j.A_lo) {
      rc = p.a_mn ? make_tmap(&tmA2, j.A_lo, p.M, p.K, j.lda, 32, true, Cfg::BK) : make_tmap(&tmA2, j.A_lo, p.K, p.M, j.lda, GEMM_BM, false, Cfg::BK);
      if (rc) return rc;
    }
    if (p.derive_lo < 2)             // (mixed / fp16 three-product: no second fp32 B tensor)
      rc = p.b_mn ? make_tmap(&tmB2, j.B_lo, p.N, p.K, j.ldb, 32, true, Cfg::BK) : make_tmap(&tmB2, j.B_lo, p.K, p.N, j.ldb, Cfg::BN_CTA, false, Cfg::BK);
    else if (p.b_presplit)           // 16-bit [2 rows16][K]: boxes of BN_CTA rows x 32 elements (64-byte rows, 64-byte swizzle, zero fill beyond K)
      rc = make_tmap(&tmB2, reinterpret_cast<const float*>(j.B16), p.K, 2ll * j.rows16, j.ld16, Cfg::BN_CTA, false, Cfg::BK, true);
    if (rc) return rc;
  }
  CUtensorMap tmO = tmA;
  p.tma_store = 0;
  constexpr bool kPlain = KIND == K_PLAIN || KIND == K_PLAIN16 || KIND == K_PLAIN_H3;
  if (kPlain && Cfg::CHUNK == 32 && p.out && !p.out2 && !p.out_lo && !p.transpose_out && gemm_tma_store() && p.ldo % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
    rc = make_tmap_out(&tmO, p.out, p.N, p.M, p.ldo);
    if (rc) return rc;
    p.tma_store = 1;
  }
  auto kernel = gemm_tf32_kernel<BN, FOUR, KIND, TWO>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "gemm: cudaFuncSetAttribute");
    attr_set = true;
  }
  const long long total = (long long)n_tiles * m_tiles * splits;
  const int workers = gemm_workers(TWO);
  const int active = (int)(total < workers ? total : workers);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(TWO ? 2 * active : active), 1, 1);
  cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TWO ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see common.cuh: launch_kernel / pdl_sync
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmA2, tmB2, tmO, p);
  ++g_launches;
  if (e != cudaSuccess) return cuda_fail(e, "gemm_tf32 launch");
  return 0;
}

static bool gemm_no_mn3d() {          // MOLCLR_GEMM_MN3D=0: per-block 2-D copies for MN-major operands (A/B timing)
  static int v = -1;
  if (v < 0) { const char* e = debug_env("MOLCLR_GEMM_MN3D"); v = (e && atoi(e) == 0) ? 1 : 0; }
  return v != 0;
}

// epilogue through TMA stores unless MOLCLR_GEMM_TMA_STORE=0 (debug-switch builds: A/B timing)
static bool gemm_tma_store() {
  static int v = -1;
  if (v < 0) { const char* e = debug_env("MOLCLR_GEMM_TMA_STORE"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

static int gemm_debug_flags() {
  static int v = -1;
  if (v < 0) { const char* e = debug_env("MOLCLR_GEMM_DEBUG"); v = e ? atoi(e) : 0; }
  return v;
}

static int gemm_impl_simt() {
  static int v = -1;
  if (v < 0) { const char* e = debug_env("MOLCLR_GEMM_IMPL"); v = (e && strcmp(e, "simt") == 0) ? 1 : 0; }
  return v;
}

// CTA pairs (cluster of 2, 256-row tiles) unless MOLCLR_GEMM_PAIR=0
static bool gemm_pair() {
  static int v = -1;
  if (v < 0) { const char* e = debug_env("MOLCLR_GEMM_PAIR"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

// Column-tile width.  Pairs: each CTA stages BN/2 rows of B; an MN-major B needs BN/2 to be whole 32-column blocks.
static int gemm_bn(long long N, bool b_mn, bool pair, bool allow224 = false) {
  if (!pair) return (N % 256 == 0 || N > 640) ? 256 : 160;
  // fewest column tiles first (every extra tile re-reads the A rows), then the narrowest width that still covers N
  // (the compensated forward GEMM is tensor-bound: N = 600 as 3 x 224 instead of 3 x 256 saves 12% of its MMAs)
  const long long nt = (N + 255) / 256;
  for (int bn : {160, 192, 224, 256})
    if (nt * bn >= N && (!b_mn || (bn / 2) % 32 == 0) && (bn != 224 || allow224)) return bn;
  return 256;
}

bool gemm_f16_ok() { return !gemm_impl_simt(); }

int gemm_make_tmap_f16(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  return make_tmap(m, reinterpret_cast<const float*>(base), inner, outer, ld, box_outer, false, 64, true);
}

int gemm_n_tiles(long long N) {     // number of NT-Xent forward partials per row
  if (gemm_impl_simt()) return (int)((N + 31) / 32);
  const int bn = gemm_bn(N, false, gemm_pair());
  return 2 * (int)((N + bn - 1) / bn);          // two epilogue warps share the columns of a tile
}

// Number of K splits a launch really uses for a requested split_k (no empty split).
static int effective_splits(int num_kb, int split_k) {
  int splits = split_k > 1 ? split_k : 1;
  if (splits > num_kb) splits = num_kb;
  const int per = (num_kb + splits - 1) / splits;
  return (num_kb + per - 1) / per;
}

int gemm_run(const GemmJob& job, cudaStream_t stream) {
  GemmParams p = job.p;
  MOLCLR_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  MOLCLR_REQUIRE(p.N % 4 == 0, "gemm: N=%d must be a multiple of 4", p.N);
  const bool atomic = job.split_k > 1 || p.transpose_out;
  MOLCLR_REQUIRE(job.A_lo == nullptr || job.B_lo != nullptr, "gemm: A_lo needs B_lo");
  MOLCLR_REQUIRE(job.compensate >= 0 && job.compensate <= 2, "gemm: compensate must be 0, 1 or 2");
  MOLCLR_REQUIRE(!job.compensate || (!job.A_lo && !job.B_lo && !p.a_mn && !p.b_mn), "gemm: compensate takes unrounded K-major A and B and no lo tensors");
  const bool h3 = job.compensate == 2;
  MOLCLR_REQUIRE(!h3 || job.B16, "gemm: compensate = 2 needs B16, the fp16 tiles of B (molclr_prepare_weights, b16_kind = 1)");
  MOLCLR_REQUIRE(h3 || job.B, "gemm: B is null");
  p.segments = (job.B_lo || job.compensate) ? 3 : 1;
  // compensated product with the low halves derived on chip: 1 = A_lo (fp32 tile) from an unrounded A, B_hi/B_lo from the caller;
  // 2 = "mixed": both operands unrounded, bf16 correction tiles of both formed on chip
  p.derive_lo = h3 ? 3 : job.compensate ? 2 : (job.B_lo && !job.A_lo) ? 1 : 0;
  p.status = job.status;
  if (h3) p.alpha *= 1.f / (float)MOLCLR_H16_SCALE;      // B16 holds the weights times 2^6 (exact); the epilogue undoes it
  p.b_presplit = (job.compensate && job.B16) ? 1 : 0;
  p.rows16 = job.rows16;
  MOLCLR_REQUIRE(!job.B16 || job.compensate, "gemm: B16 (pre-split bf16 tiles of B) belongs to the compensated product");
  if (p.b_presplit)
    MOLCLR_REQUIRE(job.ld16 % 8 == 0 && job.ld16 >= p.K && job.rows16 >= (p.N + 255) / 256 * 256 && (reinterpret_cast<uintptr_t>(job.B16) & 15) == 0,
                   "gemm: B16 needs ld16 %% 8 == 0, ld16 >= K, rows16 >= N rounded up to 256, a 16-byte aligned base");
  MOLCLR_REQUIRE((!p.bits_in && !p.bits_out) || (p.epi == EPI_GENERIC && !atomic && !p.mask && !p.addend),
                 "gemm: ReLU bit masks need the plain epilogue (no float mask / addend / split-K)");
  MOLCLR_REQUIRE(!p.bits_in || p.segments == 1, "gemm: mask_bits is not supported by the compensated product");
  if (p.bits_in || p.bits_out)
    MOLCLR_REQUIRE(p.ld_bits >= molclr_gemm_mask_words(p.N), "gemm: ld_bits=%lld < molclr_gemm_mask_words(N)=%d", (long long)p.ld_bits,
                   molclr_gemm_mask_words(p.N));
  if (atomic)
    MOLCLR_REQUIRE(p.segments == 1 && !p.out_lo && p.epi == EPI_GENERIC && !p.bias && !p.addend && !p.mask && !p.relu && !p.round_out && !p.out2 && !p.colstat && p.out &&
                       p.alpha == 1.f,
                   "gemm: split-K / transposed output supports no fused epilogue");
  else if (p.epi != EPI_NTX_FWD) {
    MOLCLR_REQUIRE(p.out != nullptr || p.out2 != nullptr || p.colstat != nullptr || p.out16 != nullptr, "gemm: no output requested");
    MOLCLR_REQUIRE((!p.out || p.ldo % 4 == 0) && (!p.out2 || p.ldo2 % 4 == 0), "gemm: output leading dimensions must be multiples of 4");
    MOLCLR_REQUIRE((!p.addend || p.ldadd % 4 == 0) && (!p.mask || p.ldmask % 4 == 0), "gemm: addend/mask leading dimensions must be multiples of 4");
  }
  if (p.half16)
    MOLCLR_REQUIRE(p.segments == 1 && !p.a_mn && !p.b_mn && !atomic && !gemm_impl_simt(), "gemm: fp16 operands: single pass, K-major, no split-K");
  MOLCLR_REQUIRE(!p.out16 || (p.epi == EPI_NTX_W && p.ldo16 % 8 == 0 && !p.out), "gemm: out16 is the fp16 output of the NT-Xent weight epilogue");
  const int bk = p.half16 ? 2 * GEMM_BK : p.segments > 1 ? GEMM_BK4 : GEMM_BK;
  p.num_kb = (p.K + bk - 1) / bk;
  const int splits = effective_splits(p.num_kb, job.split_k);
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  p.atomic_out = !atomic ? 0 : (!p.transpose_out && p.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) ? 2 : 1;
  if (job.ordered_ws) {           // ordered split-K: partial tiles to the workspace (the caller sums them in split order)
    MOLCLR_REQUIRE(atomic && !p.transpose_out && p.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(job.ordered_ws) & 15) == 0,
                   "gemm: ordered split-K writes [splits][M][ldo] partials (no transposed output), 16-byte aligned");
    p.out = job.ordered_ws;
    p.atomic_out = 3;
  }
  p.debug = gemm_debug_flags();
  // L2 prefetch of the A operand 12 k-blocks ahead of the loads (row products with one K split and a K-major A): OFF.  Tried at the end
  // of round 2 on the theory that the HBM-streamed A operands (u, g_u) make the ring latency-bound: the step got SLOWER, 10.31 ->
  // 11.09 ms (backward row products 111 -> 141 us, forward 142 -> 152 us in situ; tools/gpu_r2_pf.sh).  Kept behind
  // MOLCLR_GEMM_DEBUG bit 32 (debug-switch builds) for further experiments with the distance.
  p.pf = (splits == 1 && !p.a_mn && (p.segments == 1 || job.compensate) && (p.debug & 32)) ? 12 : 0;
  // split-K weight gradients stay on single CTAs: their 300/600-wide outputs pad badly to 256-row pair tiles (MMA-bound)
  const bool wide = atomic && job.wide && gemm_pair() && !gemm_impl_simt();
  if (wide) MOLCLR_REQUIRE(p.a_mn && p.b_mn, "gemm: wide split-K tiles need both operands MN-major");
  const bool pair = gemm_pair() && !gemm_impl_simt() && (!atomic || wide);
  const int tile_m = pair ? 2 * GEMM_BM : GEMM_BM;
  const int m_tiles = (p.M + tile_m - 1) / tile_m;
  p.stat_groups = molclr_gemm_colstat_tiles(p.M);
  if (atomic && p.atomic_out != 3 && !job.accumulate) {
    const size_t w = (size_t)(p.transpose_out ? p.M : p.N) * sizeof(float), h = (size_t)(p.transpose_out ? p.N : p.M);
    cudaError_t e = cudaMemset2DAsync(p.out, (size_t)p.ldo * sizeof(float), 0, w, h, stream);
    if (e != cudaSuccess) return cuda_fail(e, "gemm: zeroing split-K output");
  }
  if (gemm_impl_simt()) {
    p.kb_per_split = p.num_kb;
    MOLCLR_LAUNCH(gemm_simt_kernel, dim3((unsigned)((p.N + 31) / 32), (p.M + GEMM_BM - 1) / GEMM_BM, 1), 128, 0, stream, job.A, job.lda, job.B, job.ldb, job.A_lo, job.B_lo, p);
    MOLCLR_CHECK_LAUNCH("gemm_simt");
    return 0;
  }
  const int kind = p.epi == EPI_NTX_FWD ? K_NTX_FWD : atomic ? K_ATOMIC : p.epi == EPI_NTX_W ? K_NTX_W
                   : (p.mask || p.addend) ? K_LATE : K_PLAIN;
  MOLCLR_REQUIRE(p.segments == 1 || kind == K_PLAIN, "gemm: the compensated product supports the plain epilogue only");
  MOLCLR_REQUIRE(!p.half16 || kind == K_PLAIN || kind == K_NTX_W || kind == K_NTX_FWD, "gemm: fp16 operands: plain and NT-Xent epilogues only");
  MOLCLR_REQUIRE(kind != K_NTX_W || (p.out16 != nullptr) == (p.half16 != 0), "gemm: the NT-Xent weight epilogue writes fp16 iff its operands are fp16");
  MOLCLR_REQUIRE(!h3 || !gemm_impl_simt(), "gemm: compensate = 2 has no scalar debug implementation");
  const int kind_i = h3 ? K_PLAIN_H3 : !p.half16 ? kind : kind == K_PLAIN ? K_PLAIN16 : kind == K_NTX_W ? K_NTX_W16 : K_NTX_FWD16;
  const int bn = wide ? 320 : (job.bn_hint && pair) ? job.bn_hint : gemm_bn(p.N, p.b_mn != 0, pair, kind == K_PLAIN || kind == K_LATE), nt = (p.N + bn - 1) / bn;
  if (wide) MOLCLR_REQUIRE((long long)nt * m_tiles * splits <= gemm_workers(true), "gemm: a wide split-K launch must be one wave");
#define MOLCLR_GEMM_CASE(BN_, FOUR_, KIND_, TWO_) \
  if (bn == BN_ && (p.segments > 1) == FOUR_ && kind_i == KIND_ && pair == TWO_) return launch_tc<BN_, FOUR_, KIND_, TWO_>(job, p, nt, m_tiles, splits, stream);
#define MOLCLR_GEMM_KINDS(BN_, TWO_) \
  MOLCLR_GEMM_CASE(BN_, false, K_PLAIN, TWO_) MOLCLR_GEMM_CASE(BN_, true, K_PLAIN, TWO_) MOLCLR_GEMM_CASE(BN_, false, K_LATE, TWO_) \
  MOLCLR_GEMM_CASE(BN_, false, K_NTX_W, TWO_) MOLCLR_GEMM_CASE(BN_, false, K_NTX_FWD, TWO_) MOLCLR_GEMM_CASE(BN_, false, K_ATOMIC, TWO_)
  MOLCLR_GEMM_KINDS(160, false) MOLCLR_GEMM_KINDS(256, false)
  MOLCLR_GEMM_KINDS(160, true) MOLCLR_GEMM_KINDS(192, true) MOLCLR_GEMM_KINDS(256, true)
  MOLCLR_GEMM_CASE(224, false, K_PLAIN, true) MOLCLR_GEMM_CASE(224, true, K_PLAIN, true) MOLCLR_GEMM_CASE(224, false, K_LATE, true)
  MOLCLR_GEMM_CASE(128, false, K_PLAIN, true) MOLCLR_GEMM_CASE(320, false, K_ATOMIC, true)
  MOLCLR_GEMM_CASE(160, true, K_PLAIN_H3, true) MOLCLR_GEMM_CASE(192, true, K_PLAIN_H3, true) MOLCLR_GEMM_CASE(224, true, K_PLAIN_H3, true)
  MOLCLR_GEMM_CASE(256, true, K_PLAIN_H3, true) MOLCLR_GEMM_CASE(160, true, K_PLAIN_H3, false) MOLCLR_GEMM_CASE(256, true, K_PLAIN_H3, false)
#define MOLCLR_GEMM_KINDS16(BN_, TWO_) \
  MOLCLR_GEMM_CASE(BN_, false, K_PLAIN16, TWO_) MOLCLR_GEMM_CASE(BN_, false, K_NTX_W16, TWO_) MOLCLR_GEMM_CASE(BN_, false, K_NTX_FWD16, TWO_)
  MOLCLR_GEMM_KINDS16(160, false) MOLCLR_GEMM_KINDS16(256, false)
  MOLCLR_GEMM_KINDS16(160, true) MOLCLR_GEMM_KINDS16(192, true) MOLCLR_GEMM_KINDS16(256, true) MOLCLR_GEMM_CASE(128, false, K_PLAIN16, true)
#undef MOLCLR_GEMM_KINDS16
#undef MOLCLR_GEMM_KINDS
#undef MOLCLR_GEMM_CASE
  set_error("gemm: no kernel instance for bn=%d segments=%d kind=%d", bn, p.segments, kind_i);
  return -2;
}

}  // namespace molclr

using namespace molclr;

extern "C" int molclr_gemm_colstat_tiles(int64_t M) { return 4 * (int)((M + GEMM_BM - 1) / GEMM_BM); }
extern "C" int molclr_gemm_colstat_tile_rows(void) { return GEMM_STAT_ROWS; }
// 32-bit words per row of a ReLU bit mask over N columns (covers the column tiles the kernel will use)
extern "C" int molclr_gemm_mask_words(int64_t N) {
  int words = 0;
  for (int bn : {160, 192, 224, 256}) { const int w = (int)((N + bn - 1) / bn) * (bn / 32); if (w > words) words = w; }
  return words;
}
// Work decomposition the launcher will use (for callers that size a split-K): number of output tiles of an [M][N] product
// and the number of concurrent workers (CTAs, or CTA pairs).
extern "C" int molclr_gemm_tile_count(int64_t M, int64_t N, int b_mn) {
  const int bn = gemm_bn(N, b_mn != 0, false);
  return (int)(((M + GEMM_BM - 1) / GEMM_BM) * ((N + bn - 1) / bn));
}
extern "C" int molclr_gemm_workers(void) { return gemm_workers(false); }

// dW [O][I] = dY^T X for row-major dY [R][O], X [R][I]: the weight gradient of a Linear.  Both operands are consumed MN-major
// in place; the reduction over R is split across one wave of workers and accumulated atomically (molclr_gemm_dw) or written as
// per-split partial tiles that are then summed in split order (molclr_gemm_dw_ordered: bit-reproducible).  Orientation (which
// operand is "M") and tile shape (128 x 160/256 on single CTAs, or 256 x 320 on CTA pairs) are chosen to minimise the operand
// bytes every K row costs on the L2 -> shared-memory path, n_tiles * M + m_tiles * N, which is what bounds this kernel.
struct DwPlan { int swap, wide, split, splits; long long M, N; };

static int dw_plan(int64_t R, int64_t O, int64_t I, DwPlan* out) {
  MOLCLR_REQUIRE(R > 0 && O > 0 && I > 0 && R < (1ll << 31) && O < (1ll << 31) && I < (1ll << 31), "gemm_dw: bad extents");
  long long best = -1;
  int best_swap = 0, best_wide = 0, best_tiles = 1;
  for (int swap = 0; swap < 2; ++swap)
    for (int wide = 0; wide < 2; ++wide) {
      const long long M = swap ? I : O, N = swap ? O : I;
      if (wide && (!gemm_pair() || gemm_impl_simt() || M <= GEMM_BM)) continue;
      const int bn = wide ? 320 : gemm_bn(N, true, false), tm = wide ? 2 * GEMM_BM : GEMM_BM;
      const long long mt = (M + tm - 1) / tm, nt = (N + bn - 1) / bn, cost = nt * M + mt * N;
      if (mt * nt > gemm_workers(wide != 0)) continue;
      if (best < 0 || cost < best) { best = cost; best_swap = swap; best_wide = wide; best_tiles = (int)(mt * nt); }
    }
  MOLCLR_REQUIRE(best >= 0, "gemm_dw: output [%lld x %lld] has more tiles than the GPU has workers", (long long)O, (long long)I);
  const int num_kb = (int)((R + GEMM_BK - 1) / GEMM_BK);
  int split = gemm_workers(best_wide != 0) / best_tiles;
  if (split > num_kb / 8) split = num_kb / 8;
  if (split < 1) split = 1;
  // `splits`: what the ORDERED variant uses -- it always takes the split-K kernel kind (split_k >= 2), also for a single K split
  out->swap = best_swap; out->wide = best_wide; out->split = split; out->splits = effective_splits(num_kb, split < 2 ? 2 : split);
  out->M = best_swap ? I : O; out->N = best_swap ? O : I;
  return 0;
}

static void dw_job(GemmJob& j, const DwPlan& pl, const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R) {
  memset(&j, 0, sizeof(j));
  j.A = pl.swap ? X : dY; j.lda = pl.swap ? ldx : ldy;
  j.B = pl.swap ? dY : X; j.ldb = pl.swap ? ldy : ldx;
  j.p.M = (int)pl.M; j.p.N = (int)pl.N; j.p.K = (int)R; j.p.a_mn = 1; j.p.b_mn = 1;
  j.p.alpha = 1.f; j.p.epi = EPI_GENERIC;
  j.split_k = pl.split; j.wide = pl.wide;
}

static int gemm_dw_atomic(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW, int64_t ldw,
                          int accumulate, cudaStream_t stream) {
  DwPlan pl;
  if (int rc = dw_plan(R, O, I, &pl)) return rc;
  GemmJob j;
  dw_job(j, pl, dY, ldy, X, ldx, R);
  j.p.out = dW; j.p.ldo = ldw; j.p.transpose_out = pl.swap;
  if (accumulate) {
    j.accumulate = 1;
    if (j.split_k < 2 && !j.p.transpose_out) j.split_k = 2;       // the accumulating (atomic) kernel kind, also for a single K split
  }
  return gemm_run(j, stream);
}

extern "C" int molclr_gemm_dw(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW,
                              int64_t ldw, cudaStream_t stream) {
  return gemm_dw_atomic(dY, ldy, X, ldx, R, O, I, dW, ldw, 0, stream);
}

extern "C" int molclr_gemm_dw_acc(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW,
                                  int64_t ldw, cudaStream_t stream) {
  return gemm_dw_atomic(dY, ldy, X, ldx, R, O, I, dW, ldw, 1, stream);
}

// out[m][n] (or out[n][m]) = sum_s ws[s][m][n], s in increasing order.  32 x 32 tiles, block (32, 8).
__global__ void __launch_bounds__(256) dw_reduce_kernel(const float* __restrict__ ws, int S, int M, int N, long long ldws, float* __restrict__ out,
                                                        long long ldo, int transpose) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int m = blockIdx.y * 32 + threadIdx.y + 8 * k;
    float a = 0.f;
    if (m < M && n < N)
      for (int s = 0; s < S; ++s) a += ws[((size_t)s * M + m) * ldws + n];
    if (!transpose) { if (m < M && n < N) out[(size_t)m * ldo + n] = a; }
    else tile[threadIdx.y + 8 * k][threadIdx.x] = a;
  }
  if (transpose) {
    __syncthreads();
    const int mo = blockIdx.y * 32 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int no = blockIdx.x * 32 + threadIdx.y + 8 * k;
      if (mo < M && no < N) out[(size_t)no * ldo + mo] = tile[threadIdx.x][threadIdx.y + 8 * k];
    }
  }
}

extern "C" size_t molclr_gemm_dw_workspace_bytes(int64_t R, int64_t O, int64_t I) {
  DwPlan pl;
  if (dw_plan(R, O, I, &pl)) return 0;
  return (size_t)pl.splits * (size_t)pl.M * (size_t)((pl.N + 3) / 4 * 4) * sizeof(float);
}

extern "C" int molclr_gemm_dw_ordered(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW,
                                      int64_t ldw, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DwPlan pl;
  if (int rc = dw_plan(R, O, I, &pl)) return rc;
  const long long ldws = (pl.N + 3) / 4 * 4;
  MOLCLR_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)pl.splits * pl.M * ldws * sizeof(float),
                 "gemm_dw_ordered: workspace too small (%zu bytes, need molclr_gemm_dw_workspace_bytes)", workspace_bytes);
  GemmJob j;
  dw_job(j, pl, dY, ldy, X, ldx, R);
  j.p.out = reinterpret_cast<float*>(workspace); j.p.ldo = ldws;
  j.ordered_ws = reinterpret_cast<float*>(workspace);
  if (j.split_k < 2) j.split_k = 2;           // (selects the split-K kernel kind; one effective split is fine)
  const int S = effective_splits((int)((R + GEMM_BK - 1) / GEMM_BK), j.split_k);
  MOLCLR_REQUIRE(S == pl.splits, "gemm_dw_ordered: internal: split count %d != planned %d", S, pl.splits);
  if (int rc = gemm_run(j, stream)) return rc;
  MOLCLR_LAUNCH(dw_reduce_kernel, dim3((unsigned)((pl.N + 31) / 32), (unsigned)((pl.M + 31) / 32), 1), dim3(32, 8, 1), 0, stream,
                reinterpret_cast<const float*>(workspace), S, (int)pl.M, (int)pl.N, ldws, dW, ldw, pl.swap);
  MOLCLR_CHECK_LAUNCH("gemm_dw_ordered reduce");
  return 0;
}

extern "C" int molclr_gemm_tf32(const molclr_gemm_args* args, cudaStream_t stream) {
  const molclr_gemm_args& a = *args;
  MOLCLR_REQUIRE(a.M < (1ll << 31) && a.N < (1ll << 31) && a.K < (1ll << 31), "gemm: extent exceeds int32");
  GemmJob j;
  memset(&j, 0, sizeof(j));
  j.A = a.A; j.lda = a.lda; j.B = a.B; j.ldb = a.ldb; j.split_k = a.split_k;
  j.A_lo = a.A_lo; j.B_lo = a.B_lo;
  j.compensate = a.compensate; j.status = a.status;
  j.B16 = a.B16; j.ld16 = a.ld16; j.rows16 = (int)a.rows16;
  GemmParams& p = j.p;
  p.M = (int)a.M; p.N = (int)a.N; p.K = (int)a.K; p.a_mn = a.a_mn; p.b_mn = a.b_mn;
  p.out = a.out; p.ldo = a.ldo; p.transpose_out = a.transpose_out; p.out2 = a.out2; p.ldo2 = a.ldo2;
  p.out_lo = a.out_lo; p.ldo_lo = a.ldo_lo;
  p.bias = a.bias; p.addend = a.addend; p.ldadd = a.ldadd; p.mask = a.mask; p.ldmask = a.ldmask;
  p.relu = a.relu; p.round_out = a.round_out; p.colstat = a.colstat; p.colstat_mode = a.colstat_mode;
  p.bits_out = a.relu_bits; p.bits_in = a.mask_bits; p.ld_bits = a.ld_bits;
  p.alpha = 1.f; p.epi = EPI_GENERIC;
  return gemm_run(j, stream);
}
