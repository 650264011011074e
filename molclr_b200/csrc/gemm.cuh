// Internal GEMM job description shared by gemm.cu and ntxent.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace molclr {

enum GemmEpilogue : int { EPI_GENERIC = 0, EPI_NTX_FWD = 1, EPI_NTX_W = 2 };

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int a_mn3d, b_mn3d;          // MN-major operand fetched with ONE 3-D TMA box per group of [k][32 mn] blocks (needs a row pitch >= the extent rounded
                               // up to 32: the last block reads the row's own padding); the 3-D map travels in the tmA2 / tmB2 slot
  int half16;                  // operands are fp16 (K-major only, 64 elements per 128-byte tile row): tcgen05 kind::f16, fp32 accumulation
  int num_kb, kb_per_split;
  int n_tiles, m_tiles, splits;   // tile grid walked by the persistent CTAs (filled by the launcher)
  int derive_lo;               // compensated product with low halves derived on chip: 1 = A_lo from an unrounded A; 2 = "mixed": A and B
                               // both unrounded, bf16 correction tiles of both formed in shared memory (see gemm.cu); 3 = fp16
                               // three-product form (K_PLAIN_H3): A split into two fp16 tiles on chip, B pre-split (b_presplit)
  int* status;                 // derive_lo 3: sticky status word (MOLCLR_STATUS_FP16_RANGE), may be null
  int segments;                // 1: plain TF32.  3: error-compensated  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (~fp32 accuracy)
  int b_presplit;              // mixed: the bf16 correction tiles of B arrive by TMA (tmB2 over [2][rows16][K] bf16); the converters touch A only
  int rows16;                  // rows per half of that tensor
  float* out; long long ldo; int transpose_out;
  float* out2; long long ldo2;
  float* out_lo; long long ldo_lo;   // tf32-rounded residual  v - round_tf32(v)  (compensated-precision consumers)
  void* out16; long long ldo16;      // EPI_NTX_W only: the softmax-weight tile as fp16, scaled by 2^10 (ldo16 in halves, % 8 == 0)
  const float* bias;
  const float* addend; long long ldadd;
  const float* mask; long long ldmask;
  uint32_t* bits_out; const uint32_t* bits_in; long long ld_bits;   // ReLU bit masks [M][ld_bits] (bit c%32 of word c/32)
  int relu, round_out;
  float* colstat; int colstat_mode;
  int stat_groups;             // number of 32-row groups the colstat buffer holds (filled by the launcher)
  int tma_store;               // plain epilogue, one output: staged 32 x 32 chunks leave through TMA stores (tensor map tmO) instead of ld.shared + st.global
  int atomic_out;
  int pf;                      // > 0: the TMA producer prefetches the K-major A tile `pf` k-blocks ahead of its loads into L2 (debug-switch builds
                               // only: measured slower, see gemm_run)
  int debug;                   // MOLCLR_GEMM_DEBUG bit 0: skip the generic epilogue (timing experiments only)
  float alpha;                 // out = alpha * acc (+ bias + addend ...)
  // ---- NT-Xent epilogues (nt_xent.py:47-65); logits l = acc * inv_tau
  int epi;
  float inv_tau;
  float ntx_bound2;            // > 0: every |logit| * log2(e) <= ntx_bound2 (unit-norm rows, moderate 1/tau): exponentials are taken relative to this
                               // a-priori bound instead of a running maximum (forward), and factored ex2(t) * (ex2(-lse_r) + ex2(-lse_k)) (weights)
  long long row_offset;        // global candidate index of A row r is r + row_offset (its own column: masked out) ...
  long long row_split, row_offset2;   // ... for r < row_split, and r - row_split + row_offset2 for the remaining rows
  long long col_offset;        // global candidate index of B row n is n + col_offset
  long long num_cand;          // Rc; positive of global row g is (g + Rc/2) mod Rc
  const float* row_lse;        // [M]   (EPI_NTX_W)
  const float* col_lse;        // [num_cand] indexed by global candidate index (EPI_NTX_W)
  float* part_max; float* part_sum;   // [n_tiles][M] (EPI_NTX_FWD)
  float* row_pos;              // [M] logit of the positive (EPI_NTX_FWD)
};

struct GemmJob {
  const float* A; long long lda; const float* B; long long ldb;   // p.half16: __half tensors, leading dimensions in halves (% 8 == 0)
  const float* A_lo; const float* B_lo;   // both non-null selects the 3-segment compensated product (same layout/ld as A, B)
  int compensate;              // 1: A and B are unrounded K-major fp32; ~fp32-accurate product, everything derived on chip ("mixed");
                               // 2: the fp16 three-product form: A unrounded fp32, B16 = fp16 tiles of B (mandatory), B unused
  int* status;                 // compensate 2: see GemmParams::status
  const void* B16; long long ld16; int rows16;   // compensate: optional pre-split bf16 correction tiles of B (molclr_prepare_weights)
  int accumulate;              // split-K: add into `out` as it is (the caller zero-filled it, or wants out += product) instead of zero-filling it first
  float* ordered_ws;           // split-K: write the per-split partial products here ([splits][M][ldws] fp32) with plain stores and sum them
                               // in split order afterwards (bit-reproducible) instead of accumulating atomically
  int split_k;
  int wide;                    // split-K only: 256 x 320 tiles on CTA pairs (both operands MN-major)
  int bn_hint;                 // 0, or a column-tile width to use instead of the default (more tiles for narrow outputs)
  GemmParams p;                // M,N,K,a_mn,b_mn and the epilogue fields filled by the caller
};

// Validates, builds the tensor maps and launches.  Returns 0 or a negative error code.
int gemm_run(const GemmJob& job, cudaStream_t stream);
// Tensor map over a row-major fp16 matrix [outer][inner] (ld in halves, % 8 == 0): boxes of box_outer rows x 64 halves (128-byte
// rows, 128-byte swizzle, zero fill out of bounds)
int gemm_make_tmap_f16(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_outer);
bool gemm_f16_ok();               // false only under the scalar debug implementation (MOLCLR_GEMM_IMPL=simt)
int gemm_n_tiles(long long N);   // number of column tiles the launcher will use for this N

}  // namespace molclr
