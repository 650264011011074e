// Small-table work of the hot path that must be EXACT fp32 and bit-reproducible:
//   * molclr_prepare_weights: the tensor-core operand forms of every Linear / GCNConv weight, one launch per forward;
//   * molclr_edge_table_grad / molclr_embed_nodes_bwd: the gradients of the bond / atom embedding tables
//     (embedding_dense_backward in the reference; ginet_molclr.py:33-39,103).  They are heavy-cancellation sums over all
//     nodes into 8 / 122 rows, so they run as plain fp32 FMAs with a fixed summation order (contiguous node blocks summed
//     sequentially by one CTA each, block partials summed in block order) instead of a split-K tensor-core contraction.
#include <cuda_bf16.h>
#include <cstring>

#include "common.cuh"
#include "molclr_b200.h"

namespace molclr {

// ------------------------------------------------------------------------------------------------ weight shadows
constexpr int kWeightBatch = 24;
struct WeightBatch { molclr_weight_desc d[kWeightBatch]; };

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(256) prepare_weights_kernel(const __grid_constant__ WeightBatch wb) {
  const molclr_weight_desc& d = wb.d[blockIdx.y];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (d.hi || d.lo) {
    const long long n = (long long)d.rows * d.ld_hi;
    for (long long i = tid; i < n; i += nth) {
      const int r = (int)(i / d.ld_hi), c = (int)(i % d.ld_hi);
      const float v = c < d.cols ? __ldg(d.src + (size_t)r * d.ld_src + c) : 0.f;
      const float h = round_tf32(v);
      if (d.hi) d.hi[i] = h;
      if (d.lo) d.lo[i] = round_tf32(v - h);
    }
  }
  const int rows_t = d.transpose_raw ? d.cols : d.rows, cols_t = d.transpose_raw ? d.rows : d.cols;
  auto at = [&](int r, int c) -> float {           // element (r, c) of the raw orientation, 0 outside
    if (r >= rows_t || c >= cols_t) return 0.f;
    return d.transpose_raw ? __ldg(d.src + (size_t)c * d.ld_src + r) : __ldg(d.src + (size_t)r * d.ld_src + c);
  };
  if (d.raw) {
    const long long n = (long long)rows_t * d.ld_raw;
    for (long long i = tid; i < n; i += nth) d.raw[i] = at((int)(i / d.ld_raw), (int)(i % d.ld_raw));
  }
  if (d.b16) {
    __nv_bfloat16* hi16 = reinterpret_cast<__nv_bfloat16*>(d.b16);
    __nv_bfloat16* lo16 = hi16 + (size_t)d.rows16 * d.ld16;
    const long long n = (long long)d.rows16 * d.ld16;
    for (long long i = tid; i < n; i += nth) {
      const float v = at((int)(i / d.ld16), (int)(i % d.ld16));
      hi16[i] = __float2bfloat16_rn(v);                     // operand of (A - trunc A) * B
      lo16[i] = __float2bfloat16_rn(v - trunc_tf32(v));     // B - trunc B: what the TF32 pass on the raw tile leaves out
    }
  }
}

// ------------------------------------------------------------------------------------------------ table gradients
constexpr int kTableUnroll = 8;

// partials[blockIdx.x][c][col] = sum over this block's rows n (increasing) of w[n][c] * g[n][col], c < 8.  Thread = column.
__global__ void __launch_bounds__(1024) class_weighted_colsum_kernel(const float* __restrict__ g, long long ld_g, const float* __restrict__ w,
                                                                     int N, int D, int rows_per_block, float* __restrict__ partials) {
  const int col = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  if (col < D) {
    for (int n = r0; n < r1; n += kTableUnroll) {
      float gv[kTableUnroll];
#pragma unroll
      for (int u = 0; u < kTableUnroll; ++u) gv[u] = (n + u < r1) ? __ldg(g + (size_t)(n + u) * ld_g + col) : 0.f;
#pragma unroll
      for (int u = 0; u < kTableUnroll; ++u) {
        if (n + u >= r1) break;
        const float4 w0 = ldg_f4(w + (size_t)(n + u) * 8), w1 = ldg_f4(w + (size_t)(n + u) * 8 + 4);     // same address in every thread: broadcast
        acc[0] = fmaf(w0.x, gv[u], acc[0]); acc[1] = fmaf(w0.y, gv[u], acc[1]); acc[2] = fmaf(w0.z, gv[u], acc[2]); acc[3] = fmaf(w0.w, gv[u], acc[3]);
        acc[4] = fmaf(w1.x, gv[u], acc[4]); acc[5] = fmaf(w1.y, gv[u], acc[5]); acc[6] = fmaf(w1.z, gv[u], acc[6]); acc[7] = fmaf(w1.w, gv[u], acc[7]);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) partials[((size_t)blockIdx.x * 8 + c) * D + col] = acc[c];
  }
}

// partials[blockIdx.x][t][col0 + col] for t < 122: rows 0..118 by atom type (key & 0xff), 119..121 by chirality (key >> 8).
// The atom-type rows live in a [119][CW] shared-memory tile (thread = column: its read-modify-write chain over the block's rows
// is sequential, hence ordered); the three chirality rows in registers.
constexpr int kOnehotCW = 160;      // columns per CTA: 119 x 160 x 4 B = 76 KB -> two CTAs per SM
__global__ void __launch_bounds__(kOnehotCW) onehot_colsum_kernel(const float* __restrict__ g, long long ld_g, const int32_t* __restrict__ key,
                                                                  int N, int D, int rows_per_block, float* __restrict__ partials) {
  extern __shared__ float tile[];          // [119][kOnehotCW]
  const int tx = threadIdx.x, col = blockIdx.y * kOnehotCW + tx;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  for (int t = 0; t < kNumAtomType; ++t) tile[t * kOnehotCW + tx] = 0.f;
  float ch0 = 0.f, ch1 = 0.f, ch2 = 0.f;
  if (col < D) {
    constexpr int U = 16;
    for (int n = r0; n < r1; n += U) {
      float gv[U];
      int kv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = n + u < r1;
        gv[u] = ok ? __ldg(g + (size_t)(n + u) * ld_g + col) : 0.f;
        kv[u] = ok ? __ldg(key + n + u) : 0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (n + u >= r1) break;
        const int a = kv[u] & 0xff, c = kv[u] >> 8;
        if (a < kNumAtomType) tile[a * kOnehotCW + tx] += gv[u];
        ch0 += c == 0 ? gv[u] : 0.f; ch1 += c == 1 ? gv[u] : 0.f; ch2 += c == 2 ? gv[u] : 0.f;
      }
    }
    float* out = partials + (size_t)blockIdx.x * (kNumAtomType + kNumChirality) * D + col;
    for (int t = 0; t < kNumAtomType; ++t) out[(size_t)t * D] = tile[t * kOnehotCW + tx];
    out[(size_t)kNumAtomType * D] = ch0; out[(size_t)(kNumAtomType + 1) * D] = ch1; out[(size_t)(kNumAtomType + 2) * D] = ch2;
  }
}

static int table_blocks() { return 4 * sm_count(); }
static int onehot_blocks() { return sm_count(); }

}  // namespace molclr

using namespace molclr;

extern "C" int molclr_prepare_weights(const molclr_weight_desc* descs, int n, cudaStream_t stream) {
  MOLCLR_REQUIRE(n >= 0 && (n == 0 || descs != nullptr), "prepare_weights: bad arguments");
  for (int i = 0; i < n; ++i) {
    const molclr_weight_desc& d = descs[i];
    MOLCLR_REQUIRE(d.src && d.rows > 0 && d.cols > 0 && d.ld_src >= d.cols, "prepare_weights: descriptor %d: bad source", i);
    MOLCLR_REQUIRE((!d.hi && !d.lo) || d.ld_hi >= d.cols, "prepare_weights: descriptor %d: ld_hi < cols", i);
    MOLCLR_REQUIRE(!d.raw || d.ld_raw >= (d.transpose_raw ? d.rows : d.cols), "prepare_weights: descriptor %d: ld_raw too small", i);
    MOLCLR_REQUIRE(!d.b16 || (d.ld16 % 8 == 0 && d.ld16 >= (d.transpose_raw ? d.rows : d.cols) && d.rows16 >= (d.transpose_raw ? d.cols : d.rows)),
                   "prepare_weights: descriptor %d: bf16 tile extents (ld16 %% 8 == 0, ld16 >= K, rows16 >= N)", i);
  }
  for (int i0 = 0; i0 < n; i0 += kWeightBatch) {
    WeightBatch wb;
    memset(&wb, 0, sizeof(wb));
    const int m = n - i0 < kWeightBatch ? n - i0 : kWeightBatch;
    long long biggest = 0;
    for (int i = 0; i < m; ++i) {
      wb.d[i] = descs[i0 + i];
      const long long e = (long long)wb.d[i].rows * wb.d[i].cols;
      if (e > biggest) biggest = e;
    }
    int bx = (int)((biggest + 256 * 4 - 1) / (256 * 4));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    prepare_weights_kernel<<<dim3((unsigned)bx, (unsigned)m, 1), 256, 0, stream>>>(wb);
    MOLCLR_CHECK_LAUNCH("prepare_weights");
  }
  return 0;
}

extern "C" size_t molclr_edge_table_grad_workspace_bytes(int D) { return (size_t)table_blocks() * 8 * D * sizeof(float); }

extern "C" int molclr_edge_table_grad(const float* ga, int64_t ld_ga, const float* cnt, int64_t N, int D, float* dB, void* workspace,
                                      cudaStream_t stream) {
  MOLCLR_REQUIRE(D > 0 && D <= 1024, "edge_table_grad: D=%d must be in 1..1024", D);
  MOLCLR_REQUIRE(N > 0 && N < (1ll << 31), "edge_table_grad: N out of range");
  MOLCLR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "edge_table_grad: workspace must be 16-byte aligned");
  const int P = table_blocks();
  const int rpb = (int)((N + P - 1) / P);
  const int used = (int)((N + rpb - 1) / rpb);
  float* partials = reinterpret_cast<float*>(workspace);
  class_weighted_colsum_kernel<<<used, (D + 31) / 32 * 32, 0, stream>>>(ga, ld_ga, cnt, (int)N, D, rpb, partials);
  MOLCLR_CHECK_LAUNCH("edge_table_grad");
  return molclr_reduce_partials(partials, used, 8 * D, 1.f, 0, dB, stream);
}

extern "C" size_t molclr_embed_nodes_bwd_workspace_bytes(int64_t N) {
  (void)N;
  return (size_t)onehot_blocks() * (kNumAtomType + kNumChirality) * 1024 * sizeof(float);      // D <= 1024
}

extern "C" int molclr_embed_nodes_bwd(const int32_t* xpacked, const float* g, int64_t ld_g, int64_t N, int D, float* dE,
                                      void* workspace, cudaStream_t stream) {
  MOLCLR_REQUIRE(D > 0 && D <= 1024, "embed_nodes_bwd: D=%d must be in 1..1024", D);
  MOLCLR_REQUIRE(N > 0 && N < (1ll << 31), "embed_nodes_bwd: N out of range");
  MOLCLR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "embed_nodes_bwd: workspace must be 16-byte aligned");
  const int P = onehot_blocks();
  const int rpb = (int)((N + P - 1) / P);
  const int used = (int)((N + rpb - 1) / rpb);
  const size_t smem = (size_t)kNumAtomType * kOnehotCW * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(onehot_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "embed_nodes_bwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  float* partials = reinterpret_cast<float*>(workspace);
  onehot_colsum_kernel<<<dim3((unsigned)used, (unsigned)((D + kOnehotCW - 1) / kOnehotCW), 1), kOnehotCW, smem, stream>>>(
      g, ld_g, xpacked, (int)N, D, rpb, partials);
  MOLCLR_CHECK_LAUNCH("embed_nodes_bwd");
  return molclr_reduce_partials(partials, used, (kNumAtomType + kNumChirality) * D, 1.f, 0, dE, stream);
}
