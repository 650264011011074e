// Small-table work of the hot path that must be EXACT fp32 and bit-reproducible:
//   * molclr_prepare_weights: the tensor-core operand forms of every Linear / GCNConv weight, one launch per forward;
//   * molclr_edge_table_grad / molclr_embed_nodes_bwd: the gradients of the bond / atom embedding tables
//     (embedding_dense_backward in the reference; ginet_molclr.py:33-39,103).  They are heavy-cancellation sums over all
//     nodes into 8 / 122 rows, so they run as plain fp32 FMAs with a fixed summation order (row tiles dealt round-robin to the
//     CTAs and summed in increasing order by each, CTA partials summed in CTA order) instead of a split-K tensor-core contraction.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstring>

#include "common.cuh"
#include "molclr_b200.h"
#include "ptx.cuh"

namespace molclr {

// ------------------------------------------------------------------------------------------------ weight shadows
constexpr int kWeightBatch = 24;
struct WeightBatch { molclr_weight_desc d[kWeightBatch]; };

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(256) prepare_weights_kernel(const __grid_constant__ WeightBatch wb) {
  pdl_sync();
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
  if (d.hi_t) {
    const long long n = (long long)d.cols * d.ld_hi_t;
    for (long long i = tid; i < n; i += nth) {
      const int r = (int)(i / d.ld_hi_t), c = (int)(i % d.ld_hi_t);          // element (r, c) of W^T = W[c][r]
      d.hi_t[i] = c < d.rows ? round_tf32(__ldg(d.src + (size_t)c * d.ld_src + r)) : 0.f;
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
  if (d.b16 && d.b16_kind == 1) {
    // fp16 halves of 2^6 W (exact scaling): h = fp16(s w), l = fp16(s w - h); |s w| beyond fp16's range saturates (weights of
    // magnitude >= 1024 do not occur)
    uint16_t* hi16 = reinterpret_cast<uint16_t*>(d.b16);
    uint16_t* lo16 = hi16 + (size_t)d.rows16 * d.ld16;
    const long long n = (long long)d.rows16 * d.ld16;
    for (long long i = tid; i < n; i += nth) {
      const float v = at((int)(i / d.ld16), (int)(i % d.ld16)) * (float)MOLCLR_H16_SCALE;
      const uint32_t h = ptx::cvt_f16x2_sat(v, 0.f) & 0xFFFFu;
      const float hf = __half2float(__ushort_as_half((unsigned short)h));
      hi16[i] = (uint16_t)h;
      lo16[i] = (uint16_t)(ptx::cvt_f16x2_sat(v - hf, 0.f) & 0xFFFFu);
    }
  } else if (d.b16) {
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
constexpr int kTabRY = 4;          // row lanes per CTA (at most)
constexpr int kTabUnroll = 6;

// partials[blockIdx.x][c][:] = sum over this CTA's rows n of w[n][c] * g[n][:], c < 8.  Thread (tx, ty): float4 column chunk tx, row
// lane ty takes rows r0 + ty, r0 + ty + RY, ... in increasing order (registers); the RY lane sums are then added in lane order.
__global__ void __launch_bounds__(384, 2) class_weighted_colsum_kernel(const float* __restrict__ g, long long ld_g, const float* __restrict__ w,
                                                                     int N, int D, int rows_per_block, float* __restrict__ partials) {
  pdl_sync();
  extern __shared__ float red[];          // [RY][8][D], then the weight rows of the current tile: [RY * U][8]
  const int tx = threadIdx.x, ty = threadIdx.y, D4 = D >> 2, RY = blockDim.y;
  const int tid = ty * blockDim.x + tx, nth = blockDim.x * blockDim.y;
  // row tiles of RY * kTabUnroll rows are dealt round-robin to the CTAs, so that at any time the grid reads ONE contiguous,
  // DRAM-page-friendly window of the matrix; a CTA still adds its tiles in a fixed (increasing) order
  const int tile_rows = RY * kTabUnroll;
  (void)rows_per_block;
  float4* wt = reinterpret_cast<float4*>(red + (size_t)RY * 8 * D);       // [tile_rows][2] float4
  float4 acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = f4_zero();
  for (long long t0 = (long long)blockIdx.x * tile_rows; t0 < N; t0 += (long long)gridDim.x * tile_rows) {
    const int n = (int)t0 + ty;
    // the tile's feature rows go to registers first (their DRAM latency then overlaps the staging of the weight rows) ...
    float4 gv[kTabUnroll];
#pragma unroll
    for (int u = 0; u < kTabUnroll; ++u) {
      const int row = n + u * RY;
      gv[u] = (row < N && tx < D4) ? ldg_f4(g + (size_t)row * ld_g + 4 * tx) : f4_zero();
    }
    // ... the [tile_rows][8] weights to shared memory: one coalesced load per tile instead of a dependent L2 round trip per row
    __syncthreads();
    if (tid < 2 * tile_rows) wt[tid] = ((int)t0 + (tid >> 1) < N) ? ldg_f4(w + ((size_t)t0 + (tid >> 1)) * 8 + 4 * (tid & 1)) : f4_zero();
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kTabUnroll; ++u) {
      const float4 w0 = wt[2 * (ty + u * RY)], w1 = wt[2 * (ty + u * RY) + 1];              // broadcast
      const float wc[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        acc[c].x = fmaf(wc[c], gv[u].x, acc[c].x); acc[c].y = fmaf(wc[c], gv[u].y, acc[c].y);
        acc[c].z = fmaf(wc[c], gv[u].z, acc[c].z); acc[c].w = fmaf(wc[c], gv[u].w, acc[c].w);
      }
    }
  }
  if (tx < D4) {
#pragma unroll
    for (int c = 0; c < 8; ++c) st_f4(red + ((size_t)ty * 8 + c) * D + 4 * tx, acc[c]);
  }
  __syncthreads();
  for (int e = tid; e < 8 * D; e += nth) {
    float s = red[e];
    for (int l = 1; l < RY; ++l) s += red[(size_t)l * 8 * D + e];
    partials[(size_t)blockIdx.x * 8 * D + e] = s;
  }
}

// Atom / chirality table gradients: partials[blockIdx.x][t][cols of this CTA] for t < 122 (rows 0..118 by atom type = key & 0xff,
// 119..121 by chirality = key >> 8).  A CTA walks its nodes in chunks of 128: the chunk's rows are ordered by atom type with a stable
// in-CTA counting rank, each row lane then sums a contiguous piece of that order in REGISTERS, run by run, and adds a finished
// run to the CTA's [119][CW] shared-memory tile -- one read-modify-write per run instead of per node.  Only the first run of a
// lane can share its class with another lane (the order is sorted): it goes to a per-lane spill row that is added in lane order.
// Everything has a fixed order: bit-reproducible.
constexpr int kOhChunk = 128, kOhRY = 4, kOhUnroll = 16, kOhMaxCW = 384;
__global__ void __launch_bounds__(512, 1) onehot_colsum_kernel(const float* __restrict__ g, long long ld_g, const int32_t* __restrict__ key,
                                                             int N, int D, int CW, int chunks_per_block, float* __restrict__ partials) {
  pdl_sync();
  extern __shared__ float sm[];
  float* tile = sm;                                        // [119][CW]
  float* spill = tile + (size_t)kNumAtomType * CW;         // [RY][CW]
  float* chbuf = spill + (size_t)kOhRY * CW;               // [RY][3][CW]
  int* keys = reinterpret_cast<int*>(chbuf + (size_t)kOhRY * 3 * CW);     // [128]
  int* order = keys + kOhChunk;                            // [128] chunk-local row index at each sorted position
  int* spill_cls = order + kOhChunk;                       // [RY]
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * blockDim.x + tx, nth = blockDim.x * blockDim.y;
  const int col0 = blockIdx.y * CW, cw = min(CW, D - col0), cw4 = cw >> 2;
  const bool active = tx < cw4;
  for (int e = tid; e < kNumAtomType * CW; e += nth) tile[e] = 0.f;
  float4 ch0 = f4_zero(), ch1 = f4_zero(), ch2 = f4_zero();
  // chunks are dealt round-robin to the CTAs (one contiguous window of the matrix in flight at any time); fixed order per CTA
  const int nchunks = (N + kOhChunk - 1) / kOhChunk;
  (void)chunks_per_block;
  int next_key = (tid < kOhChunk && (int)blockIdx.x * kOhChunk + tid < N) ? __ldg(key + (size_t)blockIdx.x * kOhChunk + tid) : 0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int base = chunk * kOhChunk, cnt = min(kOhChunk, N - base);
    __syncthreads();                                       // previous chunk fully consumed (keys / order / spill reuse)
    if (tid < kOhChunk) keys[tid] = next_key;
    if (tid < kOhRY) spill_cls[tid] = -1;
    __syncthreads();
    {                                                      // the keys of this CTA's NEXT chunk: in flight while this one is summed
      const long long nb = ((long long)chunk + gridDim.x) * kOhChunk + tid;
      next_key = (tid < kOhChunk && nb < N) ? __ldg(key + nb) : 0;
    }
    if (tid < cnt) {                                       // stable rank of row tid in atom-type order
      const int a = keys[tid] & 0xff;
      int rank = 0;
      for (int j = 0; j < cnt; ++j) { const int aj = keys[j] & 0xff; rank += (aj < a || (aj == a && j < tid)) ? 1 : 0; }
      order[rank] = tid;
    }
    __syncthreads();
    const int seg = (cnt + kOhRY - 1) / kOhRY, p0 = ty * seg, p1 = min(cnt, p0 + seg);
    if (active && p0 < p1) {
      float4 acc = f4_zero();
      int cur = -1;
      bool first = true;
      auto flush = [&]() {
        if (cur < 0) return;
        if (first && ty > 0) { st_f4(spill + (size_t)ty * CW + 4 * tx, acc); if (tx == 0) spill_cls[ty] = cur; }
        else { float* t = tile + (size_t)cur * CW + 4 * tx; st_f4(t, f4_add(*reinterpret_cast<const float4*>(t), acc)); }
        first = false;
      };
      for (int p = p0; p < p1; p += kOhUnroll) {
        float4 v[kOhUnroll];
        int kk[kOhUnroll];
#pragma unroll
        for (int u = 0; u < kOhUnroll; ++u) {
          const bool ok = p + u < p1;
          const int r = ok ? order[p + u] : 0;
          kk[u] = ok ? keys[r] : -1;
          v[u] = ok ? ldg_f4(g + (size_t)(base + r) * ld_g + col0 + 4 * tx) : f4_zero();
        }
#pragma unroll
        for (int u = 0; u < kOhUnroll; ++u) {
          if (kk[u] < 0) break;
          const int a = kk[u] & 0xff, c = kk[u] >> 8;
          if (a != cur) { flush(); cur = a; acc = f4_zero(); }
          acc = f4_add(acc, v[u]);
          if (c == 0) ch0 = f4_add(ch0, v[u]); else if (c == 1) ch1 = f4_add(ch1, v[u]); else if (c == 2) ch2 = f4_add(ch2, v[u]);
        }
      }
      flush();
    }
    __syncthreads();
    if (ty == 0 && active) {                               // split runs, in lane order
      for (int l = 1; l < kOhRY; ++l) {
        const int cls = spill_cls[l];
        if (cls < 0) continue;
        float* t = tile + (size_t)cls * CW + 4 * tx;
        st_f4(t, f4_add(*reinterpret_cast<const float4*>(t), *reinterpret_cast<const float4*>(spill + (size_t)l * CW + 4 * tx)));
      }
    }
  }
  if (active) {
    st_f4(chbuf + ((size_t)ty * 3 + 0) * CW + 4 * tx, ch0);
    st_f4(chbuf + ((size_t)ty * 3 + 1) * CW + 4 * tx, ch1);
    st_f4(chbuf + ((size_t)ty * 3 + 2) * CW + 4 * tx, ch2);
  }
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * (kNumAtomType + kNumChirality) * D + col0;
  for (int e = tid; e < kNumAtomType * cw; e += nth) { const int t = e / cw, c = e - t * cw; out[(size_t)t * D + c] = tile[(size_t)t * CW + c]; }
  for (int e = tid; e < kNumChirality * cw; e += nth) {
    const int t = e / cw, c = e - t * cw;
    float s = chbuf[(size_t)t * CW + c];
#pragma unroll
    for (int l = 1; l < kOhRY; ++l) s += chbuf[((size_t)l * 3 + t) * CW + c];
    out[(size_t)(kNumAtomType + t) * D + c] = s;
  }
}

static int table_blocks() { return 2 * sm_count(); }
static int onehot_blocks() { return sm_count(); }

}  // namespace molclr

using namespace molclr;

extern "C" int molclr_prepare_weights(const molclr_weight_desc* descs, int n, cudaStream_t stream) {
  MOLCLR_REQUIRE(n >= 0 && (n == 0 || descs != nullptr), "prepare_weights: bad arguments");
  for (int i = 0; i < n; ++i) {
    const molclr_weight_desc& d = descs[i];
    MOLCLR_REQUIRE(d.src && d.rows > 0 && d.cols > 0 && d.ld_src >= d.cols, "prepare_weights: descriptor %d: bad source", i);
    MOLCLR_REQUIRE((!d.hi && !d.lo) || d.ld_hi >= d.cols, "prepare_weights: descriptor %d: ld_hi < cols", i);
    MOLCLR_REQUIRE(!d.hi_t || d.ld_hi_t >= d.rows, "prepare_weights: descriptor %d: ld_hi_t < rows", i);
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
    MOLCLR_LAUNCH(prepare_weights_kernel, dim3((unsigned)bx, (unsigned)m, 1), 256, 0, stream, wb);
    MOLCLR_CHECK_LAUNCH("prepare_weights");
  }
  return 0;
}

extern "C" size_t molclr_edge_table_grad_workspace_bytes(int D) { return (size_t)table_blocks() * 8 * D * sizeof(float); }

extern "C" int molclr_edge_table_grad(const float* ga, int64_t ld_ga, const float* cnt, int64_t N, int D, float* dB, void* workspace,
                                      cudaStream_t stream) {
  MOLCLR_REQUIRE(D > 0 && D % 4 == 0 && D <= 512 && ld_ga % 4 == 0 && (reinterpret_cast<uintptr_t>(ga) & 15) == 0,
                 "edge_table_grad: D=%d must be a multiple of 4, <= 512; rows 16-byte aligned", D);
  MOLCLR_REQUIRE(N > 0 && N < (1ll << 31), "edge_table_grad: N out of range");
  MOLCLR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "edge_table_grad: workspace must be 16-byte aligned");
  const int P = table_blocks();
  const int tx = (D / 4 + 31) / 32 * 32, ry = 384 / tx < kTabRY ? 384 / tx : kTabRY;      // <= 384 threads (two CTAs per SM)
  const int rpb = 0;
  const long long tiles = (N + ry * kTabUnroll - 1) / (ry * kTabUnroll);
  const int used = (int)(tiles < P ? tiles : P);
  float* partials = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)ry * 8 * D * sizeof(float) + (size_t)ry * kTabUnroll * 8 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(class_weighted_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTabRY * 8 * 512 * (int)sizeof(float) + kTabRY * kTabUnroll * 8 * (int)sizeof(float));
    if (e != cudaSuccess) return cuda_fail(e, "edge_table_grad: cudaFuncSetAttribute");
    attr_set = true;
  }
  MOLCLR_LAUNCH(class_weighted_colsum_kernel, used, dim3((unsigned)tx, (unsigned)ry, 1), smem, stream, ga, ld_ga, cnt, (int)N, D, rpb, partials);
  MOLCLR_CHECK_LAUNCH("edge_table_grad");
  return molclr_reduce_partials(partials, used, 8 * D, 1.f, 0, dB, stream);
}

extern "C" size_t molclr_embed_nodes_bwd_workspace_bytes(int64_t N) {
  (void)N;
  return (size_t)onehot_blocks() * (kNumAtomType + kNumChirality) * 1024 * sizeof(float);      // D <= 1024
}

extern "C" int molclr_embed_nodes_bwd(const int32_t* xpacked, const float* g, int64_t ld_g, int64_t N, int D, float* dE,
                                      void* workspace, cudaStream_t stream) {
  MOLCLR_REQUIRE(D > 0 && D % 4 == 0 && D <= 1024 && ld_g % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0,
                 "embed_nodes_bwd: D=%d must be a multiple of 4, <= 1024; rows 16-byte aligned", D);
  MOLCLR_REQUIRE(N > 0 && N < (1ll << 31), "embed_nodes_bwd: N out of range");
  MOLCLR_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "embed_nodes_bwd: workspace must be 16-byte aligned");
  const int slices = (D + kOhMaxCW - 1) / kOhMaxCW;
  const int CW = ((D + slices - 1) / slices + 3) / 4 * 4;               // columns per CTA (a multiple of 4, <= 384)
  const int nchunks = (int)((N + kOhChunk - 1) / kOhChunk);
  const int P = onehot_blocks();
  const int cpb = 0;
  const int used = nchunks < P ? nchunks : P;
  const size_t smem = ((size_t)kNumAtomType + kOhRY + 3 * kOhRY) * CW * sizeof(float) + (2 * kOhChunk + kOhRY) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    const size_t mx = ((size_t)kNumAtomType + kOhRY + 3 * kOhRY) * kOhMaxCW * sizeof(float) + (2 * kOhChunk + kOhRY) * sizeof(int);
    cudaError_t e = cudaFuncSetAttribute(onehot_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx);
    if (e != cudaSuccess) return cuda_fail(e, "embed_nodes_bwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  float* partials = reinterpret_cast<float*>(workspace);
  int tx = (CW / 4 + 31) / 32 * 32;
  if (tx * kOhRY < kOhChunk) tx = kOhChunk / kOhRY;                       // the first 128 threads load / rank the chunk's keys
  MOLCLR_LAUNCH(onehot_colsum_kernel, dim3((unsigned)used, (unsigned)slices, 1), dim3((unsigned)tx, kOhRY, 1), smem, stream,
                g, ld_g, xpacked, (int)N, D, CW, cpb, partials);
  MOLCLR_CHECK_LAUNCH("embed_nodes_bwd");
  return molclr_reduce_partials(partials, used, (kNumAtomType + kNumChirality) * D, 1.f, 0, dE, stream);
}
