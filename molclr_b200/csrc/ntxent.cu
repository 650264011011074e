// NT-Xent (utils/nt_xent.py:47-65) without the 2N x 2N similarity matrix.
//
// forward : one tensor-core pass S = rep . cols^T whose epilogue keeps, per row and column tile, the
//           running (max, sum exp) of the logits with the row's own column masked, and picks the
//           positive logit; a merge kernel turns the partials into row log-sum-exp and the loss.
// backward: g_rep = (1/(tau*Rc)) * W . cols,  W[r][k] = P[r][k] + P[k][r] - 2 [k = pos(r)]  (the loss is
//           symmetric, so the column-softmax term is recomputed from the same S tile with the other
//           row's log-sum-exp).  W is produced and consumed in L2-resident column stripes of
//           kStripe candidates: a stripe GEMM with the W epilogue followed by a stripe GEMM into that stripe's partial
//           gradient; the partials are summed in stripe order at the end (deterministic).
//
// unit_rows != 0 (cosine similarity: every operand row has norm <= 1): the three tensor-core passes run on FP16 copies of the
// rows (kind::f16, fp32 accumulation) -- the same 11-bit significand as TF32, so results are those of the TF32 passes, at
// twice the tensor rate and half the shared-memory bytes -- and W is staged as fp16 scaled by 2^10 (stripes of 4096
// candidates, the same 64 MB).  Otherwise
// (dot-product similarity on rows of unknown magnitude) everything stays TF32 / fp32.
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"
#include "molclr_b200.h"
#include "ntxent.cuh"

namespace molclr {

constexpr int kStripe = 2048;     // W stripe [R][2048] fp32: 64 MB at R = 8192, L2-resident between the two GEMMs
constexpr int kStripe16Max = 4096;   // fp16 W stripe [R][4096]: the same bytes

static int stripe16() {              // MOLCLR_NTX_STRIPE16 = 1024 | 2048 | 4096 (tuning)
  static int v = 0;
  if (!v) {
    const char* e = debug_env("MOLCLR_NTX_STRIPE16");
    const int x = e ? atoi(e) : 0;
    v = (x == 1024 || x == 2048 || x == 4096) ? x : kStripe16Max;
  }
  return v;
}

// dst[r][0..C) = fp16(src[r][0..C)), row pitch ld16 halves (a multiple of 8); one thread per 4 elements
__global__ void __launch_bounds__(256) ntx_to_half_kernel(const float* __restrict__ src, long long rows, int C, int ld16, __half* __restrict__ dst) {
  pdl_sync();
  const int c4 = C / 4;
  const long long total = rows * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    const float4 v = ld_stream_f4(src + r * C + c);
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + r * ld16 + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

// dstT[c][r] = fp16(src[r][c]), row pitch ldT halves: the K-major "B" operand of the dZ contraction over candidates
__global__ void __launch_bounds__(256) ntx_to_half_t_kernel(const float* __restrict__ src, long long rows, int C, long long ldT, __half* __restrict__ dstT) {
  pdl_sync();
  __shared__ float tile[64][33];
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 64; k += 8) {
    const long long r = r0 + k;
    const int c = c0 + threadIdx.x;
    tile[k][threadIdx.x] = (r < rows && c < C) ? src[r * C + c] : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int c = c0 + k;
    const long long r = r0 + 2 * threadIdx.x;
    if (c < C && r < rows) {       // rows is even: r + 1 < rows too
      const __half2 h = __floats2half2_rn(tile[2 * threadIdx.x][k], tile[2 * threadIdx.x + 1][k]);
      *reinterpret_cast<__half2*>(dstT + (size_t)c * ldT + r) = h;
    }
  }
}

// One block handles 32 rows: threadIdx.x = row (coalesced partial reads), threadIdx.y strides over the column-tile partials;
// the 8 y-partials of a row are merged through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) ntx_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, int tiles, int R,
                                                        float* __restrict__ row_lse) {
  pdl_sync();
  __shared__ float s_m[8][33], s_s[8][33];
  const int r = blockIdx.x * 32 + threadIdx.x;
  float m = -INFINITY, s = 0.f;
  if (r < R)
    for (int t = threadIdx.y; t < tiles; t += 8) {
      const float pm = part_max[(size_t)t * R + r];
      if (pm > -INFINITY) {
        const float nm = fmaxf(m, pm);
        s = s * __expf(m - nm) + part_sum[(size_t)t * R + r] * __expf(pm - nm);
        m = nm;
      }
    }
  s_m[threadIdx.y][threadIdx.x] = m; s_s[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && r < R) {
    for (int k = 1; k < 8; ++k) {
      const float pm = s_m[k][threadIdx.x];
      if (pm > -INFINITY) {
        const float nm = fmaxf(m, pm);
        s = s * __expf(m - nm) + s_s[k][threadIdx.x] * __expf(pm - nm);
        m = nm;
      }
    }
    row_lse[r] = m + logf(s);
  }
}

// loss = (1/Rc) * sum_r (row_lse[r] - row_pos[r]); single block, fixed order -> deterministic.
__global__ void __launch_bounds__(1024) ntx_loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ row_pos, int R,
                                                        float inv_rc, float* __restrict__ loss) {
  pdl_sync();
  __shared__ double red[1024];
  double s = 0.0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) s += (double)row_lse[r] - (double)row_pos[r];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(red[0] * (double)inv_rc);
}

// out[i] = sum_s partials[s][i] in stripe order (float4, coalesced)
__global__ void __launch_bounds__(256) ntx_sum_partials_kernel(const float* __restrict__ partials, int ns, long long len4,
                                                               float* __restrict__ out) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = ld_stream_f4(partials + 4 * i);
    for (int s = 1; s < ns; ++s) a = f4_add(a, ld_stream_f4(partials + (size_t)s * len4 * 4 + 4 * i));
    st_f4(out + 4 * i, a);
  }
}

}  // namespace molclr

using namespace molclr;

// unit-norm rows: |logit| <= 1/tau (with a margin for rounding).  For moderate 1/tau the kernels take exponentials relative to
// this bound (no running maximum, one EX2 per weight); 0 = fall back to the general forms (tau < ~0.045: 2^(-2 bound) would
// leave the fp32 range comfortably representable sums)
static float ntx_bound2(float inv_temperature) {
  const float b = 1.4426950408889634f * inv_temperature * 1.001f;
  return (b > 0.f && b <= 32.f && !debug_env("MOLCLR_NTX_NOBOUND")) ? b : 0.f;
}

static bool ntx_fused_enabled() {      // MOLCLR_NTX_FUSED=0: the striped two-GEMM backward (A/B timing, and the path of C > 256)
  static int v = -1;
  if (v < 0) { const char* e = debug_env("MOLCLR_NTX_FUSED"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

static int64_t num_stripes(int64_t Rc, int w = kStripe) { return (Rc + w - 1) / w; }
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Workspace layout (one buffer serves either call and either operand mode)
struct NtxLayout {
  int ld16; long long ldT;
  size_t part, f_rep16, f_cols16, fwd_total;                          // forward
  size_t stripe, partials, b_rep16, b_cols16, b_colsT16, b_ecol, bwd_total;   // backward
};
static NtxLayout ntx_layout(int64_t R, int64_t Rc, int C) {
  NtxLayout l;
  l.ld16 = (C + 7) & ~7; l.ldT = (Rc + 7) & ~(long long)7;
  const size_t rep16 = align256((size_t)R * l.ld16 * 2), cols16 = align256((size_t)Rc * l.ld16 * 2), colsT16 = align256((size_t)C * l.ldT * 2);
  l.part = 0;
  l.f_rep16 = align256((size_t)2 * gemm_n_tiles(Rc) * R * sizeof(float));
  l.f_cols16 = l.f_rep16 + rep16;
  l.fwd_total = l.f_cols16 + cols16;
  l.stripe = 0;                                                       // [R][2048] fp32 or [R][4096] fp16
  l.partials = align256((size_t)R * kStripe * sizeof(float));
  // one partial-gradient slot per stripe of the NARROWEST stripe width either path may use (fp32: kStripe; fp16: stripe16())
  const int64_t nstr = num_stripes(Rc, stripe16() < kStripe ? stripe16() : kStripe);
  const int64_t slots = nstr > kNtxFusedMaxSplits ? nstr : kNtxFusedMaxSplits;
  l.b_rep16 = l.partials + align256((size_t)slots * R * C * sizeof(float));
  l.b_cols16 = l.b_rep16 + rep16;
  l.b_colsT16 = l.b_cols16 + cols16;
  l.b_ecol = l.b_colsT16 + colsT16;
  l.bwd_total = l.b_ecol + align256(ntx_fused_ecol_floats(Rc) * sizeof(float));
  return l;
}

extern "C" size_t molclr_ntxent_workspace_bytes(int64_t R, int64_t Rc, int C) {
  const NtxLayout l = ntx_layout(R, Rc, C);
  return (l.fwd_total > l.bwd_total ? l.fwd_total : l.bwd_total) + 256;
}

static int to_half(const float* src, int64_t rows, int C, int ld16, __half* dst, cudaStream_t stream) {
  const long long total = rows * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
  MOLCLR_LAUNCH(ntx_to_half_kernel, (int)blocks, 256, 0, stream, src, rows, C, ld16, dst);
  MOLCLR_CHECK_LAUNCH("ntx_to_half");
  return 0;
}

// fp16 copies of rep and cols; rep rows that ARE rows of cols (one GPU: rep == cols) are not converted twice
static int ntx_operands16(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int ld16, __half* rep16_buf, __half* cols16,
                          const __half** rep16, cudaStream_t stream) {
  int rc = to_half(cols, Rc, C, ld16, cols16, stream);
  if (rc) return rc;
  if (rep >= cols && rep + (size_t)R * C <= cols + (size_t)Rc * C && (rep - cols) % C == 0) {
    *rep16 = cols16 + (size_t)((rep - cols) / C) * ld16;
    return 0;
  }
  *rep16 = rep16_buf;
  return to_half(rep, R, C, ld16, rep16_buf, stream);
}

// rep16 / cols16 (optional, both or neither): fp16 copies of the operands supplied by the caller (row pitch ld16_in halves); then rep /
// cols may be NULL and nothing is converted here.
static int ntxent_fwd_impl(const float* rep, const float* cols, const __half* rep16_in, const __half* cols16_in, int64_t ld16_in, int64_t R,
                           int64_t Rc, int C, int64_t row_offset, int64_t row_offset2, float inv_temperature, int unit_rows, float* row_lse,
                           float* row_pos, float* loss, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MOLCLR_REQUIRE(R > 0 && R % 2 == 0 && Rc >= R && Rc % 4 == 0 && C % 4 == 0, "ntxent: need R > 0 and even, Rc >= R, Rc %% 4 == 0, C %% 4 == 0 (R=%lld Rc=%lld C=%d)",
                 (long long)R, (long long)Rc, C);
  MOLCLR_REQUIRE(workspace_bytes >= molclr_ntxent_workspace_bytes(R, Rc, C), "ntxent_fwd: workspace too small");
  const NtxLayout l = ntx_layout(R, Rc, C);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int tiles = gemm_n_tiles(Rc);
  float* part_max = reinterpret_cast<float*>(ws + l.part);
  float* part_sum = part_max + (size_t)tiles * R;
  const bool f16 = unit_rows != 0 && gemm_f16_ok();
  GemmJob j;
  memset(&j, 0, sizeof(j));
  j.A = rep; j.lda = C; j.B = cols; j.ldb = C; j.split_k = 1;
  if (f16) {
    const __half* rep16 = rep16_in;
    const __half* cols16 = cols16_in;
    int64_t ld16 = ld16_in;
    if (!cols16_in) {
      int rc = ntx_operands16(rep, cols, R, Rc, C, l.ld16, reinterpret_cast<__half*>(ws + l.f_rep16), reinterpret_cast<__half*>(ws + l.f_cols16), &rep16, stream);
      if (rc) return rc;
      cols16 = reinterpret_cast<const __half*>(ws + l.f_cols16); ld16 = l.ld16;
    }
    j.A = reinterpret_cast<const float*>(rep16); j.lda = ld16;
    j.B = reinterpret_cast<const float*>(cols16); j.ldb = ld16;
    j.p.half16 = 1; j.p.ntx_bound2 = ntx_bound2(inv_temperature);
  }
  GemmParams& p = j.p;
  p.M = (int)R; p.N = (int)Rc; p.K = C; p.alpha = 1.f;
  p.epi = EPI_NTX_FWD; p.inv_tau = inv_temperature; p.row_offset = row_offset; p.row_split = R / 2; p.row_offset2 = row_offset2; p.col_offset = 0; p.num_cand = Rc;
  p.part_max = part_max; p.part_sum = part_sum; p.row_pos = row_pos;
  int rc = gemm_run(j, stream);
  if (rc) return rc;
  MOLCLR_LAUNCH(ntx_merge_kernel, (int)((R + 31) / 32), dim3(32, 8), 0, stream, part_max, part_sum, tiles, (int)R, row_lse);
  MOLCLR_CHECK_LAUNCH("ntx_merge");
  if (loss) {
    MOLCLR_LAUNCH(ntx_loss_kernel, 1, 1024, 0, stream, row_lse, row_pos, (int)R, 1.0f / (float)Rc, loss);
    MOLCLR_CHECK_LAUNCH("ntx_loss");
  }
  return 0;
}

extern "C" int molclr_ntxent_fwd(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int64_t row_offset,
                                 int64_t row_offset2, float inv_temperature, int unit_rows, float* row_lse, float* row_pos, float* loss,
                                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return ntxent_fwd_impl(rep, cols, nullptr, nullptr, 0, R, Rc, C, row_offset, row_offset2, inv_temperature, unit_rows, row_lse, row_pos, loss,
                         workspace, workspace_bytes, stream);
}

// fp16 operands supplied by the caller (unit-norm rows): what the data-parallel path all-gathers, at half the bytes of fp32
extern "C" int molclr_ntxent_h_supported(int C, float inv_temperature) {
  return (gemm_f16_ok() && ntx_bound2(inv_temperature) > 0.f && C <= 256 && C % 8 == 0 && ntx_fused_enabled()) ? 1 : 0;
}

extern "C" int molclr_ntxent_fwd_h(const void* rep16, const void* cols16, int64_t ld16, int64_t R, int64_t Rc, int C, int64_t row_offset,
                                   int64_t row_offset2, float inv_temperature, float* row_lse, float* row_pos, float* loss, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
  MOLCLR_REQUIRE(rep16 && cols16 && ld16 >= C && ld16 % 8 == 0, "ntxent_fwd_h: fp16 operands with a row pitch that is a multiple of 8 halves");
  MOLCLR_REQUIRE(molclr_ntxent_h_supported(C, inv_temperature), "ntxent_fwd_h: not supported for C=%d, 1/tau=%g (see molclr_ntxent_h_supported)", C, inv_temperature);
  return ntxent_fwd_impl(nullptr, nullptr, reinterpret_cast<const __half*>(rep16), reinterpret_cast<const __half*>(cols16), ld16, R, Rc, C, row_offset,
                         row_offset2, inv_temperature, 1, row_lse, row_pos, loss, workspace, workspace_bytes, stream);
}

extern "C" int molclr_ntxent_bwd_h(const void* rep16, const void* cols16, int64_t ld16, int64_t R, int64_t Rc, int C, int64_t row_offset,
                                   int64_t row_offset2, float inv_temperature, const float* row_lse, const float* col_lse, float gscale, float* g_rep,
                                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MOLCLR_REQUIRE(R > 0 && R % 2 == 0 && Rc >= R && Rc % 4 == 0 && C % 4 == 0, "ntxent: need R > 0 and even, Rc >= R, Rc %% 4 == 0, C %% 4 == 0");
  MOLCLR_REQUIRE(rep16 && cols16 && ld16 >= C && ld16 % 8 == 0, "ntxent_bwd_h: fp16 operands with a row pitch that is a multiple of 8 halves");
  MOLCLR_REQUIRE(molclr_ntxent_h_supported(C, inv_temperature), "ntxent_bwd_h: not supported for C=%d, 1/tau=%g (see molclr_ntxent_h_supported)", C, inv_temperature);
  MOLCLR_REQUIRE(workspace_bytes >= molclr_ntxent_workspace_bytes(R, Rc, C), "ntxent_bwd_h: workspace too small");
  const NtxLayout l = ntx_layout(R, Rc, C);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* partials = reinterpret_cast<float*>(ws + l.partials);
  const int splits = ntx_fused_splits(R, Rc);
  int rc = ntx_bwd_fused(reinterpret_cast<const __half*>(rep16), reinterpret_cast<const __half*>(cols16), (int)ld16, R, Rc, C, row_offset, row_offset2,
                         inv_temperature, ntx_bound2(inv_temperature), row_lse, col_lse, gscale, reinterpret_cast<float*>(ws + l.b_ecol), partials, splits, stream);
  if (rc) return rc;
  const long long len4 = (long long)R * C / 4;
  long long blocks = (len4 + 255) / 256;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  MOLCLR_LAUNCH(ntx_sum_partials_kernel, (int)blocks, 256, 0, stream, partials, splits, len4, g_rep);
  MOLCLR_CHECK_LAUNCH("ntx_sum_partials");
  return 0;
}

extern "C" int molclr_ntxent_bwd(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int64_t row_offset,
                                 int64_t row_offset2, float inv_temperature, int unit_rows, const float* row_lse, const float* col_lse, float gscale,
                                 float* g_rep, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MOLCLR_REQUIRE(R > 0 && R % 2 == 0 && Rc >= R && Rc % 4 == 0 && C % 4 == 0, "ntxent: need R > 0 and even, Rc >= R, Rc %% 4 == 0, C %% 4 == 0");
  MOLCLR_REQUIRE(workspace_bytes >= molclr_ntxent_workspace_bytes(R, Rc, C), "ntxent_bwd: workspace too small");
  const NtxLayout l = ntx_layout(R, Rc, C);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* stripe = reinterpret_cast<float*>(ws + l.stripe);
  float* partials = reinterpret_cast<float*>(ws + l.partials);    // [stripes][R][C]
  const bool f16 = unit_rows != 0 && gemm_f16_ok();
  const int sw = f16 ? stripe16() : kStripe;
  const int64_t ns = num_stripes(Rc, sw);
  const __half* rep16 = nullptr;
  const __half* cols16 = reinterpret_cast<const __half*>(ws + l.b_cols16);
  __half* colsT16 = reinterpret_cast<__half*>(ws + l.b_colsT16);
  const float bound2 = ntx_bound2(inv_temperature);
  if (f16) {
    int rc = ntx_operands16(rep, cols, R, Rc, C, l.ld16, reinterpret_cast<__half*>(ws + l.b_rep16), reinterpret_cast<__half*>(ws + l.b_cols16), &rep16, stream);
    if (rc) return rc;
    if (bound2 > 0.f && C <= 256 && ntx_fused_enabled()) {
      // fused: S tile -> softmax weights -> second product, all on chip (ntxent_fused.cu)
      const int splits = ntx_fused_splits(R, Rc);
      rc = ntx_bwd_fused(rep16, cols16, l.ld16, R, Rc, C, row_offset, row_offset2, inv_temperature, bound2, row_lse, col_lse, gscale,
                         reinterpret_cast<float*>(ws + l.b_ecol), partials, splits, stream);
      if (rc) return rc;
      const long long len4 = (long long)R * C / 4;
      long long blocks = (len4 + 255) / 256;
      if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
      MOLCLR_LAUNCH(ntx_sum_partials_kernel, (int)blocks, 256, 0, stream, partials, splits, len4, g_rep);
      MOLCLR_CHECK_LAUNCH("ntx_sum_partials");
      return 0;
    }
    MOLCLR_LAUNCH(ntx_to_half_t_kernel, dim3((unsigned)((Rc + 63) / 64), (unsigned)((C + 31) / 32)), dim3(32, 8), 0, stream, cols, Rc, C, l.ldT, colsT16);
    MOLCLR_CHECK_LAUNCH("ntx_to_half_t");
  }
  for (int64_t si = 0; si < ns; ++si) {
    const int64_t c0 = si * sw;
    const int kc = (int)((Rc - c0) < sw ? (Rc - c0) : sw);
    GemmJob w;
    memset(&w, 0, sizeof(w));
    w.A = rep; w.lda = C; w.B = cols + (size_t)c0 * C; w.ldb = C; w.split_k = 1;
    w.p.M = (int)R; w.p.N = kc; w.p.K = C; w.p.alpha = 1.f;
    w.p.epi = EPI_NTX_W; w.p.inv_tau = inv_temperature; w.p.row_offset = row_offset; w.p.row_split = R / 2; w.p.row_offset2 = row_offset2; w.p.col_offset = c0; w.p.num_cand = Rc;
    w.p.row_lse = row_lse; w.p.col_lse = col_lse;
    if (f16) {
      w.A = reinterpret_cast<const float*>(rep16); w.lda = l.ld16;
      w.B = reinterpret_cast<const float*>(cols16 + (size_t)c0 * l.ld16); w.ldb = l.ld16;
      w.p.half16 = 1; w.p.out16 = stripe; w.p.ldo16 = sw; w.p.ntx_bound2 = ntx_bound2(inv_temperature);
    } else {
      w.p.out = stripe; w.p.ldo = kStripe; w.p.round_out = 1;
    }
    int rc = gemm_run(w, stream);
    if (rc) return rc;
    GemmJob g;
    memset(&g, 0, sizeof(g));
    g.split_k = 1;
    g.p.M = (int)R; g.p.N = C; g.p.K = kc; g.p.epi = EPI_GENERIC;
    if (f16) {     // both operands K-major over the candidates: W stripe [R][sw] and cols^T [C][ldT]; W carries a factor 2^10
      g.A = stripe; g.lda = sw; g.B = reinterpret_cast<const float*>(colsT16 + c0); g.ldb = l.ldT;
      g.p.half16 = 1; g.p.alpha = inv_temperature * gscale * (1.f / 1024.f);
    } else {
      g.A = stripe; g.lda = kStripe; g.B = cols + (size_t)c0 * C; g.ldb = C;
      g.p.a_mn = 0; g.p.b_mn = 1; g.p.alpha = inv_temperature * gscale;
    }
    g.p.out = partials + (size_t)si * R * C; g.p.ldo = C;
    g.bn_hint = 128;             // C = 256 is one 256-wide tile per row tile: halve it so that the stripe fills the GPU
    rc = gemm_run(g, stream);
    if (rc) return rc;
  }
  const long long len4 = (long long)R * C / 4;
  long long blocks = (len4 + 255) / 256;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  MOLCLR_LAUNCH(ntx_sum_partials_kernel, (int)blocks, 256, 0, stream, partials, (int)ns, len4, g_rep);
  MOLCLR_CHECK_LAUNCH("ntx_sum_partials");
  return 0;
}
