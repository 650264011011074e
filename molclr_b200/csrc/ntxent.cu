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
#include "common.cuh"
#include "gemm.cuh"
#include "molclr_b200.h"

namespace molclr {

constexpr int kStripe = 2048;     // W stripe [R][2048] fp32: 64 MB at R = 8192, L2-resident between the two GEMMs

// One block handles 32 rows: threadIdx.x = row (coalesced partial reads), threadIdx.y strides over the column-tile partials;
// the 8 y-partials of a row are merged through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) ntx_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, int tiles, int R,
                                                        float* __restrict__ row_lse) {
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
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = ld_stream_f4(partials + 4 * i);
    for (int s = 1; s < ns; ++s) a = f4_add(a, ld_stream_f4(partials + (size_t)s * len4 * 4 + 4 * i));
    st_f4(out + 4 * i, a);
  }
}

}  // namespace molclr

using namespace molclr;

static int64_t num_stripes(int64_t Rc) { return (Rc + kStripe - 1) / kStripe; }

extern "C" size_t molclr_ntxent_workspace_bytes(int64_t R, int64_t Rc, int C) {
  const size_t fwd = (size_t)2 * gemm_n_tiles(Rc) * R * sizeof(float);
  const size_t bwd = (size_t)R * kStripe * sizeof(float) + (size_t)num_stripes(Rc) * R * C * sizeof(float);
  return (fwd > bwd ? fwd : bwd) + 256;
}

extern "C" int molclr_ntxent_fwd(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int64_t row_offset,
                                 int64_t row_offset2, float inv_temperature, float* row_lse, float* row_pos, float* loss, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
  MOLCLR_REQUIRE(R > 0 && R % 2 == 0 && Rc >= R && Rc % 4 == 0 && C % 4 == 0, "ntxent: need R > 0 and even, Rc >= R, Rc %% 4 == 0, C %% 4 == 0 (R=%lld Rc=%lld C=%d)",
                 (long long)R, (long long)Rc, C);
  MOLCLR_REQUIRE(workspace_bytes >= molclr_ntxent_workspace_bytes(R, Rc, C), "ntxent_fwd: workspace too small");
  const int tiles = gemm_n_tiles(Rc);
  float* part_max = reinterpret_cast<float*>(workspace);
  float* part_sum = part_max + (size_t)tiles * R;
  GemmJob j;
  memset(&j, 0, sizeof(j));
  j.A = rep; j.lda = C; j.B = cols; j.ldb = C; j.split_k = 1;
  GemmParams& p = j.p;
  p.M = (int)R; p.N = (int)Rc; p.K = C; p.alpha = 1.f;
  p.epi = EPI_NTX_FWD; p.inv_tau = inv_temperature; p.row_offset = row_offset; p.row_split = R / 2; p.row_offset2 = row_offset2; p.col_offset = 0; p.num_cand = Rc;
  p.part_max = part_max; p.part_sum = part_sum; p.row_pos = row_pos;
  int rc = gemm_run(j, stream);
  if (rc) return rc;
  ntx_merge_kernel<<<(int)((R + 31) / 32), dim3(32, 8), 0, stream>>>(part_max, part_sum, tiles, (int)R, row_lse);
  MOLCLR_CHECK_LAUNCH("ntx_merge");
  if (loss) {
    ntx_loss_kernel<<<1, 1024, 0, stream>>>(row_lse, row_pos, (int)R, 1.0f / (float)Rc, loss);
    MOLCLR_CHECK_LAUNCH("ntx_loss");
  }
  return 0;
}

extern "C" int molclr_ntxent_bwd(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int64_t row_offset,
                                 int64_t row_offset2, float inv_temperature, const float* row_lse, const float* col_lse, float gscale, float* g_rep,
                                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MOLCLR_REQUIRE(R > 0 && R % 2 == 0 && Rc >= R && Rc % 4 == 0 && C % 4 == 0, "ntxent: need R > 0 and even, Rc >= R, Rc %% 4 == 0, C %% 4 == 0");
  MOLCLR_REQUIRE(workspace_bytes >= molclr_ntxent_workspace_bytes(R, Rc, C), "ntxent_bwd: workspace too small");
  float* stripe = reinterpret_cast<float*>(workspace);
  float* partials = stripe + (size_t)R * kStripe;                 // [stripes][R][C]
  const int64_t ns = num_stripes(Rc);
  for (int64_t si = 0; si < ns; ++si) {
    const int64_t c0 = si * kStripe;
    const int kc = (int)((Rc - c0) < kStripe ? (Rc - c0) : kStripe);
    GemmJob w;
    memset(&w, 0, sizeof(w));
    w.A = rep; w.lda = C; w.B = cols + (size_t)c0 * C; w.ldb = C; w.split_k = 1;
    w.p.M = (int)R; w.p.N = kc; w.p.K = C; w.p.alpha = 1.f;
    w.p.epi = EPI_NTX_W; w.p.inv_tau = inv_temperature; w.p.row_offset = row_offset; w.p.row_split = R / 2; w.p.row_offset2 = row_offset2; w.p.col_offset = c0; w.p.num_cand = Rc;
    w.p.row_lse = row_lse; w.p.col_lse = col_lse;
    w.p.out = stripe; w.p.ldo = kStripe; w.p.round_out = 1;
    int rc = gemm_run(w, stream);
    if (rc) return rc;
    GemmJob g;
    memset(&g, 0, sizeof(g));
    g.A = stripe; g.lda = kStripe; g.B = cols + (size_t)c0 * C; g.ldb = C; g.split_k = 1;
    g.p.M = (int)R; g.p.N = C; g.p.K = kc; g.p.a_mn = 0; g.p.b_mn = 1;
    g.p.alpha = inv_temperature * gscale; g.p.epi = EPI_GENERIC;
    g.p.out = partials + (size_t)si * R * C; g.p.ldo = C;
    g.bn_hint = 128;             // C = 256 is one 256-wide tile per row tile: halve it so that the stripe fills the GPU
    rc = gemm_run(g, stream);
    if (rc) return rc;
  }
  const long long len4 = (long long)R * C / 4;
  long long blocks = (len4 + 255) / 256;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  ntx_sum_partials_kernel<<<(int)blocks, 256, 0, stream>>>(partials, (int)ns, len4, g_rep);
  MOLCLR_CHECK_LAUNCH("ntx_sum_partials");
  return 0;
}

