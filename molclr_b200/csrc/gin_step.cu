// Host-side orchestration of the GIN encoder and the projection head: ONE C call enqueues the whole kernel sequence of
// GINet.forward (ginet_molclr.py:98-117) or of its autograd backward, instead of ~50 / ~80 calls made one by one from Python
// (at 512 pairs per step the Python call overhead, not the GPU, set the step time).  Nothing here touches the device
// directly: it sequences the entry points of this library on the caller's stream, carving every intermediate tensor out of
// caller-owned buffers (`ctx`: what the backward needs; `scratch`: temporaries).
#include <cstring>
#include <vector>

#include "common.cuh"
#include "molclr_b200.h"

namespace molclr {

static inline size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t r32(int64_t x) { return (x + 31) / 32 * 32; }

// Bump allocator over a caller buffer, in floats, 256-byte granules.
struct Carve {
  uint8_t* base; size_t off, cap; bool ok;
  explicit Carve(void* b, size_t cap_) : base(reinterpret_cast<uint8_t*>(b)), off(0), cap(cap_), ok(true) {}
  void* bytes(size_t n) {
    const size_t o = off;
    off = up(off + n, 256);
    if (off > cap) ok = false;
    return base ? base + o : nullptr;
  }
  float* f(size_t n) { return reinterpret_cast<float*>(bytes(n * sizeof(float))); }
};

struct GinDims {
  int64_t N, G; int L, D, H, F, comp, pool_mode;
  int64_t ldD, ldH; int words, T, TG;
};
static GinDims gin_dims(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode) {
  GinDims d;
  d.N = N; d.G = G; d.L = m->num_layer; d.D = m->emb_dim; d.H = 2 * m->emb_dim; d.F = m->feat_dim; d.comp = comp; d.pool_mode = pool_mode;
  d.ldD = r32(d.D); d.ldH = r32(d.H);
  d.words = molclr_gemm_mask_words(d.H);
  d.T = molclr_gemm_colstat_tiles(N); d.TG = molclr_gemm_colstat_tiles(G);
  return d;
}

// ctx: tensors produced by the forward that the backward reads
struct GinCtx {
  float* h0;
  float* a[MOLCLR_MAX_LAYERS]; float* u[MOLCLR_MAX_LAYERS]; uint32_t* ubits[MOLCLR_MAX_LAYERS]; float* z[MOLCLR_MAX_LAYERS]; float* coef[MOLCLR_MAX_LAYERS];
  float* p; float* p_lo; int32_t* argmax;
  float* h_r; float* h_lo; float* r; float* r_lo;      // projection head
};
static bool gin_ctx_carve(const GinDims& d, Carve& c, GinCtx* x) {
  x->h0 = c.f((size_t)d.N * d.D);
  for (int l = 0; l < d.L; ++l) {
    x->a[l] = c.f((size_t)d.N * d.ldD);
    x->u[l] = c.f((size_t)d.N * d.ldH);
    x->ubits[l] = reinterpret_cast<uint32_t*>(c.bytes((size_t)d.N * d.words * 4));
    x->z[l] = c.f((size_t)d.N * d.D);
    x->coef[l] = c.f((size_t)4 * d.D);
  }
  x->p = c.f((size_t)d.G * d.ldD);
  x->p_lo = d.comp ? c.f((size_t)d.G * d.ldD) : nullptr;
  x->argmax = d.pool_mode == 2 ? reinterpret_cast<int32_t*>(c.bytes((size_t)d.G * d.D * 4)) : nullptr;
  x->h_r = c.f((size_t)d.G * d.F);
  x->h_lo = d.comp ? c.f((size_t)d.G * d.F) : nullptr;
  x->r = c.f((size_t)d.G * d.F);
  x->r_lo = d.comp ? c.f((size_t)d.G * d.F) : nullptr;
  return c.ok;
}

static void gemm_args_init(molclr_gemm_args& a) { memset(&a, 0, sizeof(a)); a.split_k = 1; }

// Optional in-situ timing (bench.py's roofline lines): CUDA event pairs around selected launches of the whole-pass calls, on the
// stream they are launched on.  Off by default; costs nothing then.
enum : int { TIME_AGG_FWD = 0, TIME_GEMM_FWD = 1, TIME_GEMM_BWD = 2, TIME_GEMM_DW = 3, TIME_GEMM_FWD2 = 4, TIME_CATS = 5 };
struct StepTimer {
  bool on = false;
  std::vector<cudaEvent_t> ev[TIME_CATS];
};
static StepTimer g_timer;
struct TimeScope {
  int cat; cudaStream_t st; bool on;
  TimeScope(int c, cudaStream_t s, bool enable = true) : cat(c), st(s), on(g_timer.on && enable) {
    if (on) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); g_timer.ev[cat].push_back(e); }
  }
  ~TimeScope() {
    if (on) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); g_timer.ev[cat].push_back(e); }
  }
};

}  // namespace molclr

using namespace molclr;

#define GIN_CALL(expr) do { if (int _rc = (expr)) return _rc; } while (0)

extern "C" size_t molclr_gin_ctx_bytes(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode) {
  const GinDims d = gin_dims(m, N, G, comp, pool_mode);
  Carve c(nullptr, (size_t)-1);
  GinCtx x;
  gin_ctx_carve(d, c, &x);
  return c.off + 256;
}

// temporaries of either pass: BatchNorm tile statistics + merge workspace (forward); gradients in flight, partial sums, the
// table-gradient and ordered split-K workspaces (backward)
extern "C" size_t molclr_gin_scratch_bytes(const molclr_gin_model* m, int64_t N, int64_t G, int ordered) {
  const GinDims d = gin_dims(m, N, G, 1, 0);
  const size_t mb = (size_t)molclr_rowwise_max_blocks();
  size_t fwd = up((size_t)d.T * 2 * d.D * 4, 256) + up(molclr_bn_finalize_workspace_bytes(d.D), 256);
  size_t dw = 0;
  if (ordered) {
    const size_t c1 = molclr_gemm_dw_workspace_bytes(N, d.D, d.H), c2 = molclr_gemm_dw_workspace_bytes(N, d.H, d.D);
    const size_t c3 = molclr_gemm_dw_workspace_bytes(G, d.F, d.F), c4 = molclr_gemm_dw_workspace_bytes(G, d.F, d.D);
    dw = c1 > c2 ? c1 : c2;
    if (c3 > dw) dw = c3;
    if (c4 > dw) dw = c4;
  }
  size_t bwd = up((size_t)N * d.ldD * 4, 256) /* g_z */ + up((size_t)N * d.ldH * 4, 256) /* g_u */ + 2 * up((size_t)N * d.D * 4, 256) /* g_a, g_y */ +
               up((size_t)d.T * d.H * 4, 256) + up(mb * 2 * d.D * 4, 256) + up(mb * d.D * 4, 256) + up((size_t)3 * d.D * 4, 256) +
               up(molclr_edge_table_grad_workspace_bytes(d.D), 256) + up(molclr_embed_nodes_bwd_workspace_bytes(N), 256) + up(dw, 256) +
               up((size_t)G * (d.F / 2) * 4, 256) + 2 * up((size_t)G * d.F * 4, 256) + 2 * up((size_t)d.TG * d.F * 4, 256) + up((size_t)G * d.D * 4, 256);
  return (fwd > bwd ? fwd : bwd) + 4096;
}

// Number of floats of the flat gradient buffer and the offset of every parameter in it, in the order
// [x_embedding1, x_embedding2] + per layer [mlp.0.weight, mlp.0.bias, mlp.2.weight, mlp.2.bias, edge_embedding1, edge_embedding2,
// bn.weight, bn.bias] + [feat_lin.weight, .bias, out_lin.0.weight, .bias, out_lin.2.weight, .bias].
extern "C" int64_t molclr_gin_grad_layout(const molclr_gin_model* m, int64_t* offsets /* [2 + 8 L + 6], optional */) {
  const int64_t D = m->emb_dim, H = 2 * D, F = m->feat_dim;
  int64_t off = 0;
  int k = 0;
  auto put = [&](int64_t n) { if (offsets) offsets[k] = off; ++k; off += n; };
  put(kNumAtomType * D); put(kNumChirality * D);
  for (int l = 0; l < m->num_layer; ++l) { put(H * D); put(H); put(D * H); put(D); put(kNumBondType * D); put(kNumBondDir * D); put(D); put(D); }
  put(F * D); put(F); put(F * F); put(F); put((F / 2) * F); put(F / 2);
  return off;
}

static int gin_check(const molclr_gin_model* m, const molclr_plan_view* pl) {
  MOLCLR_REQUIRE(m && pl && m->layers, "gin: null model / plan");
  MOLCLR_REQUIRE(m->num_layer >= 1 && m->num_layer <= MOLCLR_MAX_LAYERS, "gin: num_layer=%d must be in 1..%d", m->num_layer, MOLCLR_MAX_LAYERS);
  MOLCLR_REQUIRE(m->emb_dim > 0 && m->emb_dim % 4 == 0 && m->emb_dim <= 512 && m->feat_dim > 0 && m->feat_dim % 8 == 0,
                 "gin: emb_dim=%d must be a multiple of 4 (<= 512), feat_dim=%d a multiple of 8", m->emb_dim, m->feat_dim);
  MOLCLR_REQUIRE(pl->N > 0 && pl->G > 0, "gin: empty batch (N=%lld, G=%lld)", (long long)pl->N, (long long)pl->G);
  return 0;
}

extern "C" int molclr_gin_encoder_fwd(const molclr_gin_model* m, const molclr_plan_view* pl, int comp, int training, int pool_mode,
                                      const uint32_t* drop_seeds, float drop_p, void* ctx, size_t ctx_bytes, void* scratch,
                                      size_t scratch_bytes, cudaStream_t stream) {
  GIN_CALL(gin_check(m, pl));
  const GinDims d = gin_dims(m, pl->N, pl->G, comp, pool_mode);
  Carve cc(ctx, ctx_bytes);
  GinCtx x;
  MOLCLR_REQUIRE(gin_ctx_carve(d, cc, &x), "gin_encoder_fwd: ctx too small (%zu bytes, need molclr_gin_ctx_bytes)", ctx_bytes);
  Carve sc(scratch, scratch_bytes);
  float* stats = sc.f((size_t)d.T * 2 * d.D);
  void* bn_ws = sc.bytes(molclr_bn_finalize_workspace_bytes(d.D));
  MOLCLR_REQUIRE(sc.ok, "gin_encoder_fwd: scratch too small");
  const int64_t N = d.N;
  const int D = d.D, H = d.H, L = d.L;
  auto seed = [&](int l) -> uint32_t { return (drop_seeds && drop_p > 0.f) ? drop_seeds[l] : 0u; };
  const float dp = (drop_seeds && drop_p > 0.f) ? drop_p : 0.f;
  GIN_CALL(molclr_embed_nodes_fwd(pl->xpacked, m->x_emb1, m->x_emb2, N, D, x.h0, stream));          // ginet_molclr.py:103
  const float* src = x.h0;
  const float* coef_prev = nullptr;
  for (int l = 0; l < L; ++l) {
    const molclr_gin_layer& ly = m->layers[l];
    // a_l = sum_j f(z_{l-1})[j] + bond table, self loop last (:29-44); f = BatchNorm + ReLU (+ dropout) of layer l-1, fused
    {
      TimeScope ts(TIME_AGG_FWD, stream, l > 0);
      GIN_CALL(molclr_gine_aggregate_fwd(src, coef_prev, 1, pl->rowptr, pl->col, pl->eattr, pl->nbr, ly.bond_type, ly.bond_dir, N, D, x.a[l], d.ldD, 0,
                                         nullptr, l > 0 ? seed(l - 1) : 0u, l > 0 ? dp : 0.f, stream));
    }
    molclr_gemm_args g;
    gemm_args_init(g);                                                                                // u = relu(a W1^T + b1)  (:19-23,46-47)
    g.A = x.a[l]; g.lda = d.ldD; g.B = comp == 2 ? nullptr : comp ? ly.w1_raw : ly.w1_hi; g.ldb = d.ldD; g.M = N; g.N = H; g.K = D;
    g.compensate = comp; g.B16 = comp ? ly.w1_b16 : nullptr; g.ld16 = m->w1_ld16; g.rows16 = m->w1_rows16; g.status = m->status;
    g.out = x.u[l]; g.ldo = d.ldH; g.bias = ly.b1; g.relu = 1; g.relu_bits = x.ubits[l]; g.ld_bits = d.words;
    { TimeScope ts(TIME_GEMM_FWD, stream); GIN_CALL(molclr_gemm_tf32(&g, stream)); }
    gemm_args_init(g);                                                                                // z = u W2^T + b2 (+ BatchNorm tile statistics)
    g.A = x.u[l]; g.lda = d.ldH; g.B = comp == 2 ? nullptr : comp ? ly.w2_raw : ly.w2_hi; g.ldb = d.ldH; g.M = N; g.N = D; g.K = H;
    g.compensate = comp; g.B16 = comp ? ly.w2_b16 : nullptr; g.ld16 = m->w2_ld16; g.rows16 = m->w2_rows16; g.status = m->status;
    g.out = x.z[l]; g.ldo = D; g.bias = ly.b2;
    if (training) { g.colstat = stats; g.colstat_mode = 2; }
    { TimeScope ts(TIME_GEMM_FWD2, stream); GIN_CALL(molclr_gemm_tf32(&g, stream)); }
    if (training)                                                                                     // :107
      GIN_CALL(molclr_bn_fwd_finalize(stats, d.T, molclr_gemm_colstat_tile_rows(), N, D, ly.gamma, ly.beta, ly.running_mean, ly.running_var,
                                      ly.num_batches_tracked, ly.momentum, ly.eps, x.coef[l], bn_ws, stream));
    else
      GIN_CALL(molclr_bn_eval_coef(ly.gamma, ly.beta, ly.running_mean, ly.running_var, ly.eps, D, x.coef[l], stream));
    src = x.z[l]; coef_prev = x.coef[l];
  }
  // p = pool( dropout( BN_L(z_L) ) )  (:110-113: no ReLU after the last layer)
  GIN_CALL(molclr_pool_fwd(src, coef_prev, 0, pl->gptr, pl->gperm, pool_mode, d.G, D, x.p, d.ldD, 1, x.p_lo, x.argmax, seed(L - 1), dp, stream));
  return 0;
}

// pointers into ctx a caller needs: the pooled operand pair (input of any head)
extern "C" int molclr_gin_ctx_pooled(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode, void* ctx, float** p, float** p_lo,
                                     int64_t* ld) {
  const GinDims d = gin_dims(m, N, G, comp, pool_mode);
  Carve cc(ctx, (size_t)-1);
  GinCtx x;
  gin_ctx_carve(d, cc, &x);
  if (p) *p = x.p;
  if (p_lo) *p_lo = x.p_lo;
  if (ld) *ld = d.ldD;
  return 0;
}

// h = feat_lin(p); out = out_lin(h)   (ginet_molclr.py:90-96,114-115)
extern "C" int molclr_proj_head_fwd(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode, void* ctx, float* h, float* out,
                                    cudaStream_t stream) {
  const GinDims d = gin_dims(m, N, G, comp, pool_mode);
  Carve cc(ctx, (size_t)-1);
  GinCtx x;
  gin_ctx_carve(d, cc, &x);
  const int D = d.D, F = d.F;
  molclr_gemm_args g;
  gemm_args_init(g);
  g.A = x.p; g.lda = d.ldD; g.A_lo = x.p_lo; g.B = m->wf_hi; g.B_lo = comp ? m->wf_lo : nullptr; g.ldb = r32(D); g.M = d.G; g.N = F; g.K = D;
  g.out = h; g.ldo = F; g.out2 = x.h_r; g.ldo2 = F; g.out_lo = x.h_lo; g.ldo_lo = F; g.bias = m->bf;
  GIN_CALL(molclr_gemm_tf32(&g, stream));
  gemm_args_init(g);
  g.A = x.h_r; g.lda = F; g.A_lo = x.h_lo; g.B = m->w0_hi; g.B_lo = comp ? m->w0_lo : nullptr; g.ldb = r32(F); g.M = d.G; g.N = F; g.K = F;
  g.out = x.r; g.ldo = F; g.out_lo = x.r_lo; g.ldo_lo = F; g.bias = m->b0; g.relu = 1; g.round_out = 1;
  GIN_CALL(molclr_gemm_tf32(&g, stream));
  gemm_args_init(g);
  g.A = x.r; g.lda = F; g.A_lo = x.r_lo; g.B = m->w2_hi; g.B_lo = comp ? m->w2_lo : nullptr; g.ldb = r32(F); g.M = d.G; g.N = F / 2; g.K = F;
  g.out = out; g.ldo = F / 2; g.bias = m->b2;
  GIN_CALL(molclr_gemm_tf32(&g, stream));
  return 0;
}

static int dw(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* out, int ordered, void* ws,
              size_t ws_bytes, cudaStream_t stream) {
  if (ordered) return molclr_gemm_dw_ordered(dY, ldy, X, ldx, R, O, I, out, I, ws, ws_bytes, stream);
  return molclr_gemm_dw_acc(dY, ldy, X, ldx, R, O, I, out, I, stream);       // (the gradient slices were zero-filled once, see zero_slices)
}

// The weight-gradient products accumulate atomically into the flat gradient buffer: ONE contiguous zero-fill per backward call instead of a
// 2-D memset in front of every product.
static int zero_slices(float* grads, int64_t lo, int64_t hi, int ordered, cudaStream_t stream) {
  if (ordered || hi <= lo) return 0;
  cudaError_t e = cudaMemsetAsync(grads + lo, 0, (size_t)(hi - lo) * sizeof(float), stream);
  if (e != cudaSuccess) return cuda_fail(e, "gin backward: zero-filling the gradient buffer");
  return 0;
}

// scratch carving shared by the two backward entry points (so that the head's g_p is where the encoder expects it)
struct GinBwdScratch {
  float* g_z; float* g_u; float* g_a; float* g_y; float* part; float* partials; float* partials2; float* bcoef;
  void* tab_ws; void* emb_ws; void* dw_ws; size_t dw_bytes;
  float* g_out_r; float* g_r; float* g_hh_r; float* hpart; float* hpart2; float* g_p;
};
static bool gin_bwd_carve(const GinDims& d, int ordered, Carve& c, GinBwdScratch* s) {
  const size_t mb = (size_t)molclr_rowwise_max_blocks();
  s->g_z = c.f((size_t)d.N * d.ldD); s->g_u = c.f((size_t)d.N * d.ldH); s->g_a = c.f((size_t)d.N * d.D); s->g_y = c.f((size_t)d.N * d.D);
  s->part = c.f((size_t)d.T * d.H); s->partials = c.f(mb * 2 * d.D); s->partials2 = c.f(mb * d.D); s->bcoef = c.f((size_t)3 * d.D);
  s->tab_ws = c.bytes(molclr_edge_table_grad_workspace_bytes(d.D));
  s->emb_ws = c.bytes(molclr_embed_nodes_bwd_workspace_bytes(d.N));
  s->dw_bytes = 0;
  if (ordered) {
    const size_t c1 = molclr_gemm_dw_workspace_bytes(d.N, d.D, d.H), c2 = molclr_gemm_dw_workspace_bytes(d.N, d.H, d.D);
    const size_t c3 = molclr_gemm_dw_workspace_bytes(d.G, d.F, d.F), c4 = molclr_gemm_dw_workspace_bytes(d.G, d.F, d.D);
    s->dw_bytes = c1 > c2 ? c1 : c2;
    if (c3 > s->dw_bytes) s->dw_bytes = c3;
    if (c4 > s->dw_bytes) s->dw_bytes = c4;
  }
  s->dw_ws = c.bytes(s->dw_bytes);
  s->g_out_r = c.f((size_t)d.G * (d.F / 2)); s->g_r = c.f((size_t)d.G * d.F); s->g_hh_r = c.f((size_t)d.G * d.F);
  s->hpart = c.f((size_t)d.TG * d.F); s->hpart2 = c.f((size_t)d.TG * d.F); s->g_p = c.f((size_t)d.G * d.D);
  return c.ok;
}

// Backward of molclr_proj_head_fwd: writes the six head gradients into `grads` (flat layout of molclr_gin_grad_layout) and the
// gradient of the pooled vectors into the scratch slot molclr_gin_encoder_bwd reads (g_p = NULL there).
extern "C" int molclr_proj_head_bwd(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode, void* ctx, const float* g_h /* optional */,
                                    const float* g_out, int ordered, float* grads, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  const GinDims d = gin_dims(m, N, G, comp, pool_mode);
  Carve cc(ctx, (size_t)-1);
  GinCtx x;
  gin_ctx_carve(d, cc, &x);
  Carve sc(scratch, scratch_bytes);
  GinBwdScratch s;
  MOLCLR_REQUIRE(gin_bwd_carve(d, ordered, sc, &s), "proj_head_bwd: scratch too small (need molclr_gin_scratch_bytes)");
  int64_t off[2 + 8 * MOLCLR_MAX_LAYERS + 6];
  molclr_gin_grad_layout(m, off);
  const int hb = 2 + 8 * d.L;
  float* dWf = grads + off[hb]; float* dbf = grads + off[hb + 1]; float* dW0 = grads + off[hb + 2]; float* db0 = grads + off[hb + 3];
  float* dW2 = grads + off[hb + 4]; float* db2 = grads + off[hb + 5];
  const int D = d.D, F = d.F, F2 = d.F / 2;
  const int64_t Gn = d.G;
  GIN_CALL(zero_slices(grads, off[hb], molclr_gin_grad_layout(m, nullptr), ordered, stream));
  GIN_CALL(molclr_round_tf32(g_out, s.g_out_r, nullptr, Gn * F2, stream));
  GIN_CALL(dw(s.g_out_r, F2, x.r, F, Gn, F2, F, dW2, ordered, s.dw_ws, s.dw_bytes, stream));
  GIN_CALL(molclr_reduce_partials(g_out, (int)Gn, F2, 1.f, 0, db2, stream));
  molclr_gemm_args g;
  gemm_args_init(g);                                  // g_r = (g_out W2) * [r > 0], db0 = column sums
  g.A = s.g_out_r; g.lda = F2; g.B = m->w2_hi; g.ldb = r32(F); g.b_mn = 1; g.M = Gn; g.N = F; g.K = F2;
  g.out = s.g_r; g.ldo = F; g.mask = x.r; g.ldmask = F; g.round_out = 1; g.colstat = s.hpart; g.colstat_mode = 1;
  GIN_CALL(molclr_gemm_tf32(&g, stream));
  GIN_CALL(molclr_reduce_partials(s.hpart, d.TG, F, 1.f, 0, db0, stream));
  GIN_CALL(dw(s.g_r, F, x.h_r, F, Gn, F, F, dW0, ordered, s.dw_ws, s.dw_bytes, stream));
  gemm_args_init(g);                                  // g_h(total) = g_r W0 (+ the gradient arriving on the returned representation h)
  g.A = s.g_r; g.lda = F; g.B = m->w0_hi; g.ldb = r32(F); g.b_mn = 1; g.M = Gn; g.N = F; g.K = F;
  g.out2 = s.g_hh_r; g.ldo2 = F; g.addend = g_h; g.ldadd = F; g.colstat = s.hpart2; g.colstat_mode = 1;
  GIN_CALL(molclr_gemm_tf32(&g, stream));
  GIN_CALL(molclr_reduce_partials(s.hpart2, d.TG, F, 1.f, 0, dbf, stream));
  GIN_CALL(dw(s.g_hh_r, F, x.p, d.ldD, Gn, F, D, dWf, ordered, s.dw_ws, s.dw_bytes, stream));
  gemm_args_init(g);
  g.A = s.g_hh_r; g.lda = F; g.B = m->wf_hi; g.ldb = r32(D); g.b_mn = 1; g.M = Gn; g.N = D; g.K = F;
  g.out = s.g_p; g.ldo = D;
  GIN_CALL(molclr_gemm_tf32(&g, stream));
  return 0;
}

// Backward of molclr_gin_encoder_fwd.  g_p [G][D]: gradient of the pooled vectors (NULL: the one molclr_proj_head_bwd left in
// scratch).  Writes the 2 + 8 L encoder gradients into `grads`.
extern "C" int molclr_gin_encoder_bwd(const molclr_gin_model* m, const molclr_plan_view* pl, int comp, int training, int pool_mode,
                                      const uint32_t* drop_seeds, float drop_p, void* ctx, const float* g_p, int ordered, float* grads,
                                      void* scratch, size_t scratch_bytes, molclr_layer_cb on_layer_done, void* user, cudaStream_t stream) {
  GIN_CALL(gin_check(m, pl));
  const GinDims d = gin_dims(m, pl->N, pl->G, comp, pool_mode);
  Carve cc(ctx, (size_t)-1);
  GinCtx x;
  gin_ctx_carve(d, cc, &x);
  Carve sc(scratch, scratch_bytes);
  GinBwdScratch s;
  MOLCLR_REQUIRE(gin_bwd_carve(d, ordered, sc, &s), "gin_encoder_bwd: scratch too small (need molclr_gin_scratch_bytes)");
  if (!g_p) g_p = s.g_p;
  int64_t off[2 + 8 * MOLCLR_MAX_LAYERS + 6];
  molclr_gin_grad_layout(m, off);
  const int64_t N = d.N;
  const int D = d.D, H = d.H, L = d.L;
  auto seed = [&](int l) -> uint32_t { return (drop_seeds && drop_p > 0.f) ? drop_seeds[l] : 0u; };
  const float dp = (drop_seeds && drop_p > 0.f) ? drop_p : 0.f;
  auto G_ = [&](int l, int k) -> float* { return grads + off[2 + 8 * l + k]; };     // k: 0 W1, 1 b1, 2 W2, 3 b2, 4 E1, 5 E2, 6 gamma, 7 beta
  int P = 0;
  GIN_CALL(zero_slices(grads, 0, off[2 + 8 * L], ordered, stream));
  // last layer: BatchNorm backward fed by the pool backward (the pooled gradient is expanded on the fly)
  GIN_CALL(molclr_pool_bwd_stats(g_p, pl->node2graph, pl->gptr, pool_mode, x.argmax, x.z[L - 1], x.coef[L - 1], N, D, s.partials, &P, seed(L - 1), dp, stream));
  GIN_CALL(molclr_bn_bwd_finalize(s.partials, P, N, D, m->layers[L - 1].gamma, x.coef[L - 1], training, G_(L - 1, 6), G_(L - 1, 7), s.bcoef, stream));
  GIN_CALL(molclr_bn_bwd_apply(nullptr, g_p, pl->node2graph, pl->gptr, pool_mode, x.argmax, x.z[L - 1], s.bcoef, N, D, s.g_z, d.ldD, 1, G_(L - 1, 3),
                               s.partials2, seed(L - 1), dp, stream));
  for (int l = L - 1; l >= 0; --l) {
    const molclr_gin_layer& ly = m->layers[l];
    molclr_gemm_args g;
    gemm_args_init(g);                                // g_u = (g_z W2) * [u > 0];  db1 = colsum(g_u)
    g.A = s.g_z; g.lda = d.ldD; g.M = N; g.N = H; g.K = D;
    if (ly.w2_hi_t) { g.B = ly.w2_hi_t; g.ldb = d.ldD; }              // W2^T [H][D] K-major: 224-column tiles (12 % padding) instead of 256 (28 %)
    else { g.B = ly.w2_hi; g.ldb = d.ldH; g.b_mn = 1; }
    g.out = s.g_u; g.ldo = d.ldH; g.mask_bits = x.ubits[l]; g.ld_bits = d.words; g.round_out = 1; g.colstat = s.part; g.colstat_mode = 1;
    { TimeScope ts(TIME_GEMM_BWD, stream); GIN_CALL(molclr_gemm_tf32(&g, stream)); }
    GIN_CALL(molclr_reduce_partials(s.part, d.T, H, 1.f, 0, G_(l, 1), stream));
    { TimeScope ts(TIME_GEMM_DW, stream); GIN_CALL(dw(s.g_z, d.ldD, x.u[l], d.ldH, N, D, H, G_(l, 2), ordered, s.dw_ws, s.dw_bytes, stream)); }   // dW2 [D][H]
    gemm_args_init(g);                                // g_a = g_u W1
    g.A = s.g_u; g.lda = d.ldH; g.M = N; g.N = D; g.K = H;
    if (ly.w1_hi_t) { g.B = ly.w1_hi_t; g.ldb = d.ldH; }              // W1^T [D][H] K-major: 160-column tiles (6 % padding) instead of 192 (22 %)
    else { g.B = ly.w1_hi; g.ldb = d.ldD; g.b_mn = 1; }
    g.out = s.g_a; g.ldo = D;
    { TimeScope ts(TIME_GEMM_BWD, stream); GIN_CALL(molclr_gemm_tf32(&g, stream)); }
    { TimeScope ts(TIME_GEMM_DW, stream); GIN_CALL(dw(s.g_u, d.ldH, x.a[l], d.ldD, N, H, D, G_(l, 0), ordered, s.dw_ws, s.dw_bytes, stream)); }   // dW1 [H][D]
    GIN_CALL(molclr_edge_table_grad(s.g_a, D, pl->cnt, N, D, G_(l, 4), s.tab_ws, stream));                  // [8][D]: edge_embedding1 | edge_embedding2
    if (on_layer_done) on_layer_done(l, user);        // all eight gradients of layer l are enqueued (data-parallel: launch their all-reduce now)
    if (l > 0) {
      GIN_CALL(molclr_gine_aggregate_bwd(s.g_a, pl->rowptr_t, pl->col_t, pl->nbr_t, x.z[l - 1], x.coef[l - 1], 1, N, D, s.g_y, 0, s.partials, &P,
                                         seed(l - 1), dp, stream));
      GIN_CALL(molclr_bn_bwd_finalize(s.partials, P, N, D, m->layers[l - 1].gamma, x.coef[l - 1], training, G_(l - 1, 6), G_(l - 1, 7), s.bcoef, stream));
      GIN_CALL(molclr_bn_bwd_apply(s.g_y, nullptr, nullptr, nullptr, 0, nullptr, x.z[l - 1], s.bcoef, N, D, s.g_z, d.ldD, 1, G_(l - 1, 3), s.partials2,
                                   0u, 0.f, stream));
    } else {
      GIN_CALL(molclr_gine_aggregate_bwd(s.g_a, pl->rowptr_t, pl->col_t, pl->nbr_t, nullptr, nullptr, 1, N, D, s.g_y, 0, nullptr, &P, 0u, 0.f, stream));
      GIN_CALL(molclr_embed_nodes_bwd(pl->xpacked, s.g_y, D, N, D, grads + off[0], s.emb_ws, stream));      // [122][D]: x_embedding1 | x_embedding2
      if (on_layer_done) on_layer_done(-1, user);     // the node-embedding tables
    }
  }
  return 0;
}

// In-situ timing of the whole-pass calls (measurement aid of bench.py): molclr_step_timing(1) starts collecting CUDA event pairs around
// the BatchNorm-fused aggregation launches (category 0), the first forward MLP products u = relu(a W1^T + b1) (1), the backward row
// products (2), the weight-gradient products (3) and the second forward products z = u W2^T + b2 (4); molclr_step_timing(0) stops.  molclr_step_timing_read synchronises the events of a category and
// returns their summed milliseconds and the number of timed launches.
extern "C" int molclr_step_timing(int enable) {
  for (int c = 0; c < TIME_CATS; ++c) {
    for (cudaEvent_t e : g_timer.ev[c]) cudaEventDestroy(e);
    g_timer.ev[c].clear();
  }
  g_timer.on = enable != 0;
  return 0;
}

extern "C" int molclr_step_timing_read(int category, double* total_ms, int* count) {
  MOLCLR_REQUIRE(category >= 0 && category < TIME_CATS && total_ms && count, "step_timing_read: bad arguments");
  const std::vector<cudaEvent_t>& ev = g_timer.ev[category];
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < ev.size(); i += 2) {
    cudaError_t e = cudaEventSynchronize(ev[i + 1]);
    if (e != cudaSuccess) return cuda_fail(e, "step_timing_read");
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
    if (e != cudaSuccess) return cuda_fail(e, "step_timing_read");
    tot += ms; ++n;
  }
  *total_ms = tot; *count = n;
  return 0;
}
