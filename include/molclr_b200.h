/* molclr_b200 -- C ABI of the B200-native MolCLR pre-training hot path.
 *
 * The reference (CameronDiao/MolCLR) is pure Python and has no FFI boundary of its own; what this
 * library replaces are the ATen / torch-scatter / cuBLAS calls its Python hot path reaches
 * (SURVEY.md section 2b, rows K1-K18).  Each entry point cites the reference line(s) whose work it
 * does.  The Python classes in molclr_b200/ (GINet, GCN, NTXentLoss -- same constructor signatures
 * and state_dict keys as the reference) call these through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current device unless marked "host";
 *   - the caller owns every buffer, including workspaces; nothing here allocates or frees;
 *   - work is enqueued on `stream`; no entry point synchronises the device or the stream;
 *   - return 0 on success, negative on error; molclr_last_error() (thread-local) has the message;
 *   - feature matrices are row-major fp32 [rows][D] with D % 4 == 0 and 16-byte aligned bases;
 *   - "tf32-rounded" means rounded to nearest to 10 mantissa bits (cvt.rna.tf32.f32): such a tensor is
 *     consumed by the tensor cores without further loss.
 */
#ifndef MOLCLR_B200_H
#define MOLCLR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define MOLCLR_ABI_VERSION 3

/* ---- library ------------------------------------------------------------------------------- */
int molclr_abi_version(void);
const char* molclr_last_error(void);
/* number of kernels this library has launched so far in this process (diagnostics) */
uint64_t molclr_launch_count(void);
/* Programmatic dependent launch (default on): every kernel of the library is launched with programmatic stream serialization, so
 * that its launch overlaps the tail of the kernel before it; each kernel waits for the COMPLETION of its predecessors before its
 * first global-memory access, so results are those of plain stream order.  enable = 0 / 1 sets the switch, < 0 only queries;
 * returns the previous value.  (A measurement aid and an escape hatch; the reference has no counterpart: it is eager PyTorch.) */
int molclr_set_pdl(int enable);
/* host out-params; cc = compute capability major*10+minor */
int molclr_device_info(int* sm_count, int* cc);

/* ---- batch -> CSR plan (once per batch) ----------------------------------------------------------
 * Replaces, for all layers at once: add_self_loops + self_loop_attr + cat (ginet_molclr.py:31-37,
 * gcn_molclr.py:64-70) and the COO gather/scatter bookkeeping of PyG's MessagePassing.propagate.
 * Inputs are the int64 tensors of a PyG Batch (dataset.py:93-109 layout):
 *   x [N,2] (atom 0..118, chirality 0..2), edge_index [2,E], edge_attr [E,2] (type 0..4, dir 0..2),
 *   batch [N] (graph id 0..G-1).
 * Outputs (all int32 unless noted):
 *   xpacked[N] = atom | chirality << 8;  node2graph[N];
 *   rowptr[N+1], col[E] (source of each in-edge), eattr[E] uint8 = type*3+dir: destination-sorted,
 *     in-edges in INPUT order; the self loop is implicit (summed last by the aggregation kernel);
 *   rowptr_t[N+1], col_t[E] (destination of each out-edge): source-sorted transpose for backward;
 *   cnt[N][8] float: in-edge counts per bond type (0..4, self loop = type 4) and direction (5..7), clamped to 2048
 *     (exact in TF32: they are an operand of the table-gradient contraction);
 *   nbr[N][8] uint32 (optional, 32-byte aligned): fixed-width copy of rows with <= 8 in-edges, entry = source << 4 | eattr,
 *     0xFFFFFFFF = empty, [0] = 0xFFFFFFFE = "row too long, use the CSR": lets the aggregation kernel fetch a row's neighbour
 *     list with ONE load (no rowptr -> col dependency); nbr_t[N][8] (optional): the same for the out-edges (destination << 4),
 *     used by the backward aggregation;
 *   gptr[G+1], gperm[N]: nodes grouped by graph (identity permutation for sorted `batch`).
 *   status[4]: [0] = error bits (1 node feature, 2 edge endpoint, 4 edge attr, 8 batch id out of
 *              range, 16 a per-class in-degree above 2048); [1] = 1 if `batch` was not sorted.
 */
size_t molclr_plan_workspace_bytes(int64_t N, int64_t E, int64_t G);
int molclr_plan_build(const int64_t* x, const int64_t* edge_index, const int64_t* edge_attr, const int64_t* batch,
                      int64_t N, int64_t E, int64_t G, int32_t* xpacked, int32_t* node2graph, int32_t* rowptr,
                      int32_t* col, uint8_t* eattr, int32_t* rowptr_t, int32_t* col_t, float* cnt, uint32_t* nbr, uint32_t* nbr_t,
                      int32_t* gptr, int32_t* gperm, void* workspace, size_t workspace_bytes, int32_t* status, cudaStream_t stream);

/* ---- on-device batch construction + augmentation: dataset/dataset.py:112-145 and the DataLoader collate
 * (dataset.py:179-184) from a packed molecule store in HBM ("molclr-packed v1": atom_ptr[M+1], atoms = type | chirality << 8;
 * bond_ptr[M+1], bonds = begin | end << 12 | type << 24 | dir << 27).  For batch slot s = molecule mol_ids[s], views i and j
 * independently: max(1, floor(N/4)) atoms -> [118, 0], floor(M/4) bonds deleted (both directions), survivors in order, each as
 * two consecutive directed edges.  node_off / edge_off / bond_off [B]: exclusive prefix sums of atoms, surviving directed
 * edges 2 (M - floor(M/4)) and bonds M over the batch (host-computed from the pointer arrays).  Outputs are the int64
 * tensors of a PyG Batch (edge_index row-major [2][E_total]).  The random k-subsets are the k smallest of counter-based
 * keys of (seed, view, slot, item); node_masked [2][N_total] / bond_deleted [2][M_total] (optional) export the selection so
 * that a CPU oracle can replay it.  status[0] bit 0: a molecule id out of range. */
int molclr_augment_views(const int32_t* atom_ptr, const int32_t* atoms, const int32_t* bond_ptr, const int32_t* bonds,
                         int64_t n_mols, const int64_t* mol_ids, int64_t B, const int32_t* node_off, const int32_t* edge_off,
                         const int32_t* bond_off, uint64_t seed, int64_t N_total, int64_t E_total, int64_t M_total, int64_t* x_i,
                         int64_t* edge_index_i, int64_t* edge_attr_i, int64_t* batch_i, int64_t* x_j, int64_t* edge_index_j,
                         int64_t* edge_attr_j, int64_t* batch_j, uint8_t* node_masked, uint8_t* bond_deleted, int32_t* status,
                         cudaStream_t stream);

/* ---- subgraph-removal / mixed augmentation: dataset/dataset_subgraph.py:70-88,125-172, dataset/dataset_mix.py:45-68,128-215 ----
 * mode 1: breadth-first removal of floor(0.25 * graph nodes) atoms from a random start atom (removed atoms are masked to [118, 0],
 * not deleted); a bond survives iff neither endpoint was removed AND its begin atom entered the networkx graph before its end atom
 * (the reference's `(start, end) in list(G.edges)` test).  mode 2: removal fraction ~ U(0, 0.2), either orientation survives, then
 * random atom masking / bond deletion up to floor(0.25 N) hidden atoms / ceil(0.75 M) bonds.  The kernel emulates the ordered
 * containers the reference's result depends on (networkx insertion order, CPython's set of small ints): see csrc/augment.cu.
 * Two phases because the number of surviving edges is data dependent: _select writes atoms, batch vectors, per-bond keep flags
 * (0 gone, 1 kept, 2 deleted by the mixed variant's random masking), edge counts / offsets [2][B] and totals[2] (device: the
 * caller reads them to size the edge tensors), plus the draws made (centres [2][B], fractions [2][B] double, removed /
 * extra-masked atoms [2][N]) so that the oracle can replay them; _fill emits the edges.
 * status bits: 1 molecule id out of range, 2 start atom without bonds (the reference raises), 4 molecule too large for the kernel
 * (more than 128 atoms or an atom with more than 8 distinct neighbours). */
int molclr_subgraph_select(const int32_t* atom_ptr, const int32_t* atoms, const int32_t* bond_ptr, const int32_t* bonds,
                           int64_t n_mols, const int64_t* mol_ids, int64_t B, const int32_t* node_off, const int32_t* bond_off,
                           uint64_t seed, int mode, int64_t N_total, int64_t M_total, int64_t* x_i, int64_t* batch_i, int64_t* x_j,
                           int64_t* batch_j, uint8_t* bond_keep, int32_t* edge_count, int32_t* edge_off, int32_t* totals,
                           int32_t* centers, double* percents, uint8_t* removed, uint8_t* extra_masked, int32_t* status,
                           cudaStream_t stream);
int molclr_subgraph_fill(const int32_t* bond_ptr, const int32_t* bonds, int64_t n_mols, const int64_t* mol_ids, int64_t B,
                         const int32_t* node_off, const int32_t* bond_off, const int32_t* edge_off, const uint8_t* bond_keep,
                         int64_t M_total, int64_t* edge_index_i, int64_t* edge_attr_i, int64_t E_i, int64_t* edge_index_j,
                         int64_t* edge_attr_j, int64_t E_j, cudaStream_t stream);

/* ---- node embedding: ginet_molclr.py:103 / gcn_molclr.py:144 ------------------------------------- */
int molclr_embed_nodes_fwd(const int32_t* xpacked, const float* E1, const float* E2, int64_t N, int D, float* out,
                           cudaStream_t stream);
/* dE is ONE buffer [(119+3)][D]: rows 0..118 = grad of x_embedding1, 119..121 = grad of x_embedding2 (g rows ld_g floats apart):
 * embedding_dense_backward of the reference, in plain fp32 with a FIXED summation order (bit-reproducible): chunks of 128 nodes are
 * dealt round-robin to the CTAs; a CTA orders a chunk by atom type (stable in-CTA rank), sums it run by run in registers and adds
 * the runs to its [119][D] shared-memory tile; the per-CTA tiles go to `workspace` and are summed in CTA order.
 * workspace: molclr_embed_nodes_bwd_workspace_bytes(N) bytes, 16-byte aligned. */
size_t molclr_embed_nodes_bwd_workspace_bytes(int64_t N);
int molclr_embed_nodes_bwd(const int32_t* xpacked, const float* g, int64_t ld_g, int64_t N, int D, float* dE, void* workspace,
                           cudaStream_t stream);

/* ---- GINE neighbour aggregation: ginet_molclr.py:39-44 + PyG propagate (index_select, add, scatter_add_) ----
 * out[i] = sum_{in-edges e of i, input order}( f(src[col[e]]) + (B1[t_e] + B2[d_e]) ) + ( f(src[i]) + (B1[4] + B2[0]) )
 * f = identity if bn_coef == NULL, else f(v) = [relu](v*scale + shift) with scale = bn_coef[0..D), shift = bn_coef[D..2D)
 * (the previous layer's BatchNorm + ReLU, ginet_molclr.py:107-111, applied on the fly). */
int molclr_gine_aggregate_fwd(const float* src, const float* bn_coef, int relu, const int32_t* rowptr, const int32_t* col,
                              const uint8_t* eattr, const uint32_t* nbr /* optional, see molclr_plan_build */, const float* B1,
                              const float* B2, int64_t N, int D, float* out,
                              int64_t ld_out /* row stride of out / out_lo in floats */, int round_tf32_out,
                              float* out_lo /* optional: tf32 residual of the exact sum */,
                              uint32_t drop_seed, float drop_p /* dropout of f (see "Dropout" below); 0 = none */, cudaStream_t stream);
/* Backward (autograd of index_select/scatter_add_): gy[j] = sum_{out-edges e of j} ga[col_t[e]] + ga[j].
 * If z_prev != NULL additionally fuses the previous layer's ReLU backward and BatchNorm statistics:
 *   gy[j] *= [z_prev[j]*scale+shift > 0] (if relu);  partials[b][0] += gy, partials[b][1] += gy * (z_prev-mean)*invstd
 * with bn_coef = [scale, shift, mean, invstd][D].  partials: [molclr_rowwise_max_blocks()][2][D]; *num_partials
 * (host) receives the number of partial rows written. */
int molclr_rowwise_max_blocks(void);
int molclr_gine_aggregate_bwd(const float* ga, const int32_t* rowptr_t, const int32_t* col_t,
                              const uint32_t* nbr_t /* optional: selects the shared-memory tile kernel */, const float* z_prev,
                              const float* bn_coef, int relu, int64_t N, int D, float* gy, int round_tf32_out, float* partials,
                              int* num_partials, uint32_t drop_seed, float drop_p, cudaStream_t stream);
/* The ReLU-backward / BatchNorm-statistics stage alone (no neighbour gather): gy = g * [relu mask of z_prev], partials as
 * above.  Used by the GCN backward, where the gradient of a layer input arrives from a GEMM (gcn_molclr.py:76). */
int molclr_relu_bn_bwd_stats(const float* g, const float* z_prev, const float* bn_coef, int relu, int64_t N, int D, float* gy,
                             float* partials, int* num_partials, uint32_t drop_seed, float drop_p, cudaStream_t stream);

/* ---- GCNConv aggregation: gcn_molclr.py:72-88 (scalar bond embeddings [5][1], [3][1] broadcast over the features; the
 * degree normalisation of gcn_molclr.py:27-36,74 is computed and DISCARDED by the reference, so none is applied) ----
 * out[i] = sum_{in-edges e, input order}( src[col[e]] + (b1[t_e] + b2[d_e]) ) + ( src[i] + (b1[4] + b2[0]) ) + bias
 * (src = x @ weight; `out += bias` comes last, gcn_molclr.py:81-82).  bias may be NULL. */
int molclr_gcn_aggregate_fwd(const float* src, const int32_t* rowptr, const int32_t* col, const uint8_t* eattr, const float* b1,
                             const float* b2, const float* bias, int64_t N, int D, float* out, int64_t ld_out, cudaStream_t stream);
/* out[r] = sum_c in[r][c]: collapses the [8][D] result of molclr_edge_table_grad to the GCN's [5][1] + [3][1] gradients */
int molclr_row_sum(const float* in, int R, int C, float* out, cudaStream_t stream);
/* x = [relu](z*scale + shift) (bn_coef NULL: x = z) materialised as a tensor-core operand: hi = tf32(x) (round_hi = 0: x itself,
 * for the compensated GEMM that derives its low halves on chip), lo (optional) = tf32(x - tf32(x)); rows `ld` floats apart.
 * The GCN's GEMM input (gcn_molclr.py:146-152 then :76). */
int molclr_bn_apply_fwd(const float* z, const float* bn_coef, int relu, int64_t N, int D, float* hi, float* lo, int64_t ld,
                        int round_hi, uint32_t drop_seed, float drop_p, cudaStream_t stream);
/* tile_stats [T][2][D]: column mean and M2 of every 32-row group of z (T = molclr_gemm_colstat_tiles(N)): what the GEMM
 * epilogue emits for the GIN path, for outputs that do not come from a GEMM. */
int molclr_bn_tile_stats(const float* z, int64_t N, int D, int T, float* tile_stats, cudaStream_t stream);
/* Gradients of edge_embedding1/2 (embedding_dense_backward over E' rows in the reference):
 * dB [8][D] = cnt^T . ga: rows 0..4 = d edge_embedding1, rows 5..7 = d edge_embedding2 (ga rows ld_ga floats apart).  These are
 * heavy-cancellation sums over all nodes, so they run as exact fp32 FMAs with a FIXED order (bit-reproducible): row tiles are dealt
 * round-robin to the CTAs, each of a CTA's four row lanes sums its rows in increasing order, lanes are added in lane order, the
 * per-CTA [8][D] partials go to `workspace` and are summed in CTA order.
 * workspace: molclr_edge_table_grad_workspace_bytes(D) bytes, 16-byte aligned. */
size_t molclr_edge_table_grad_workspace_bytes(int D);
int molclr_edge_table_grad(const float* ga, int64_t ld_ga, const float* cnt, int64_t N, int D, float* dB, void* workspace,
                           cudaStream_t stream);

/* out[c] (+)= scale * sum_p partials[p][c], p in increasing order (deterministic). */
int molclr_reduce_partials(const float* partials, int P, int len, float scale, int accumulate, float* out,
                           cudaStream_t stream);

/* ---- BatchNorm1d: ginet_molclr.py:79-81,107 (torch defaults eps 1e-5, momentum 0.1) ----------------
 * tile_stats [T][2][D]: per tile of `tile_rows` rows, column mean and M2 (from the GEMM epilogue).
 * coef [4][D] = scale, shift, mean, invstd.  Updates running stats (unbiased var) and num_batches_tracked.
 * Two-level deterministic merge (Chan et al.) through `workspace` (molclr_bn_finalize_workspace_bytes(D) bytes). */
size_t molclr_bn_finalize_workspace_bytes(int D);
int molclr_bn_fwd_finalize(const float* tile_stats, int T, int tile_rows, int64_t N, int D, const float* gamma,
                           const float* beta, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                           float momentum, float eps, float* coef, void* workspace /* 8-byte aligned */, cudaStream_t stream);
int molclr_bn_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, int D, float* coef, cudaStream_t stream);
/* partials [P][2][D] = (sum gy, sum gy*xhat).  Writes dgamma, dbeta and bcoef [3][D] = (k1, A, B) with
 * gz = k1*gy + A + B*z. */
int molclr_bn_bwd_finalize(const float* partials, int P, int64_t N, int D, const float* gamma, const float* coef,
                           int use_batch_stats, float* dgamma, float* dbeta, float* bcoef, cudaStream_t stream);
/* gz (tf32-rounded if round_tf32_out: it then feeds a GEMM) = k1*gy + A + B*z.  gy from memory, or (gp != NULL) gy[n] = gp[node2graph[n]] * w_graph
 * (backward of global_mean/add/max_pool, times the last layer's dropout mask).  dbias (optional) = column sums of gz.  partials [max_blocks][D]. */
int molclr_bn_bwd_apply(const float* gy, const float* gp, const int32_t* node2graph, const int32_t* gptr, int pool_mode,
                        const int32_t* argmax, const float* z, const float* bcoef, int64_t N, int D, float* gz, int64_t ld_gz,
                        int round_tf32_out, float* dbias, float* partials, uint32_t drop_seed, float drop_p, cudaStream_t stream);

/* ---- global_mean_pool / global_add_pool / global_max_pool: ginet_molclr.py:83-88,113 (pool_mode 0 = mean, 1 = add, 2 = max) ----
 * out[g] = w_g * sum (or max) over n in graph g, node order, of dropout([relu](z[n]*scale + shift)).
 * max: argmax [G][D] int32 receives the node that attains the maximum (first one wins); the backward kernels route the
 * gradient to it. */
int molclr_pool_fwd(const float* z, const float* bn_coef, int relu, const int32_t* gptr, const int32_t* gperm,
                    int pool_mode, int64_t G, int D, float* out, int64_t ld_out, int round_tf32_out,
                    float* out_lo /* optional */, int32_t* argmax /* mode 2 */, uint32_t drop_seed, float drop_p, cudaStream_t stream);
int molclr_pool_bwd_stats(const float* gp, const int32_t* node2graph, const int32_t* gptr, int pool_mode, const int32_t* argmax,
                          const float* z, const float* bn_coef, int64_t N, int D, float* partials, int* num_partials,
                          uint32_t drop_seed, float drop_p, cudaStream_t stream);

/* ---- Dropout: ginet_molclr.py:108-111 (F.dropout(relu(BN(z))) per layer, F.dropout(BN(z)) after the last) ----
 * The normalised activations are never stored, so the mask is a counter-based hash of (drop_seed, node, feature): every
 * consumer above that takes (drop_seed, drop_p) recomputes it; kept values are scaled by 1/(1-p).  Use one seed per layer
 * and forward call, and the same seed in the matching backward call.  molclr_dropout_mask writes the mask itself
 * (0 or 1/(1-p)) -- the test oracle applies it explicitly. */
int molclr_dropout_mask(uint32_t drop_seed, float drop_p, int64_t N, int D, float* out, cudaStream_t stream);

/* ---- dense contractions: the nn.Linear / matmul calls of ginet_molclr.py:19-23,46-47,90-96,114-115,
 * gcn_molclr.py:76 and their autograd backward, on tcgen05 tensor cores (TF32 in, FP32 accumulate) ----
 * C[M][N] = sum_k A(m,k) * B(n,k).
 *   a_mn = 0: A is row-major [M][K] (ld = lda);  a_mn = 1: A is row-major [K][M].   Same for B with N.
 * Epilogue, in this order: + bias[n]; + addend[m][n]; relu; * (mask[m][n] > 0) or the mask_bits bit; column statistics of the
 * result (before the optional tf32 rounding of `out`) per 32-row group (colstat_mode 1: sums -> colstat[group][N]; 2: mean and M2 -> colstat[group][2][N]);
 * out = (round_out ? tf32-rounded : exact); out2 = tf32-rounded copy.
 * A_lo/B_lo (same shape and ld as A/B; B_lo alone is allowed, see below): the tf32-rounded residuals x - tf32(x) of the true
 * fp32 operands whose tf32-rounded values are in A/B.  When given, the product is error-compensated,
 * A*B + A_lo*B + A*B_lo (three tensor-core passes, fp32 accumulate, ~fp32 accuracy) -- used by the forward
 * pass so that ReLU masks match the fp32 reference.  out_lo receives the residual of the result.
 * B_lo without A_lo: A holds UNROUNDED fp32 values; the kernel uses the raw tile as the (truncated) hi operand and derives
 * A_lo = tf32(A - trunc_tf32(A)) in shared memory (converter warps), so the producer of A writes one tensor instead of two.
 * split_k > 1 or transpose_out: raw products accumulated atomically into out (zeroed here first);
 * transpose_out stores C^T (out[n][m]).  No other epilogue option is allowed in that mode. */
typedef struct {
  const float* A; int64_t lda; int32_t a_mn;
  const float* B; int64_t ldb; int32_t b_mn;
  const float* A_lo; const float* B_lo;
  int64_t M, N, K;
  float* out; int64_t ldo; int32_t transpose_out;
  float* out2; int64_t ldo2;
  float* out_lo; int64_t ldo_lo;
  const float* bias;
  const float* addend; int64_t ldadd;
  const float* mask; int64_t ldmask;
  int32_t relu, round_out;
  float* colstat; int32_t colstat_mode;
  int32_t split_k;
  uint32_t* relu_bits;        /* out: bit (n % 32) of word [m][n / 32] = (result[m][n] > 0)               */
  const uint32_t* mask_bits;  /* in : result[m][n] = 0 where the bit is clear (ReLU backward), after bias/relu */
  int64_t ld_bits;            /* words per row of either, >= molclr_gemm_mask_words(N)                    */
  int32_t compensate;         /* 1: A and B are UNROUNDED fp32, both K-major, A_lo = B_lo = NULL: ~fp32-accurate product with
                                 every low half derived on chip -- pass 1 in TF32 on the raw tiles (the tensor core truncates),
                                 the corrections (A - trunc A) * B and A * (B - trunc B) as kind::f16 MMAs on bf16 tiles the
                                 kernel forms in shared memory (their 2^-9 rounding applies to terms 2^-10 of the product).
                                 2: the fp16 three-product form of the same product: A (unrounded fp32, K-major) is split on chip
                                 into fp16(a) and fp16(a - fp16(a)) -- 22 significand bits --, B16 (mandatory, b16_kind = 1 of
                                 molclr_prepare_weights) holds B split the same way, B itself is not read (may be NULL), and the
                                 product runs as A_h B_h + A_l B_h + A_h B_l in three kind::f16 MMAs: 3/4 of the tensor time and
                                 2/3 of the operand bytes of form 1, smaller rounding error (hi halves rounded, not truncated;
                                 11-bit corrections).  Price: fp16's RANGE.  |a| > 65504 is clamped (the result is then wrong)
                                 and reported through `status`; halves below 6e-5 are fp16 subnormals with an absolute error of
                                 3e-8, so the relative accuracy degrades gracefully towards TF32's when ALL of A is small
                                 (2e-5 at |a| ~ 1e-3, 2e-4 at 1e-4; TF32: 8e-4).  The activations of
                                 the path (BatchNorm outputs and their neighbourhood sums) sit in the middle of that range.   */
  const void* B16;            /* compensate = 1, optional: the bf16 correction tiles of B pre-split once per optimizer step by
                                 molclr_prepare_weights -- bf16 [2][rows16][ld16]: bf16(B) then bf16(B - trunc_tf32(B)), zero padded;
                                 TMA then lands them in shared memory and the converter warps touch A only (B is a WEIGHT on every
                                 compensated product of the path: it does not change between the row tiles of a launch, nor
                                 between the launches of a step)                                                               */
  int64_t ld16, rows16;       /* row pitch (bf16 elements, % 8 == 0) and rows per half (>= N rounded up to 256)               */
  int32_t* status;            /* compensate = 2, optional: device word; MOLCLR_STATUS_FP16_RANGE is OR-ed in (sticky) when an
                                 element of A exceeded fp16's finite range                                                     */
} molclr_gemm_args;
#define MOLCLR_STATUS_FP16_RANGE 1
#define MOLCLR_H16_SCALE 64   /* the fp16 tiles of a weight hold 2^6 W: the low halves of typical weights (1e-3 .. 1) stay normal
                                 fp16 numbers, |W| up to 1023 stays finite; the GEMM epilogue multiplies by 2^-6 (exact)          */
/* column statistics are emitted per group of molclr_gemm_colstat_tile_rows() (= 32) consecutive rows;
 * molclr_gemm_colstat_tiles(M) groups are written (a multiple of 4; trailing groups may be empty). */
int molclr_gemm_colstat_tiles(int64_t M);
int molclr_gemm_colstat_tile_rows(void);
int molclr_gemm_mask_words(int64_t N);
/* work decomposition of a split-K launch, for callers that size split_k: output tiles of an [M][N] product, concurrent CTAs */
int molclr_gemm_tile_count(int64_t M, int64_t N, int b_mn);
int molclr_gemm_workers(void);
int molclr_gemm_tf32(const molclr_gemm_args* args /* host */, cudaStream_t stream);
/* dW [O][I] (rows ldw apart) = dY^T X for row-major dY [R][O], X [R][I] (tf32-rounded): the weight gradient of a Linear
 * (autograd of ginet_molclr.py:19-23,90-96; gcn_molclr.py:76).  Split-K over R, one wave, atomic accumulation. */
int molclr_gemm_dw(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW,
                   int64_t ldw, cudaStream_t stream);
/* dW += dY^T X: the same launch without the zero-fill of dW (a caller that produces many weight gradients into one buffer zero-fills
 * the buffer ONCE: the per-call 2-D memsets cost more than they look, ~20 us each in the step) */
int molclr_gemm_dw_acc(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW,
                       int64_t ldw, cudaStream_t stream);
/* The same contraction with a FIXED summation order (bit-reproducible run to run): every K split writes its partial [O][I]
 * tile product to `workspace` with plain stores and a second kernel sums the splits in split order.
 * workspace: molclr_gemm_dw_workspace_bytes(R, O, I) bytes, 16-byte aligned. */
size_t molclr_gemm_dw_workspace_bytes(int64_t R, int64_t O, int64_t I);
int molclr_gemm_dw_ordered(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t R, int64_t O, int64_t I, float* dW,
                           int64_t ldw, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- weight shadows: once per forward, ONE launch for every Linear / GCNConv weight of the model -----------------------
 * The tensor-core operand forms of a weight W (nn.Linear [out][in], ginet_molclr.py:19-23,90-96; GCNConv [in][out],
 * gcn_molclr.py:47) are functions of the parameter only, so they are derived once per forward instead of per tile:
 *   hi  [rows][ld_hi]   = tf32(W)            (single-pass products, every backward dX product), same orientation as W
 *   lo  [rows][ld_hi]   = tf32(W - hi)       (explicit 3-pass product of the head)
 *   hi_t [cols][ld_hi_t] = tf32(W^T)         (dX = dY W reads W^T K-major: column tiles of any multiple of 8 instead of whole
 *                                             32-column blocks, i.e. 6 - 12 % instead of 22 - 28 % padded tensor work at N = 300 / 600)
 *   raw [rows_t][ld_raw] = W or W^T (transpose_raw: K-major copy of a weight stored [in][out]), unrounded, 128-byte rows
 *   b16 [2][rows16][ld16] bf16 = bf16(raw), bf16(raw - trunc_tf32(raw)), zero padded (see molclr_gemm_args.B16);
 *       with b16_kind = 1: fp16 instead -- h = fp16(2^6 raw), fp16(2^6 raw - h) (molclr_gemm_args.compensate = 2)
 * Any output pointer may be NULL.  Padding columns (up to the row pitch) are written as zeros. */
typedef struct {
  const float* src; int64_t ld_src; int32_t rows, cols;
  float* hi; float* lo; int64_t ld_hi;
  float* hi_t; int64_t ld_hi_t;                 /* tf32(W^T) [cols][ld_hi_t]: the K-major B operand of the backward dX product */
  float* raw; int64_t ld_raw; int32_t transpose_raw;
  void* b16; int64_t ld16; int32_t rows16;
  int32_t b16_kind;                             /* 0: bf16 correction tiles; 1: fp16 halves of 2^6 W */
} molclr_weight_desc;
int molclr_prepare_weights(const molclr_weight_desc* descs /* host */, int n, cudaStream_t stream);

/* ---- whole-pass entry points: GINet.forward / its backward as ONE call each ---------------------------------------------
 * The kernel sequence of ginet_molclr.py:98-117 (node embedding -> L x (aggregate, MLP, BatchNorm statistics) -> pool) and of
 * its autograd backward, issued from C instead of ~50 / ~80 calls from the host language; same kernels, same order, same results
 * as calling the entry points above one by one.  All pointers are device pointers unless noted; the structs live on the host.
 * Buffers: `ctx` (molclr_gin_ctx_bytes) holds what the backward reads and stays alive between the two passes; `scratch`
 * (molclr_gin_scratch_bytes) is free after each call; `grads` is one flat fp32 buffer (molclr_gin_grad_layout). */
#define MOLCLR_MAX_LAYERS 16
typedef struct {                       /* one GINEConv + BatchNorm1d (ginet_molclr.py:16-27,79-81) */
  const float* w1_hi; const float* w1_raw; const void* w1_b16;     /* mlp.0.weight [2D][D]: molclr_prepare_weights outputs */
  const float* b1;
  const float* w2_hi; const float* w2_raw; const void* w2_b16;     /* mlp.2.weight [D][2D] */
  const float* b2;
  const float* w1_hi_t; const float* w2_hi_t;                      /* tf32(W1^T) [D][2D], tf32(W2^T) [2D][D] (optional: backward dX operands) */
  const float* bond_type; const float* bond_dir;                   /* edge_embedding1 [5][D], edge_embedding2 [3][D] */
  const float* gamma; const float* beta; float* running_mean; float* running_var; int64_t* num_batches_tracked;
  float momentum, eps;
} molclr_gin_layer;
typedef struct {
  int32_t num_layer, emb_dim, feat_dim;
  const float* x_emb1; const float* x_emb2;                         /* x_embedding1 [119][D], x_embedding2 [3][D] */
  const molclr_gin_layer* layers;                                   /* host array [num_layer] */
  int64_t w1_ld16, w1_rows16, w2_ld16, w2_rows16;                   /* extents of the bf16 tiles (see molclr_gemm_args.B16) */
  /* projection head (ginet_molclr.py:90-96): weight shadows hi / lo with row pitch = cols rounded up to 32 floats */
  const float* wf_hi; const float* wf_lo; const float* bf;          /* feat_lin [F][D] */
  const float* w0_hi; const float* w0_lo; const float* b0;          /* out_lin.0 [F][F] */
  const float* w2_hi; const float* w2_lo; const float* b2;          /* out_lin.2 [F/2][F] */
  int32_t* status;                                                  /* optional device word: see molclr_gemm_args.status (comp = 2) */
} molclr_gin_model;
typedef struct {                       /* the outputs of molclr_plan_build */
  int64_t N, E, G;
  const int32_t* xpacked; const int32_t* node2graph; const int32_t* rowptr; const int32_t* col; const uint8_t* eattr;
  const int32_t* rowptr_t; const int32_t* col_t; const float* cnt; const uint32_t* nbr; const uint32_t* nbr_t;
  const int32_t* gptr; const int32_t* gperm;
} molclr_plan_view;
size_t molclr_gin_ctx_bytes(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode);
size_t molclr_gin_scratch_bytes(const molclr_gin_model* m, int64_t N, int64_t G, int ordered);
/* floats of the flat gradient buffer; offsets (host, optional) [2 + 8 L + 6] in the order x_embedding1, x_embedding2, per layer
 * (mlp.0.weight, mlp.0.bias, mlp.2.weight, mlp.2.bias, edge_embedding1, edge_embedding2, bn.weight, bn.bias), feat_lin.weight,
 * .bias, out_lin.0.weight, .bias, out_lin.2.weight, .bias */
int64_t molclr_gin_grad_layout(const molclr_gin_model* m, int64_t* offsets);
/* comp: 1 = error-compensated forward products, TF32 pass + bf16 corrections ("tf32x3"); 2 = the same products in the fp16
 * three-product form ("fp16x3": w1_b16 / w2_b16 are then the fp16 tiles, b16_kind = 1); drop_seeds (host) [L] + drop_p: dropout (NULL / 0 = none) */
int molclr_gin_encoder_fwd(const molclr_gin_model* m, const molclr_plan_view* plan, int comp, int training, int pool_mode,
                           const uint32_t* drop_seeds, float drop_p, void* ctx, size_t ctx_bytes, void* scratch, size_t scratch_bytes,
                           cudaStream_t stream);
/* the pooled operand pair inside ctx (tf32-rounded p [G][ld] and, comp, its residual): the input of any head */
int molclr_gin_ctx_pooled(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode, void* ctx, float** p, float** p_lo,
                          int64_t* ld);
int molclr_proj_head_fwd(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode, void* ctx, float* h /* [G][F] */,
                         float* out /* [G][F/2] */, cudaStream_t stream);
int molclr_proj_head_bwd(const molclr_gin_model* m, int64_t N, int64_t G, int comp, int pool_mode, void* ctx, const float* g_h /* optional */,
                         const float* g_out, int ordered, float* grads, void* scratch, size_t scratch_bytes, cudaStream_t stream);
/* g_p [G][D] = gradient of the pooled vectors; NULL = the one molclr_proj_head_bwd left in `scratch` (same scratch buffer).
 * on_layer_done (optional, host callback): called as (l, user) right after the kernels producing the eight gradients of layer l
 * (l = L-1 .. 0) have been enqueued, and as (-1, user) after the node-embedding gradients: a data-parallel caller launches
 * that slice's all-reduce there, so that it overlaps the rest of the backward pass. */
typedef void (*molclr_layer_cb)(int layer, void* user);
int molclr_gin_encoder_bwd(const molclr_gin_model* m, const molclr_plan_view* plan, int comp, int training, int pool_mode,
                           const uint32_t* drop_seeds, float drop_p, void* ctx, const float* g_p, int ordered, float* grads, void* scratch,
                           size_t scratch_bytes, molclr_layer_cb on_layer_done, void* user, cudaStream_t stream);
/* In-situ timing of the whole-pass calls (a measurement aid; off by default): molclr_step_timing(1) starts collecting CUDA event
 * pairs, on the launching stream, around the BatchNorm-fused aggregation launches (category 0), the first forward MLP products
 * u = relu(a W1^T + b1) (1), the backward row products (2), the weight-gradient products (3) and the second forward products
 * z = u W2^T + b2 (4); (0) stops.  _read waits for a category's events and returns
 * their summed milliseconds and the number of timed launches (host out-params). */
int molclr_step_timing(int enable);
int molclr_step_timing_read(int category, double* total_ms, int* count);
/* y += x (n floats): the gradients of the two views of a step, summed by one launch */
int molclr_add_inplace(float* y, const float* x, int64_t n, cudaStream_t stream);

/* ---- small elementwise ops ---------------------------------------------------------------------- */
/* hi = tf32(src); lo (optional) = tf32(src - hi) */
int molclr_round_tf32(const float* src, float* hi, float* lo, int64_t n, cudaStream_t stream);
/* same for a [rows][cols] matrix with row strides ld_src / ld_dst (hi and lo share ld_dst): lets tensor-core
 * operands be stored with 128-byte aligned rows, which TMA streams markedly faster */
int molclr_round_tf32_2d(const float* src, int64_t ld_src, float* hi, float* lo, int64_t ld_dst, int64_t rows, int64_t cols,
                         cudaStream_t stream);
/* device-to-device 2-D copy (byte pitches / width): pads the 2- or 1-row output layer of the fine-tune head to 4 rows */
int molclr_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height, cudaStream_t stream);
/* activations of the fine-tune prediction head (ginet_finetune.py:96-127): mode 0 = Softplus (beta 1, threshold 20), 1 = ReLU.
 * forward: hi = tf32(act(x)), lo (optional) = tf32 residual; backward: gx = tf32(gy * act'(x)). */
int molclr_act_fwd(const float* x, int mode, int64_t n, float* hi, float* lo, cudaStream_t stream);
int molclr_act_bwd(const float* gy, const float* x, int mode, int64_t n, float* gx, cudaStream_t stream);
/* F.normalize(z, dim=1), molclr.py:63-64 (eps 1e-12) and its backward */
int molclr_l2_normalize_fwd(const float* z, int64_t R, int C, float eps, float* y, float* inv_norm, cudaStream_t stream);
int molclr_l2_normalize_bwd(const float* gy, const float* y, const float* inv_norm, int64_t R, int C, float eps, float* gz,
                            cudaStream_t stream);
/* the same with g_y multiplied by the DEVICE scalar *gscale first (the gradient arriving on a scalar loss: no host round trip) */
int molclr_l2_normalize_bwd_scaled(const float* gy, const float* y, const float* inv_norm, int64_t R, int C, float eps,
                                   const float* gscale, float* gz, cudaStream_t stream);
/* Row preparation of NTXentLoss.forward (nt_xent.py:48 and :40-45) in one pass: rep = cat([zA, zB]) (zA = zjs FIRST), each row
 * divided by max(||row||, eps) when `normalise` (torch.nn.CosineSimilarity, eps 1e-8); y (optional) = the rows, y_r = tf32-rounded
 * (the tensor-core operand), inv_norm (optional) [RA + RB]. */
int molclr_ntxent_rows_fwd(const float* zA, const float* zB, int64_t RA, int64_t RB, int C, float eps, int normalise, float* y,
                           float* y_r /* optional */, float* inv_norm, void* y16 /* optional: fp16 rows, pitch ld16 halves, zero padded */,
                           int64_t ld16, cudaStream_t stream);

/* ---- NT-Xent: utils/nt_xent.py:47-65 -------------------------------------------------------------
 * rep [R][C] fp32, R = 2N rows ordered [zjs; zis] (nt_xent.py:48), already L2-normalised when
 * use_cosine (the cosine similarity's own normalisation is applied by the caller-side kernels
 * molclr_l2_normalize_*).  `cols` [Rc][C] are the candidate rows (== rep for one GPU; the all-gathered
 * projections of every rank for global negatives).  rep is this rank's [zjs_local; zis_local]: its first R/2 rows are
 * candidates row_offset + r, the other R/2 rows candidates row_offset2 + (r - R/2) (one GPU: 0 and R/2); the positive of a
 * local row is the row at the same position of the OTHER block (nt_xent.py:53-55), wherever the blocks of other ranks lie.
 *   forward : row_lse[r] = log sum_{k != self} exp(S[r][k]/tau),  row_pos[r] = S[r][pos(r)]/tau,
 *             loss[0] = sum_r (row_lse[r] - row_pos[r]) / Rc  (this rank's share of the mean over Rc anchors).
 *   backward: g_rep[r] = (gscale/tau) * sum_k (P[r][k] + P[k][r] - 2*[k = pos(r)]) * cols[k],
 *             P[i][k] = exp(S[i][k]/tau - lse[i]) for k != i; col_lse[Rc] holds the log-sum-exp of every
 *             candidate row (== row_lse on one GPU, all-gathered for global negatives); gscale = 1/Rc.
 * The Rc x Rc similarity matrix is never written to memory.  Backward, unit_rows with C <= 256 and 1/tau <= ~22: one fused
 * kernel per call (S tile -> softmax weights -> second product on chip, the [128][C] gradient tile resident in TMEM) over
 * row tiles x candidate splits, whose partial gradients are summed in split order (deterministic).  Otherwise: column
 * stripes of W (2048 fp32 / 4096 fp16 candidates) staged through L2 between two GEMMs, partials summed in stripe order.
 * Requires R even, Rc % 4 == 0, C % 4 == 0.
 * unit_rows != 0 promises that every row of rep and cols has norm <= 1 (the cosine similarity of nt_xent.py:40-45, rows
 * normalised by the caller): the tensor-core passes then run on FP16 copies of the rows (the 11-bit significand of TF32,
 * fp32 accumulation, twice the tensor rate; the softmax weights are staged as fp16 x 2^10).  unit_rows == 0 (dot similarity, nt_xent.py:32-38, rows of
 * any magnitude): TF32 operands, fp32 weights. */
size_t molclr_ntxent_workspace_bytes(int64_t R, int64_t Rc, int C);
/* The same two calls on FP16 operands supplied by the caller (unit-norm rows; row pitch ld16 halves, a multiple of 8): what the
 * data-parallel path all-gathers -- half the NVLink bytes of fp32 rows and no conversion pass over the gathered candidates.
 * Available when molclr_ntxent_h_supported(C, 1/tau) (the fused-backward conditions: C <= 256, C % 8 == 0, 1/tau <= ~22). */
int molclr_ntxent_h_supported(int C, float inv_temperature);
int molclr_ntxent_fwd_h(const void* rep16, const void* cols16, int64_t ld16, int64_t R, int64_t Rc, int C, int64_t row_offset,
                        int64_t row_offset2, float inv_temperature, float* row_lse, float* row_pos, float* loss /* [1], optional */,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
int molclr_ntxent_bwd_h(const void* rep16, const void* cols16, int64_t ld16, int64_t R, int64_t Rc, int C, int64_t row_offset,
                        int64_t row_offset2, float inv_temperature, const float* row_lse, const float* col_lse, float gscale,
                        float* g_rep, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int molclr_ntxent_fwd(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int64_t row_offset,
                      int64_t row_offset2, float inv_temperature, int unit_rows, float* row_lse, float* row_pos,
                      float* loss /* [1], optional */, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int molclr_ntxent_bwd(const float* rep, const float* cols, int64_t R, int64_t Rc, int C, int64_t row_offset,
                      int64_t row_offset2, float inv_temperature, int unit_rows, const float* row_lse, const float* col_lse, float gscale,
                      float* g_rep, void* workspace, size_t workspace_bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MOLCLR_B200_H */
