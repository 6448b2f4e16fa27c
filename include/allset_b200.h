/* allset_b200.h -- C ABI of liballset_b200.so (sm_100a).
 *
 * The drop-in boundary for AllSet's V->E / E->V multiset-aggregation path.  The reference
 * (jianhao2016/AllSet) has no FFI layer of its own: the path bottoms out in three third-party
 * Python entry points.  Each function below names the reference call it replaces
 * (paths relative to the reference tree):
 *
 *   torch_scatter.scatter(src, index, dim, reduce)    src/layers.py:656 (HalfNLHconv.aggregate)
 *                                                     src/layers.py:194 (PMA.aggregate)
 *   MessagePassing.propagate -> index_select gather   src/layers.py:633, :145
 *   norm.view(-1,1) * x_j                             src/layers.py:638-639 (HalfNLHconv.message)
 *   torch_geometric.utils.softmax(alpha, index, ...)  src/layers.py:168-177 (PMA.message)
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every pointer is a DEVICE pointer owned by the caller
 *     (PyTorch's caching allocator in the Python host); the library never allocates, frees or
 *     retains memory.  Workspace sizes are queried first and passed in.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation, no host reads of device data.
 *   - Return value: 0 on success, a negative ALLSET_E* code otherwise; the message is available
 *     from allset_last_error() (thread-local).  No C++ exception crosses the boundary.
 *   - Incidence lists are CSR by TARGET row: rowptr[n_tgt+1] (int32), col[nnz] (int32 source row
 *     of each incidence), in the stable order of the caller's COO list, so per-segment summation
 *     order equals the reference's CPU scatter_add_ order.
 *   - dtype: ALLSET_F32 or ALLSET_BF16 is the STORAGE type of feature rows; accumulation is
 *     always fp32.  Scores, weights and statistics are fp32.
 */
#ifndef ALLSET_B200_H_
#define ALLSET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALLSET_ABI_VERSION 3

enum { ALLSET_F32 = 0, ALLSET_BF16 = 1 };
enum { ALLSET_SUM = 0, ALLSET_MEAN = 1 };
enum {
  ALLSET_OK = 0,
  ALLSET_EINVAL = -1,   /* bad argument (null pointer, negative size, unknown enum) */
  ALLSET_ERANGE = -2,   /* size does not fit the int32 CSR (nnz or rows >= 2^31) */
  ALLSET_EWORKSPACE = -3, /* workspace too small */
  ALLSET_ECUDA = -4,    /* CUDA launch / runtime error (text in allset_last_error) */
  ALLSET_EUNSUPPORTED = -5 /* shape not handled by the requested (fused-exchange) variant; use the plain call */
};

/* ABI version of the loaded library (== ALLSET_ABI_VERSION it was built with). */
int allset_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* allset_last_error(void);

/* 1 when (dtype, d, n_tgt) takes the STREAM kernels (rows of 128/256/512/1024 bytes, enough target segments).  Given a
 * workspace (below) those kernels reduce segments of ANY length at full rate -- a chunk boundary that falls inside a long
 * segment cuts it and the pieces are combined in stream order -- so a caller should then pass n_long = 0 instead of
 * bucketing long segments for the CTA kernels.  allset_pma_stream_eligible: the same question for the PMA kernel, which
 * additionally needs H % 4 == 0 and lane chunks that do not straddle heads. */
int allset_stream_eligible(int dtype, int32_t d, int64_t n_tgt);
int allset_pma_stream_eligible(int dtype, int32_t H, int32_t C, int64_t n_tgt);

/* Bytes of the optional workspace of the stream kernels for rows of d elements (any dtype, any H).  The caller
 * zero-initialises it ONCE; every launch leaves it zeroed, so it can be reused by successive launches on one stream
 * (not by concurrent launches on different streams).  It holds the partial results and ready flags of the pieces of
 * long segments that were cut at a warp-chunk boundary.  ws == NULL: segments are never cut. */
size_t allset_stream_workspace_bytes(int32_t d);

/* --- incidence container ------------------------------------------------------------------
 * Replaces the implicit work torch_scatter does on every call (unsorted COO index + a
 * D2H `index.max()` to size the output, src/layers.py:656): the COO list is sorted ONCE per
 * graph into CSR-by-target.  `perm[k]` = position in the caller's COO list of CSR slot k
 * (needed to carry per-incidence tensors -- data.norm, SetGNN.Importance, PMA's returned
 * alpha -- between the two orders).  Stable: equal targets keep COO order. */
size_t allset_csr_workspace_bytes(int64_t nnz, int64_t n_tgt);
int allset_csr_from_coo(const int64_t* tgt, const int64_t* src, int64_t nnz, int64_t n_tgt,
                        int32_t* rowptr, int32_t* col, int32_t* perm,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Segments longer than `threshold` get a CTA each instead of a lane group.  Writes their row
 * ids (ascending) to long_ids[capacity] and their count to *n_long (device int32). */
size_t allset_long_segments_workspace_bytes(int64_t n_tgt);
int allset_long_segments(const int32_t* rowptr, int64_t n_tgt, int32_t threshold,
                         int32_t* long_ids, int32_t* n_long,
                         void* workspace, size_t workspace_bytes, void* stream);

/* --- AllDeepSets: gather + weight + segmented sum / mean -----------------------------------
 * out[t, :] = reduce_{k in [rowptr[t], rowptr[t+1])} w[k] * src_scale[col[k]] * x[col[k], :]
 * Replaces propagate -> message -> aggregate of HalfNLHconv (src/layers.py:633,638-639,641-656)
 * i.e. index_select + norm*x_j + torch_scatter.scatter(reduce='add'|'sum'|'mean') in ONE launch.
 *   x         [n_src, d]  feature rows (dtype)          out  [n_tgt, d] (dtype), fully written
 *   w         [nnz] fp32 in CSR order, or NULL (= data.norm all ones, preprocessing.py:454)
 *   src_scale [n_src] fp32 or NULL (used by the backward of 'mean': 1/max(count,1) of the row)
 *   op        ALLSET_SUM | ALLSET_MEAN (mean divides by max(segment length, 1): torch_scatter)
 *   long_ids / n_long / long_threshold: output of allset_long_segments (n_long may be 0, then
 *   long_ids may be NULL).
 * The gradient w.r.t. x is this same function on the transposed CSR. */
int allset_segreduce_fwd(const void* x, int dtype, int64_t n_src, int32_t d,
                         const int32_t* rowptr, const int32_t* col,
                         const float* w, const float* src_scale,
                         int64_t n_tgt, int op,
                         const int32_t* long_ids, int32_t n_long, int32_t long_threshold,
                         void* out, void* ws, size_t ws_bytes, void* stream);

/* Fused compute + exchange (multi-GPU, one process per GPU; no counterpart in the single-device reference).
 * Same reduction as allset_segreduce_fwd, but every reduced row is ALSO sent, from the kernel's epilogue, to the same row
 * of `n_peers` (<= 7) peer replicas over NVLink, so the all-gather of the rank's row range that would follow the kernel
 * costs no extra launch and overlaps the reduce.  By default the reducing warp stores the row to the peers itself
 * (st.global on peer addresses, fire-and-forget); ALLSET_PUSH=bulk parks the row in a shared-memory staging slot and hands it
 * to the TMA instead (cp.async.bulk shared -> global, one bulk store per peer), which takes NVLink back-pressure off the
 * gathering warps at the price of a staging pass (measured slower while the kernel is compute-bound).
 *   out          this rank's rows inside its own replica (row 0 of the rank's range)
 *   peer_outs    HOST array of n_peers peer-mapped DEVICE pointers: the address of that same row in each peer replica
 *                (or ONE multicast address that the NVSwitch replicates into every replica)
 *   peer_mask    DEVICE [n_tgt] bytes or NULL: bit j of byte t set = peer j needs row t (a vertex row is needed only by
 *                the ranks whose hyperedge range contains the vertex); NULL = every peer gets every row
 * The caller must order a cross-rank barrier after the kernel before any rank reads the gathered rows.
 * Returns ALLSET_EUNSUPPORTED when the shape is not eligible for the stream kernel (row bytes not in
 * {128,256,512,1024}, too few segments): call allset_segreduce_fwd and all-gather instead. */
int allset_segreduce_fwd_bcast(const void* x, int dtype, int64_t n_src, int32_t d,
                               const int32_t* rowptr, const int32_t* col,
                               const float* w, const float* src_scale,
                               int64_t n_tgt, int op,
                               void* out, void* const* peer_outs, int32_t n_peers, const uint8_t* peer_mask,
                               void* ws, size_t ws_bytes, void* stream);

/* The exchange by itself: send rows [0, n_rows) of this rank's range (already stored in its own replica at `rows`) to the
 * same rows of n_peers (<= 7) peer replicas; peer_rows / peer_mask as in allset_segreduce_fwd_bcast.  For rows whose
 * producer is not one of the fused-exchange kernels (the output of a GEMM + glue pass in a sharded SetGNN layer).
 * row_bytes must be a multiple of 16. */
int allset_push_rows(const void* rows, int64_t n_rows, int64_t row_bytes, void* const* peer_rows, int32_t n_peers,
                     const uint8_t* peer_mask, void* stream);

/* --- dense glue of MLP / PMA ----------------------------------------------------------------
 * out[r, :] = LayerNorm_{gamma,beta,eps}( residual[r, :] + relu( x[r, :] + bias ) ), each stage optional
 * (bias / residual / gamma may be NULL, relu 0|1; beta requires gamma).  fp32, row-major [rows, d].
 * Replaces, in ONE pass over the rows, the bias add of nn.Linear, F.relu and nn.LayerNorm that surround every Linear
 * in MLP.forward (src/layers.py:571-579: Linear -> ReLU -> norm) and PMA's `ln1(out + relu(rFF(out)))`
 * (src/layers.py:155-157).  The GEMMs themselves stay on cuBLAS (the reference hands them to a library too).
 * stats [rows, 2] = (mean, rstd) of the LayerNorm, or NULL (needed only for the backward pass). */
int allset_bias_act_norm(const float* x, const float* bias, int relu, const float* residual,
                         const float* gamma, const float* beta, float eps,
                         int64_t rows, int32_t d, float* out, float* stats, void* stream);

/* Fused two-layer MLP on the tcgen05 tensor cores (bf16 operands, fp32 accumulate in TMEM), eval mode, equal widths
 * in = hid = out = d in {64, 128}:
 *     out[r, :] = [relu]( LN1?( relu( LN0?(x[r, :]) W1^T + b1 ) ) W2^T + b2 )
 * = MLP.forward with num_layers == 2 (src/layers.py:571-579: norm0 -> Linear -> ReLU -> norm -> dropout(eval) -> Linear)
 * plus the F.relu HalfNLHconv wraps around it (src/layers.py:631,634) when relu_out != 0, in ONE pass over the rows.
 * x [rows, d] f32|bf16; w1, w2 [d, d] f32 in nn.Linear layout ([out, in]); b1, b2 [d] f32 or NULL;
 * ln*_gamma NULL = no LayerNorm at that position (Normalization 'None' / InputNorm False); out [rows, d] f32|bf16.
 * Accuracy is that of bf16 operands (the 1e-2 bar of north_star's bf16 mode), not fp32: callers that need 1e-4
 * keep the cuBLAS SGEMM + allset_bias_act_norm chain.  status: device int32 or NULL, set to 1 if an internal
 * mbarrier wait timed out (diagnostic; never in a correct run).  ALLSET_EUNSUPPORTED for other widths.
 * out_pitch: bytes between output rows, 0 = dense (lets PMA.lin_V write the value part of packed [values | scores]
 * records for allset_pma_fwd_strided).
 * w2 == NULL selects ONE Linear: out = [relu]( LN0?(x) W1^T + b1 ) (b2 / ln1 must be NULL) -- nn.Linear as used by
 * PMA.lin_V (src/layers.py:129) and MLP with num_layers == 1. */
int allset_mlp2_fwd(const void* x, int x_dtype,
                    const float* ln0_gamma, const float* ln0_beta, float ln0_eps,
                    const float* w1, const float* b1,
                    const float* ln1_gamma, const float* ln1_beta, float ln1_eps,
                    const float* w2, const float* b2, int relu_out,
                    int64_t rows, int32_t d, void* out, int out_dtype, int64_t out_pitch,
                    int32_t* status, void* stream);

/* PMA's dense tail as ONE tcgen05 kernel (bf16 operands, fp32 accumulate and residual), equal widths d in {64, 128}:
 *     y   = LN0(x)                                   (PMA.ln0 on `out + att_r`, src/layers.py:155)
 *     out = [relu]( LN1( y + relu( rFF(y) ) ) )      (src/layers.py:157; rFF = Linear -> ReLU -> Linear, no norms,
 *                                                     src/layers.py:76-80; the optional ReLU is the one SetGNN.forward
 *                                                     applies to every half layer, src/models.py:475,478)
 * x [rows, d] f32|bf16 (the aggregated rows in their storage dtype); w1, w2 [d, d] f32 ([out, in]); b1, b2, ln*_beta
 * [d] f32 or NULL; out [rows, d] f32|bf16.  The residual y is re-derived from x in fp32 by the epilogue and z = y +
 * relu(.) makes a round trip through TMEM (tcgen05.st / tcgen05.ld) between the statistics pass and the write pass.
 * Replaces: dtype conversion, LayerNorm pass, 2 SGEMMs, 2 glue passes, ReLU pass.  status as in allset_mlp2_fwd. */
int allset_pma_tail_fwd(const void* x, int x_dtype,
                        const float* ln0_gamma, const float* ln0_beta, float ln0_eps,
                        const float* w1, const float* b1, const float* w2, const float* b2,
                        const float* ln1_gamma, const float* ln1_beta, float ln1_eps, int relu_final,
                        int64_t rows, int32_t d, void* out, int out_dtype, int32_t* status, void* stream);

/* PMA's two projections of the source rows in ONE launch (equal widths d in {64, 128}, heads * d <= 1024):
 *     out[r, :]   = x[r, :] W^T + b                      -- V = lin_V(x), src/layers.py:129, on tcgen05 (bf16 operands)
 *     score[r, h] = <x[r, :], w_eff[h, :]> + b_eff[h]    -- (lin_K(x).view(-1, H, C) * att_r).sum(-1), src/layers.py:128,130,
 *                                                           with lin_K folded into w_eff[h, :] = sum_c att_r[h, c] W_K[hC+c, :]
 *                                                           (one seed per head makes the score linear in x); fp32 FMAs
 *                                                           in the producer warps while they hold the row for the GEMM
 * x [rows, d] f32|bf16; W [d, d], w_eff [heads, d], b, b_eff f32; out f32|bf16 with out_pitch as in allset_mlp2_fwd;
 * score [rows, heads] f32 dense.  Replaces a skinny cuBLAS GEMM that re-reads x. */
int allset_linear_score_fwd(const void* x, int x_dtype, const float* w, const float* b,
                            const float* w_eff, const float* b_eff, int32_t heads,
                            int64_t rows, int32_t d, void* out, int out_dtype, int64_t out_pitch,
                            float* score, int32_t* status, void* stream);

/* Arithmetic of the tcgen05 Linear kernels below. */
#define ALLSET_PREC_BF16 0   /* operands rounded to bf16, fp32 accumulate (1e-2 class) */
#define ALLSET_PREC_SPLIT 1  /* f32 operands cut into bf16 terms whose products are accumulated in fp32 in TMEM: THREE terms and the
                              * six products down to 2^-18 for allset_linear_fwd (as close to fp64 as an fp32 SGEMM: 1e-6 of the
                              * output scale; two terms leave 2^-17 per product, which flips ReLUs that sit within 1e-5 of zero),
                              * TWO terms / three products for allset_linear_wgrad (a sum over all rows averages the 2^-17 errors) */

/* ONE nn.Linear on tcgen05 (equal widths d in {64, 128}), the unit the training path is built from:
 *     out = [relu]( LN?(x) op(W)^T + b ),   op(W) = W ([out, in], nn.Linear layout)      -- forward of MLP.lins[i] /
 *                                                   PMA.lin_V (src/layers.py:575,129)
 *                                           op(W) = W^T when w_transposed != 0             -- its input gradient dx = dy W
 * x [rows, d] f32|bf16; W [d, d] f32; b, ln_* [d] f32 or NULL (the LayerNorm is the one IN FRONT of the Linear:
 * MLP.normalizations[i], src/layers.py:573,577; not with w_transposed); out [rows, d] dense.
 * precision ALLSET_PREC_BF16: any dtype pair; ALLSET_PREC_SPLIT: f32 in, f32 out, out 32-byte aligned -- this is what
 * lets the DEFAULT (reference-precision) mode leave the SIMT SGEMM.  status as in allset_mlp2_fwd. */
int allset_linear_fwd(const void* x, int x_dtype, const float* ln_gamma, const float* ln_beta, float ln_eps,
                      const float* w, int w_transposed, const float* b, int relu, int precision,
                      int64_t rows, int32_t d, void* out, int out_dtype, int32_t* status, void* stream);

/* Weight gradient of the same Linear: dw[n, k] = sum_r dy[r, n] * x[r, k]  (dw [d, d] f32, nn.Linear layout).
 * dy, x [rows, d], both `dtype`, 32-byte aligned; f32 rows use ALLSET_PREC_SPLIT, bf16 rows ALLSET_PREC_BF16.
 * One pass over both inputs; every CTA keeps its share of the sum in TMEM and writes one [d, d] partial into
 * `workspace` (allset_linear_wgrad_partials(rows) * d * d floats, 32-byte aligned), a second launch adds the partials
 * in a fixed order: deterministic, no atomics. */
int allset_linear_wgrad_partials(int64_t rows);
int allset_linear_wgrad(const void* dy, const void* x, int dtype, int precision, int64_t rows, int32_t d,
                        float* dw, float* workspace, int64_t workspace_floats, int32_t* status, void* stream);

/* Backward of allset_bias_act_norm (d in {128,256,512,1024}; ALLSET_EUNSUPPORTED otherwise):
 *   dx [rows, d] = gradient w.r.t. x;  dres [rows, d] or NULL = gradient w.r.t. residual;
 *   partial [blocks, 3, d] = per-CTA column sums of (d gamma, d beta, d bias), blocks =
 *   allset_bias_act_norm_bwd_blocks(rows); the caller sums over dim 0 (deterministic, no atomics). */
int32_t allset_bias_act_norm_bwd_blocks(int64_t rows);
int allset_bias_act_norm_bwd(const float* dy, const float* x, const float* bias, int relu,
                             const float* residual, const float* gamma, const float* stats,
                             int64_t rows, int32_t d, float* dx, float* dres, float* partial, void* stream);

/* --- rowop: the same glue with mixed-precision rows, dropout and a backward for every width it takes -------------
 *   z = residual + relu(x + bias);  y = LayerNorm(z);  out = dropout_p(relu_out(y))     (every stage optional)
 * = what sits between the Linears of MLP.forward in TRAINING (src/layers.py:571-579: Linear -> ReLU -> norm -> dropout)
 * and around PMA's rFF (src/layers.py:153-157), one pass forward and one pass backward, so that the Linears themselves
 * can run as bf16 tensor-core GEMMs with bf16 activations in between.
 *   x, residual [rows, d] x_dtype;  out [rows, d] out_dtype;  bias, gamma, beta [d] f32;  stats [rows, 2] f32 or NULL
 *   d in {64, 128, 256, 512, 1024} (allset_rowop_supported), rows 16-byte aligned; ALLSET_EUNSUPPORTED otherwise.
 *   drop_p in [0, 1): element (r, c) is kept iff a 16-bit slice of a counter hash of (seed, r, c) >= p * 65536 and
 *   scaled by 1 / (1 - p) -- F.dropout's distribution, regenerated (not stored) by the backward pass from the same seed.
 * Backward: dy [rows, d] g_dtype -> dx (and dres, or NULL) in g_dtype; partial [blocks, 3, d] f32 = per-CTA column sums
 * of (d gamma, d beta, d bias), blocks = allset_bias_act_norm_bwd_blocks(rows), summed over dim 0 by the caller. */
int allset_rowop_supported(int32_t d);
int allset_rowop_fwd(const void* x, int x_dtype, const float* bias, int relu, const void* residual,
                     const float* gamma, const float* beta, float eps, int relu_out, float drop_p, uint64_t seed,
                     int64_t rows, int32_t d, void* out, int out_dtype, float* stats, void* stream);
int allset_rowop_bwd(const void* dy, int g_dtype, const void* x, int x_dtype, const float* bias, int relu,
                     const void* residual, const float* gamma, const float* beta, const float* stats, int relu_out,
                     float drop_p, uint64_t seed, int64_t rows, int32_t d, void* dx, void* dres, float* partial,
                     void* stream);

/* Gradient w.r.t. the per-incidence weights (SetGNN.LearnMask, src/models.py:451-452):
 * grad_w[k] = tgt_scale[t] * <x[col[k], :], grad_out[t, :]>  for k in segment t (CSR order).
 * tgt_scale [n_tgt] fp32 or NULL (mean: 1/max(count,1)). */
int allset_segreduce_bwd_w(const void* x, const void* grad_out, int dtype, int32_t d,
                           const int32_t* rowptr, const int32_t* col, const float* tgt_scale,
                           int64_t n_tgt, float* grad_w, void* stream);

/* --- AllSetTransformer: per-segment multi-head PMA -----------------------------------------
 * One seed query per head, so per target segment t and head h:
 *   a_k   = leaky_relu(score[col[k], h], slope)
 *   alpha = exp(a_k - max_k a_k) / (sum_k exp(a_k - max) + 1e-16)      (PyG softmax)
 *   out[t, h, :] = sum_k alpha_k * v[col[k], h, :] + seed[h, :]
 * Replaces PMA.propagate/message/aggregate + the seed residual (src/layers.py:145-153,168-194):
 * index_select x2, leaky_relu, the 6-kernel segment softmax, the weighting and the scatter-add.
 *   v     [n_src, H*C] (dtype)     score [n_src, H] fp32     seed [H*C] fp32
 *   out   [n_tgt, H*C] (dtype)     stats [n_tgt, H, 2] fp32 = (max, sum + 1e-16), or NULL
 * Empty segments yield out = seed (the reference's zero row + att_r). */
int allset_pma_fwd(const void* v, const float* score, const float* seed, int dtype,
                   int32_t H, int32_t C, float slope,
                   const int32_t* rowptr, const int32_t* col, int64_t n_tgt,
                   const int32_t* long_ids, int32_t n_long, int32_t long_threshold,
                   void* out, float* stats, void* ws, size_t ws_bytes, void* stream);

/* allset_pma_fwd with STRIDED sources: value row i at (char*)v + i * v_pitch, its H fp32 scores at (char*)score +
 * i * s_pitch (both pitches in bytes, multiples of 16).  With v_pitch == s_pitch and score == v + H*C*sizeof(T) the
 * source is ONE packed record per row [values | scores]: a gather then touches one contiguous 16-byte-granular record
 * per incidence instead of a row plus a 32-byte score in another 128-byte line (the V->E direction of a large graph is
 * DRAM-bound on exactly that).  Stream kernel only: ALLSET_EUNSUPPORTED when allset_stream_eligible() says no or the
 * graph has long segments; out is dense [n_tgt, H*C]. */
int allset_pma_fwd_strided(const void* v, int64_t v_pitch, const float* score, int64_t s_pitch, const float* seed,
                           int dtype, int32_t H, int32_t C, float slope,
                           const int32_t* rowptr, const int32_t* col, int64_t n_tgt,
                           void* out, float* stats, void* ws, size_t ws_bytes, void* stream);

/* allset_pma_fwd with the fused exchange of allset_segreduce_fwd_bcast (same contract for out / peer_outs). */
int allset_pma_fwd_bcast(const void* v, const float* score, const float* seed, int dtype,
                         int32_t H, int32_t C, float slope,
                         const int32_t* rowptr, const int32_t* col, int64_t n_tgt,
                         void* out, float* stats, void* const* peer_outs, int32_t n_peers,
                         const uint8_t* peer_mask, void* ws, size_t ws_bytes, void* stream);

/* Attention weights per incidence in CSR order (PMA.forward(return_attention_weights=True),
 * src/layers.py:159-166): alpha[k, h] for k in segment t from score and stats. */
int allset_pma_alpha(const float* score, const float* stats, int32_t H, float slope,
                     const int32_t* rowptr, const int32_t* col, int64_t n_tgt,
                     float* alpha, void* stream);

/* Per-row, per-head dot product  out[r, h] = sum_c a[r,h,c] * (b[r,h,c] - sub[h,c])
 * (sub may be NULL).  Used for the softmax-backward term D = <grad_out, out - seed>. */
int allset_rowdot_heads(const void* a, const void* b, const float* sub, int dtype,
                        int64_t n_rows, int32_t H, int32_t C, float* out, void* stream);

/* Backward of allset_pma_fwd, run over the TRANSPOSED CSR (segments = source rows s,
 * colT[k] = target row of the incidence):
 *   alpha_k       = exp(leaky_relu(score[s,h]) - stats[t,h,0]) / stats[t,h,1]
 *   grad_v[s,h,:] = sum_k alpha_k * grad_out[t_k, h, :]
 *   grad_score[s,h] = leaky_relu'(score[s,h]) * (<grad_v[s,h,:], v[s,h,:]> - sum_k alpha_k * D[t_k,h])
 * D [n_tgt, H] from allset_rowdot_heads(grad_out, out, seed).  (grad_seed = column sums of
 * grad_out is left to the host framework.) */
int allset_pma_bwd(const void* grad_out, const void* v, const float* score, const float* stats,
                   const float* D, int dtype, int32_t H, int32_t C, float slope,
                   const int32_t* rowptrT, const int32_t* colT, int64_t n_src,
                   const int32_t* long_ids, int32_t n_long, int32_t long_threshold,
                   void* grad_v, float* grad_score, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALLSET_B200_H_ */
