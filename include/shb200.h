/*
 * shb200.h -- C ABI of libshb200.so: the B200 (sm_100a) kernels behind SemanticHuman's
 * spiral-mesh-autoencoder training step.
 *
 * The reference (XiaokunSun/SemanticHuman) has no FFI of its own: its operator interface for this path is the
 * module-level Python class API of models.py, and underneath it ATen.  Each entry point below replaces one
 * group of ATen call sites; the reference line(s) it stands in for are cited per function.  A reference
 * maintainer would bind these with ctypes (INTEGRATION.md shows the stub) -- exactly what
 * semantichuman_b200/_capi.py does.
 *
 * Conventions
 *   - plain pointers + sizes only; device pointers unless a parameter says HOST.
 *   - no allocation, no global state, no synchronisation inside; every launch goes to `stream`
 *     (a cudaStream_t passed as void*).  The caller owns every buffer, including workspaces.
 *   - return value: 0 = ok; negative = argument error (SHB_E_*); positive = a cudaError_t from the launch.
 *     shb_error_string() renders either.  The Python host raises RuntimeError on any non-zero code.
 *   - callers hand over row-major (B, rows, C) tensors (batch-major, channel fastest -- the reference's own layout,
 *     models.py:37,48); inside the trunks activations live in the SLAB layout described further down, and
 *     shb_slab_from_rows / shb_slab_to_rows convert.  Accumulation is always fp32.
 *   - "table": the spiral index table restricted to the rows this call produces, int32, (rows_out, S),
 *     values in [0, rows_in): the reference's -1 (models.py:42, Python negative index -> dummy row) is
 *     normalised to rows_in-1 on the host (semantichuman_b200/indexing.py).
 */
#ifndef SHB200_H
#define SHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHB_ABI_VERSION 2

/* activation enum == the strings models.py:19-32 accepts */
enum { SHB_ACT_IDENTITY = 0, SHB_ACT_RELU = 1, SHB_ACT_ELU = 2, SHB_ACT_LEAKY_RELU = 3, SHB_ACT_SIGMOID = 4,
       SHB_ACT_TANH = 5 };
/* storage dtype */
enum { SHB_F32 = 0, SHB_BF16 = 1 };
/* argument errors */
enum { SHB_E_ARG = -1, SHB_E_DTYPE = -2, SHB_E_SHAPE = -3, SHB_E_WORKSPACE = -4, SHB_E_UNSUPPORTED = -5 };

int shb_abi_version(void);
/* SMs the persistent kernels (SpiralConv forward / input gradient / weight gradient: one CTA per SM) launch on, 1..148
 * (default 148).  A data-parallel step sets 148 - k so that the NCCL all-reduce of the gradient buckets, limited to k CTAs,
 * overlaps the rest of the backward instead of queueing behind it.  Process-wide setting; not a per-call argument. */
int shb_set_persistent_sms(int n);
const char* shb_error_string(int code);

/* ------------------------------------------------------------------ host-side index construction (CPU) */

/* Inverse-spiral CSR (replaces the index_put_(accumulate=True) that autograd runs for
 * models.py:42).  For every source row u in [0, rows_in): the flat slot positions j*S+s with table[j,s]==u,
 * ascending -- i.e. a stable counting sort of the flattened table.
 *   HOST table (rows_out*S), HOST rowptr (rows_in+1), HOST slots (rows_out*S). */
int shb_build_inverse_spiral_csr(const int32_t* table, int rows_out, int S, int rows_in, int32_t* rowptr,
                                 int32_t* slots);

/* Dense padded sampling matrix (main.py:183-193: D/U .todense(), +1 row/col, corner 1) -> CSR, dropping exact
 * zeros.  Call with colidx==NULL to count: *nnz_out receives the number of non-zeros.
 *   HOST dense (rows*cols) fp32, HOST rowptr (rows+1), HOST colidx/vals (cap). */
int shb_dense_to_csr(const float* dense, int rows, int cols, int32_t* rowptr, int32_t* colidx, float* vals,
                     int64_t cap, int64_t* nnz_out);

/* CSR (rows x cols) -> CSR of the transpose (cols x rows), entries of each output row in ascending column
 * order: the operand of the Pool backward  dx = P^T dy  (autograd of models.py:127,148). HOST pointers. */
int shb_csr_transpose(const int32_t* rowptr, const int32_t* colidx, const float* vals, int rows, int cols,
                      int32_t* t_rowptr, int32_t* t_colidx, float* t_vals);

/* ------------------------------------------------------------------ losses (train_funcs.py:135,145-152,501) */

/* loss = mean |a - b| over n elements (F.l1_loss, main.py:296,311).  Two-stage fixed-order reduction.
 *   partials: fp32 workspace of shb_l1_loss_workspace(n) bytes; loss_out: device fp32 scalar. */
size_t shb_l1_loss_workspace(int64_t n);
int shb_l1_loss_fwd(const void* a, const void* b, int64_t n, void* partials, size_t partials_bytes, float* loss_out,
                    int dtype, void* stream);
/* ga = gscale * sign(a-b)/n, gb = -ga; either may be NULL.  gscale: device fp32 scalar (upstream grad). */
int shb_l1_loss_bwd(const void* a, const void* b, int64_t n, const float* gscale, void* ga, void* gb, int dtype,
                    void* stream);

/* out[c] = sum over b of x[b][c]: the bias gradient of the two latent nn.Linear layers (models.py:129,142; autograd's
 * sum over the batch), x row-major (B, N) of `dtype`, out fp32 (N).  Fixed summation order (16-byte vector loads when N is a
 * multiple of 8, one column per thread otherwise). */
int shb_colsum(const void* x, int dtype, int B, int N, float* out, void* stream);

/* Part-measure latent loss (train_funcs.py:145-152): m[b,p] = ||z[b,p,:]||_2;
 *   relative != 0:  loss = mean_{b,i} | m[b,P[i]] / measure[b,Q[i]] - 1 |
 *   relative == 0:  loss = mean_{b,i} | m[b,P[i]] - measure[b,Q[i]] |
 * Writes the loss and d loss / d z (unscaled by the upstream grad) in one pass; fp32 only.
 *   z (B, n_parts, L); measure (B, n_measure); P,Q int32 (n_sel) device; gz (B, n_parts, L). */
int shb_partnorm_loss_fwd_bwd(const float* z, const float* measure, const int32_t* P, const int32_t* Q, float* loss_out,
                              float* gz, int B, int n_parts, int L, int n_measure, int n_sel, int relative,
                              void* stream);

/* ---- Bone-guided heads as grouped kernels (models.py:200-204, 233-236, 252-253, 269-273): the 17+17+17 per-part
 * nn.Linear layers, their fancy-index gathers, the permutation scatter of models.py:270-272 and the biases, one launch
 * per direction.  Group k owns rows idx[gptr[k] .. gptr[k+1]) of a (B, rows, C) fp32 tensor (n_k rows); its weight and
 * bias start at w + woff[k] / bias + boff[k] of packed fp32 buffers in nn.Linear layout.  idx, gptr int32 and woff, boff
 * int64 are device arrays.  fp32 only; L, Lin <= 32; B <= 65535.  Fixed-order reductions (no float atomics).
 *
 * gather form  (fc_latent_enc_list / kps_enc_list):  z[b,k,o] = bias_k[o] + sum_{p,c} W_k[o, p*C+c] * x[b, idx[gptr[k]+p], c]
 *   W_k (L, n_k*C).  bwd: gx (B, rows, C) -- rows outside every group are zeroed; groups must not overlap when gx is
 *   requested -- gw (packed like w), gb (packed like bias); each may be NULL.  max_group_rows = max_k n_k.
 * scatter form (fc_latent_dec_list + models.py:270-272):  y[b, idx[gptr[k]+p], c] = bias_k[p*C+c] + sum_i W_k[p*C+c, i] * zz[b,k,i]
 *   W_k (n_k*C, Lin); rows of y outside every group are left untouched.  bwd: gzz (B, G, Lin), gw, gb; each may be NULL. */
int shb_group_linear_gather_fwd(const float* x, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                const float* bias, const int64_t* boff, float* z, int B, int rows, int C, int G, int L,
                                void* stream);
int shb_group_linear_gather_bwd(const float* x, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                const int64_t* boff, const float* gz, float* gx, float* gw, float* gb, int B, int rows, int C,
                                int G, int L, int max_group_rows, void* stream);
int shb_group_linear_scatter_fwd(const float* zz, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                 const float* bias, const int64_t* boff, float* y, int B, int rows, int C, int G, int Lin,
                                 void* stream);
int shb_group_linear_scatter_bwd(const float* zz, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                 const int64_t* boff, const float* gy, float* gzz, float* gw, float* gb, int B, int rows, int C,
                                 int G, int Lin, int max_group_rows, void* stream);

/* ---- Orientation-adaptive pairwise-distance loss (utils_SH.py:442-478 angle_skl; utils_distance.py:366-376;
 * train_funcs.py:243-284 == :353-389), fused: per part k (vertex ids idx[gptr[k] .. gptr[k+1]) of the (B, V, 3) fp32
 * ground truth tx and reconstruction rec), over the ordered pairs i != j with w*De != 0,
 *     relative != 0:  mean | w*De_r/De - w |        relative == 0:  mean | w*De_r - w*De |
 * De / De_r = pairwise distances in tx / rec (De times scale[b,k] when scale != NULL); w from the angle (degrees) between
 * v_i - v_j and the part's bone kps[bone[k][0]] - kps[bone[k][1]] (or minus the mean of two keypoints when bone[k][2] >= 0):
 * wmode[k] = 0 all-one, 1 angle/90, 2 sin(angle), 3 angle/90 zeroed below w_threshold.  loss = sum_k part_weight[k] * mean_k.
 * fwd writes the loss and keeps the per-part normalisers in the workspace (parts must not overlap).  Fixed-order reductions.
 * kps (B, NK, 3); idx, gptr, bone (G,3), wmode (G) int32 and part_weight (G), scale (B,G) fp32 are device arrays. */
size_t shb_pair_loss_workspace(int B, int G, int max_part_rows);
/* grad_acc: a buffer of shb_pair_loss_grad_acc_bytes(B, G, max_part_rows) bytes, or NULL.  When given, the forward pass also
 * leaves the unscaled gradient there -- every unordered vertex pair is evaluated once and credited to both of its vertices,
 * per 128-vertex tile pair -- and shb_pair_loss_bwd (same idx / gptr / workspace) only sums a vertex's tile-pair slots in a
 * fixed order and scales: grec = 2 * coef[part] * gscale[0] * sum, zero for vertices outside every part.  One walk over HALF
 * the pairs serves loss and gradient. */
size_t shb_pair_loss_grad_acc_bytes(int B, int G, int max_part_rows);
int shb_pair_loss_fwd(const float* tx, const float* rec, const float* kps, const int32_t* idx, const int32_t* gptr,
                      const int32_t* bone, const int32_t* wmode, const float* part_weight, const float* scale,
                      float w_threshold, int relative, float* loss_out, float* grad_acc, void* workspace, size_t workspace_bytes,
                      int B, int V, int NK, int G, int max_part_rows, void* stream);
int shb_pair_loss_bwd(const float* grad_acc, const int32_t* idx, const int32_t* gptr, const float* gscale, float* grec,
                      const void* workspace, size_t workspace_bytes, int B, int V, int G, int max_part_rows, void* stream);

/* ==================================================================================================================
 * Slab layout: the trunks' internal activation format (batch innermost, 128-sample chunks, 8-channel planes).
 *
 *   element (r, b, c) of plane p  ->  bf16 index ((((r*NB + b/128)*P + p)*(C/8) + c/8)*128 + b%128)*8 + c%8
 *   NB = ceil(B/128) (tail chunk zero-padded), C % 8 == 0, P = planes: 1 = bf16 mode, 2 = fp32 mode (bf16 hi + bf16 lo).
 *
 * One (row, chunk) pair -- a slab -- is P*C*256 contiguous bytes; what SpiralConv gathers for an output vertex
 * (x[:, spiral_idx] of models.py:42) is S whole slabs, each moved by one cp.async.bulk (TMA) straight into the UMMA
 * operand layout.  `rows` always counts the dummy vertex.  The entry points below replace, in this layout, the same
 * reference call sites as their row-major counterparts above.
 * ================================================================================================================== */
size_t shb_slab_tensor_bytes(int rows, int B, int C, int planes);

/* rows <-> slabs (main.py feeds (B, V+1, 3) fp32 tensors; models.py:129,142 reshape to/from the FC layers' (B, rows*C)).
 * from_rows: src row-major (B, R, Cs) of `src_dtype`; internal row i takes the caller's row perm[i] (perm NULL: identity);
 *   channels Cs..Cp-1 and samples B..128*NB-1 are written as zeros; if ymul != NULL (slab tensor shaped like dst) the result
 *   is multiplied by act'(ymul) -- the activation derivative expressed through the layer OUTPUT (act_mul = SHB_ACT_*);
 *   zero_last zeroes row R-1 (the mask of models.py:48-51 on the gradient path).
 * to_rows: dst row-major (B, R, Cd), Cd <= Cp; the caller's row perm[i] receives internal row i.
 * perm_inv (optional; NULL allowed): the inverse permutation, perm_inv[perm[i]] = i.  With it the 8-channel (3-channel ends)
 *   conversions walk the row-major tensor in the caller's order -- contiguous reads / writes -- and scatter on the slab side,
 *   where a row is 2 KB; without it they fall back to the item-per-thread kernels. */
int shb_slab_from_rows(const void* src, int src_dtype, const int32_t* perm, const int32_t* perm_inv, void* dst, const void* ymul,
                       int B, int R, int Cs, int Cp, int act_mul, int zero_last, int planes, void* stream);
int shb_slab_to_rows(const void* src, const int32_t* perm, const int32_t* perm_inv, void* dst, int dst_dtype, int B, int R, int Cp,
                     int Cd, int planes, void* stream);

/* L1 reconstruction loss straight from the slab tensor of the last decoder SpiralConv (train_funcs.py:501 / main.py:296
 * F.l1_loss on the output of models.py:159): rec = slab tensor (R, B, 8 channels padded, Cs <= 8 real), target = row-major
 * (B, R, Cs) of `target_dtype` in the CALLER's row order, perm_inv[c] = internal row of the caller's row c (NULL: identity).
 *   fwd: *loss_out = mean over B*R*Cs elements of |rec - target|; fixed-order two-stage reduction; partials = fp32 workspace
 *        of shb_slab_l1_workspace() bytes.
 *   bwd: grad_slab (shaped like rec) = *gscale / (B*R*Cs) * sign(rec - target), multiplied by act'(rec) (act_mul, the
 *        producer's activation, derivative through its output) and with row R-1 zeroed when zero_last (the producer's
 *        dummy-row mask, models.py:48-51) -- i.e. already in the trunk's gradient convention; padded channels / samples zero.
 * Replaces slab_to_rows + l1_loss_fwd + l1_loss_bwd + slab_from_rows of a training step with two passes over the slab. */
size_t shb_slab_l1_workspace(void);
int shb_slab_l1_fwd(const void* rec, const void* target, int target_dtype, const int32_t* perm_inv, void* partials,
                    size_t partials_bytes, float* loss_out, int B, int R, int Cs, int planes, void* stream);
int shb_slab_l1_bwd(const void* rec, const void* target, int target_dtype, const int32_t* perm_inv, const float* gscale,
                    void* grad_slab, int B, int R, int Cs, int act_mul, int zero_last, int planes, void* stream);

/* Pool (models.py:127,148: torch.matmul(D[i] | U[i], x)) on slab tensors: dst[r] = sum_k vals[k] * src[colidx[k]],
 * k in rowptr[r]..rowptr[r+1].  Optional epilogue for the gradient path: times act'(ymul[r]) and/or zero the last row. */
int shb_slab_pool(const void* src, const int32_t* rowptr, const int32_t* colidx, const float* vals, void* dst, const void* ymul,
                  int B, int rows_out, int C, int act_mul, int zero_last, int planes, void* stream);

/* Weight operand images of one SpiralConv (nn.Linear weight (Cout, S*Cin) fp32, models.py:16): bf16, un-swizzled UMMA
 * K-major core-matrix order [plane][N/8][K/8][8][8]; forward image N = pad16(Cout_p), K = S*Cin_p; backward image (per-slot
 * transpose) N = pad16(Cin_p), K = S*Cout_p.  Cin_p / Cout_p: channel counts of the slab tensors (multiples of 8). */
size_t shb_slab_weight_image_bytes(int S, int Ck, int Cn, int planes);
int shb_slab_weight_images(const float* w, void* img_fwd, void* img_bwd, int S, int Cin, int Cout, int Cin_p, int Cout_p,
                           int planes, void* stream);
/* The same for `count` layers in one launch (HOST arrays of device pointers / dimensions): a model refreshes the images of
 * all its conv layers at once after an optimizer step. */
int shb_slab_weight_images_batch(int count, const float* const* w, void* const* img_fwd, void* const* img_bwd, const int* S,
                                 const int* Cin, const int* Cout, const int* Cin_p, const int* Cout_p, int planes, void* stream);

/* SpiralConv forward / input gradient (models.py:34-53 and the index_put_(accumulate) + mm of its autograd):
 *   dst[u] = mask(u) * act'(ymul[u]) * act( bias + sum_{e in ptr[u]..ptr[u+1]} src[entries[e] >> 5] . Wimg[slot entries[e] & 31] )
 * forward: entries of output row j = (table[j,s] << 5 | s), w_img = forward image, act = the layer's, ymul NULL;
 * input gradient: entries of source row u = (j << 5 | s) over all (j,s) with table[j,s] == u (fixed order: the sum is
 * accumulated in list order in TMEM, no atomics), w_img = backward image, bias NULL, act identity, ymul/act_mul = the
 * PRODUCER layer's output and activation (NULL: none), zero_last = producer's mask.
 * src (.., B, Cs) and dst (rows_dst, B, Cd) slab tensors; Cd_real = channels `bias` holds. */
int shb_slab_conv_supported(int S, int Cs, int Cd, int planes);
int shb_slab_conv(const void* src, const int32_t* ptr, const int32_t* entries, const void* w_img, const float* bias, void* dst,
                  const void* ymul, int B, int rows_dst, int S, int Cs, int Cd, int Cd_real, int act, int act_mul, int zero_last,
                  int planes, void* stream);

/* SpiralConv weight / bias gradient (the two mm of models.py:45's autograd), K = batch on tcgen05:
 *   gw[o][s*Cin + c] = sum_{j,b} x[table[j,s]][b][c] * gz[j][b][o],   gb[o] = sum_{j,b} gz[j][b][o]   (gb may be NULL)
 * x (rows_in, B, Cin_p), gz (rows_out, B, Cout_p) slab tensors; per-CTA partials in `workspace`, added in CTA order.
 * zero_src_row: a row of x known to be all zero (the dummy vertex behind padded spiral slots, when its producer masked it), or
 * -1; trailing 128-channel operand groups of a vertex that read nothing else are neither loaded nor multiplied. */
int shb_slab_wgrad_supported(int S, int Cin_p, int Cout_p, int planes);
size_t shb_slab_wgrad_workspace(int S, int Cin_p, int Cout_p, int planes);
int shb_slab_wgrad(const void* x, const int32_t* table, const void* gz, float* gw, float* gb, void* workspace,
                   size_t workspace_bytes, int B, int rows_out, int S, int Cin, int Cin_p, int Cout, int Cout_p, int skip_last,
                   int zero_src_row, int planes, void* stream);

/* ---- Shared-source groups of a SpiralConv pass (host, CPU).  `ptr` (rows_dst+1) / `ent` ((source row << 5) | slot) are the
 * entry lists of a pass (forward: models.py:42's x[:, spiral_idx] per output row; input gradient: the inverse-spiral lists).
 * Destination rows whose lists share sources are clustered (greedy, deterministic) into groups of <= R rows; per group the
 * distinct sources are listed once, in ascending order, each with the (destination, slot) pairs that use it.  Output is the
 * program shb_slab_gconv runs:
 *   gptr  (n_groups+1)  record range of every group
 *   recs  (n_records*48) records of <= SPS sources and <= 32 pairs:  [0] = n_src | n_pairs << 4 | first << 10 | last << 11,
 *         [1..8] source rows, [16..47] pairs = slab-in-record | dest-in-group << 3 | first-pair-of-dest << 8 | slot << 9
 *   gdst  (n_groups*R)  destination rows (-1: absent),  gmask (n_groups) bit i: destination i has no entries.
 * Two calls: with gptr == recs == gdst == gmask == NULL it only returns the counts; the second call fills the arrays. */
int shb_build_conv_groups(const int32_t* ptr, const int32_t* ent, int rows_dst, int rows_src, int R, int SPS, int32_t* n_groups,
                          int32_t* n_records, int32_t* gptr, int32_t* recs, int32_t* gdst, uint32_t* gmask);

/* ---- SpiralConv forward / input gradient with shared-source groups (shb_slab_gconv.cu): same contract as shb_slab_conv,
 * driven by the group program of shb_build_conv_groups instead of per-row entry lists.  A tile is one group x one 128-sample
 * chunk; every distinct source slab of the group is loaded once and feeds all (destination, slot) pairs that use it, each
 * destination accumulating in its own TMEM columns.  shb_slab_gconv_plan reports the largest group size R and the slabs per
 * record SPS the kernel's ring takes for a layer shape (SHB_E_UNSUPPORTED: use shb_slab_conv); the program passed to
 * shb_slab_gconv must have been built with that SPS and R <= R_max. */
int shb_slab_gconv_plan(int S, int Cs, int Cd, int planes, int* R_max, int* SPS);
int shb_slab_gconv(const void* src, const int32_t* gptr, const int32_t* recs, const int32_t* gdst, const uint32_t* gmask,
                   int n_groups, int R, int SPS, const void* w_img, const float* bias, void* dst, const void* ymul, int B,
                   int rows_dst, int S, int Cs, int Cd, int Cd_real, int act, int act_mul, int zero_last, int planes,
                   void* stream);

/* ---- Optimizer step (main.py:262: torch.optim.Adam(params, lr, weight_decay), defaults betas (0.9, 0.999), eps 1e-8):
 * g += weight_decay * p;  m = lerp(m, g, 1-beta1);  v = beta2 v + (1-beta2) g^2;
 * p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps),   t = *step (device scalar, advanced by shb_adam_tick
 * BEFORE the step so that the pair replays inside a CUDA graph).  `count` tensors, HOST arrays of DEVICE pointers (16-byte
 * aligned) and element counts; shadow[k] != NULL receives bf16(p) of the updated weights (the operand of the bf16-mode FC
 * GEMMs), so the step contains no separate weight cast.  shb_cast_bf16 is that cast on its own (initialisation, or after an
 * update by a foreign optimizer). */
int shb_adam_tick(float* step, void* stream);
int shb_adam_step(int count, float* const* p, const float* const* g, float* const* m, float* const* v, void* const* shadow,
                  const int64_t* numel, const float* step, double lr, double beta1, double beta2, float eps, float weight_decay,
                  void* stream);
/* Same step with per-tensor gradient dtypes: g_is_bf16[k] != 0 (HOST array, or NULL = all fp32) marks g[k] as a bf16 tensor
 * (a data-parallel gradient bucket that was reduced in bf16; 8-byte aligned). */
int shb_adam_step_mixed(int count, float* const* p, const void* const* g, const uint8_t* g_is_bf16, float* const* m,
                        float* const* v, void* const* shadow, const int64_t* numel, const float* step, double lr, double beta1,
                        double beta2, float eps, float weight_decay, void* stream);
int shb_cast_bf16(const float* src, void* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHB200_H */
