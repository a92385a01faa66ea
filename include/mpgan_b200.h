/* mpgan_b200 -- C ABI of the B200-native MPGAN / GAPT message-passing hot path.
 *
 * Drop-in boundary: these are the entry points a maintainer of rkansal47/MPGAN binds (ctypes /
 * torch custom op) to replace the aten op chains below `MPLayer.forward`
 * (mpgan/model.py:206-282), `LinearNet.forward` (mpgan/model.py:70-85), the generator /
 * discriminator mask + tail code (mpgan/model.py:689-699, 723-752, 810-831, 881-884),
 * `SpectralNorm._update_u_v` (mpgan/spectral_normalization.py:21-33) and GAPT's `MAB.forward`
 * (gapt/model.py:124-139).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 data unless stated otherwise; tensors are row-major;
 *    `ld*` arguments are row strides in elements, so strided views (e.g. x[:, :, :-1]) need no copy
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate, keep no
 *    global state and are safe under CUDA-graph capture; scratch memory comes from the caller
 *  - gradient outputs named d<param> are ACCUMULATED into (+=), matching autograd's .grad semantics;
 *    gradient outputs w.r.t. activations (dx) are overwritten unless stated otherwise
 *  - return value: 0 = ok; non-zero = error, message via mpg_last_error() (thread-local)
 *  - there is no CPU fallback: a missing GPU / unsupported option is an error, never a silent detour
 *  - dropout is a counter-based Philox4x32-10 stream keyed by (seed, rng_stream, row, column), so
 *    the backward entry points regenerate exactly the mask the forward call drew; the effective
 *    seed is `seed + *seed_dev` when the optional device pointer seed_dev is non-NULL, so a
 *    captured CUDA graph draws fresh masks on every replay by bumping one device word
 *  - precision: 0 = fp32-class (3xTF32 node GEMMs, fp32 SIMT edge network), 1 = fast
 *    (TF32 node GEMMs; bf16 tcgen05 edge network with fp32 accumulation when the edge network is
 *    the default 96/160/192 architecture, else the fp32 SIMT kernel)
 */
#ifndef MPGAN_B200_H_
#define MPGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPG_ACT_NONE 0
#define MPG_ACT_LRELU 1   /* LinearNet hidden layers */
#define MPG_ACT_TANH 1    /* mpg_unary / mpg_gen_tail: 1 = tanh, 2 = sigmoid */
#define MPG_ACT_SIGMOID 2

int mpg_version(void);
const char* mpg_last_error(void);
/* compiled feature flags: bit0 = tcgen05 edge forward, bit1 = tcgen05 edge backward, bit2 = GAPT */
int mpg_features(void);
/* number of kernel launches this library has issued so far in this process */
unsigned long long mpg_launch_count(void);
/* measurement hook: brackets the NEXT launch of one tcgen05 edge kernel (1 = forward, 2 = backward
 * chain, 3 = backward dW2) with the two cudaEvent_t handles; one-shot, calling thread only */
void mpg_probe(int kernel_id, void* ev_start, void* ev_stop);

/* ---- LinearNet layer (mpgan/model.py:77-83): y = dropout(act(x W^T + b)) -------------------------- */
int mpg_linear_fwd(const float* x, int ldx, const float* w, const float* b, float* y, int M, int K, int N,
                   int act, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t rng_stream,
                   int precision, void* stream);
/* given dy and the layer output y: dz = dy * d(act,dropout)(y) into `dz` (scratch [M,N], may alias dy);
 * dx = dz W (skipped if dx == NULL; accumulated if dx_accumulate); dw += dz^T x; db += colsum(dz). */
int mpg_linear_bwd(const float* dy, const float* y, const float* x, int ldx, const float* w, float* dz,
                   float* dx, int lddx, int dx_accumulate, float* dw, float* db, int M, int K, int N, int act,
                   float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t rng_stream,
                   int precision, void* stream);

/* ---- fused edge network + neighbour aggregation (mpgan/model.py:256-267, 284-317) -----------------
 * agg[b,i,:] = scale * sum_j mask[b,j] * fe(x_i | x_j | ef_ij),  scale = 1 (sum) or 1/N (mean).
 * W0 is fe.net.0.weight [H0, 2F + n_ef] with column blocks (receiver | sender | [diffs] | [dist]).
 * ef_mode: bit0 = Euclidean distance column, bit1 = difference columns; nd = number of leading
 * features the differences are taken over.  The N^2 x H tensors never leave the SM. */
size_t mpg_edge_workspace_bytes(int B, int N, int F, int H0, int H1, int H2);
int mpg_edge_fwd(const float* x, int ldx, const float* mask, const float* w0, const float* b0, const float* w1,
                 const float* b1, const float* w2, const float* b2, int B, int N, int F, int H0, int H1, int H2,
                 int ef_mode, int nd, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                 int precision, void* workspace, size_t workspace_bytes, float* agg, void* stream);
/* recomputes the edge activations tile by tile; dx [B*N, F] (row stride lddx) is overwritten */
int mpg_edge_bwd(const float* x, int ldx, const float* mask, const float* w0, const float* b0, const float* w1,
                 const float* b1, const float* w2, const float* b2, int B, int N, int F, int H0, int H1, int H2,
                 int ef_mode, int nd, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                 int precision, void* workspace, size_t workspace_bytes, const float* dagg, float* dx, int lddx, float* dw0,
                 float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream);

/* ---- fused node network fn (mpgan/model.py:268-279 with LinearNet :70-85) -------------------------
 * out = LinearNet([H1, H2] -> NO, final_linear)(cat(a, b)) for rows [M]: a [M, Ka] (the aggregated messages),
 * b [M, Kb] (the node features), w0 [H1, Ka+Kb], w1 [H2, H1], w2 [NO, H2] in the reference layout.  One
 * tcgen05 (TF32) kernel per direction; y0 [M, H1] and y1 [M, H2] (layer outputs) are saved for backward.
 * Dropout uses RNG streams 16, 17, 18 of `seed` (the streams LinearNet gives its layers), p in {0, 0.5}.
 * y0 / y1 hold the layer outputs rounded to TF32 (they only feed TF32 GEMMs and the leaky-relu sign test).
 * mpg_fn_supported() != 0 iff the shape is covered (Ka + Kb <= 256, H1, H2 in {128, 256}, NO <= 32, p in
 * {0, 0.5}); other shapes go through mpg_linear_* per layer. */
int mpg_fn_supported(int Ka, int Kb, int H1, int H2, int NO, float p_drop);
size_t mpg_fn_workspace_bytes(int Ka, int Kb, int H1, int H2, int NO);
int mpg_fn_fwd(const float* a, int lda, int Ka, const float* b, int ldb, int Kb, int M, const float* w0,
               const float* b0, const float* w1, const float* b1, const float* w2, const float* b2, int H1, int H2,
               int NO, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* workspace,
               size_t workspace_bytes, float* y0, float* y1, float* out, void* stream);
/* da [M, Ka] and db [M, Kb] (dense) are overwritten; dz0 [M, H1], dz1 [M, H2], dz2 [M, NO] are scratch (dz2 only
 * read/written when p_drop > 0).  Weight/bias gradients are ACCUMULATED (dw0 NULL: input gradients only). */
int mpg_fn_bwd(const float* dout, const float* y0, const float* y1, const float* a, int lda, int Ka, const float* b,
               int ldb, int Kb, int M, const float* w0, const float* w1, const float* w2, int H1, int H2, int NO,
               float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* workspace,
               size_t workspace_bytes, float* dz0, float* dz1, float* dz2, float* da, float* db, float* dw0,
               float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream);

/* The forward writes everything that depends only on (x, weights, mask) -- the factorised first layer P/Q, the
 * swizzled weight images, the work list -- into the first mpg_edge_fwd_workspace_bytes() bytes of its workspace
 * (a forward-only caller may pass just that much).  mpg_edge_bwd_saved() is mpg_edge_bwd() reading those from the
 * forward's still-intact workspace instead of recomputing them (4 fewer kernels per call); `workspace` is the
 * backward's own scratch of mpg_edge_workspace_bytes(). */
size_t mpg_edge_fwd_workspace_bytes(int B, int N, int F, int H0, int H1, int H2);
int mpg_edge_bwd_saved(const void* fwd_workspace, size_t fwd_workspace_bytes, const float* x, int ldx,
                       const float* mask, const float* w0, const float* b0, const float* w1, const float* b1,
                       const float* w2, const float* b2, int B, int N, int F, int H0, int H1, int H2, int ef_mode,
                       int nd, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                       int precision, void* workspace, size_t workspace_bytes, const float* dagg, float* dx, int lddx,
                       float* dw0, float* db0, float* dw1, float* db1, float* dw2, float* db2, void* stream);

/* ---- masks and tails ------------------------------------------------------------------------------ */
/* mask[b,i] = rank(x[b,i,0]) <= int(labels[b]*N) - 1   (bit-exact; mpgan/model.py:692-699) */
int mpg_rank_mask(const float* x, int ldx, const float* labels, int ldl, int B, int N, float* mask, void* stream);
/* Layout helpers: the layers are permutation-equivariant over particles (mpgan/model.py:206-282: fully connected,
 * sum / mean aggregation), so a jet's real particles may be moved first -- then a sender index that is padded in
 * every jet of a 128-particle tile is a step the edge kernels drop.  pos[b,i] (int32) = new index of particle i
 * (stable, mask != 0 first); mask_sorted = the mask in that order.  mpg_permute_rows: mode 0 scatters
 * dst[b, pos[b,i], :] = src[b, i, :], mode 1 gathers dst[b, i, :] = src[b, pos[b,i], :] (each the other's adjoint). */
int mpg_particle_order(const float* mask, int B, int N, int* pos, float* mask_sorted, void* stream);
int mpg_permute_rows(const float* src, int lds, float* dst, int ldd, const int* pos, int B, int N, int F, int mode,
                     void* stream);
/* pos[b] (int32) = index of jet b when the batch is ordered by descending key (key[b * ldk], e.g. the particle-count
 * label), ties in index order.  Jets never interact, so the batch order is a layout choice too: with
 * mpg_permute_rows(B = 1, N = batch, F = row length) it puts jets with similar padding into the same tiles. */
int mpg_batch_order(const float* key, int ldk, int B, int* pos, void* stream);
/* mask[r] = x[r, ldx-1] + 0.5   (mpgan/model.py:881) */
int mpg_split_mask(const float* x, int ldx, int rows, float* mask, void* stream);
/* out[r, :Fo] = act(h[r, :]); out[r, Fo] = mask[r] - 0.5 if mask   (mpgan/model.py:535-536, 752) */
int mpg_gen_tail_fwd(const float* h, const float* mask, float* out, int rows, int Fo, int act, void* stream);
int mpg_gen_tail_bwd(const float* dout, const float* out, float* dh, int rows, int Fo, int ldo, int act,
                     void* stream);
/* pooled[b,:] = sum_i h[b,i,:]*mask[b,i]  (/(sum mask + 1e-12) if mean)   (mpgan/model.py:810-822) */
int mpg_pool_fwd(const float* h, const float* mask, float* out, int B, int N, int C, int mean, void* stream);
int mpg_pool_bwd(const float* dout, const float* mask, float* dh, int B, int N, int C, int mean, void* stream);
int mpg_unary_fwd(const float* x, float* y, size_t n, int act, void* stream);
int mpg_unary_bwd(const float* dy, const float* y, float* dx, size_t n, int act, void* stream);

/* ---- spectral norm (spectral_normalization.py:21-33): one power iteration, u/v updated in place --- */
int mpg_sn_fwd(const float* w_bar, float* u, float* v, float* w_out, float* sigma, int H, int W, void* stream);
int mpg_sn_bwd(const float* dw, const float* w_bar, const float* u, const float* v, const float* sigma,
               float* dw_bar, int H, int W, void* stream);

/* ---- least-squares GAN loss (train.py:357-358,369-370,378 for D; :467,472 for G): MSELoss against constant
 * targets.  loss = mean_{i<n0}(d_i - t0)^2 + mean_{i>=n0}(d_i - t1)^2 (second term absent when n0 == n), d = the
 * discriminator outputs [n]; bwd: dd_i = *gout * dloss/dd_i (gout: device scalar). */
int mpg_ls_loss_fwd(const float* d, int n, int n0, float t0, float t1, float* loss, void* stream);
int mpg_ls_loss_bwd(const float* d, const float* gout, int n, int n0, float t0, float t1, float* dd, void* stream);

/* ---- optimizer: torch.optim.RMSprop(alpha, eps) on a flat buffer; g is scaled by gscale first ------ */
int mpg_rmsprop(float* p, const float* g, float* sq, size_t n, float lr, float alpha, float eps, float gscale,
                void* stream);

/* ---- GAPT set attention (gapt/model.py:124-139; nn.MultiheadAttention with a key mask) -------------
 * o[b,i,h*d:(h+1)*d] = softmax_j(q_i.k_j / sqrt(d) | key j not ignored) v_j, per head h (d = E/heads).
 * q/k/v are rows [B*N, E] with row strides ld* (a packed [B*N, 3E] in_proj output needs no split).
 * key_mask [B, Nk]: keys with mask != 1.0 are ignored (gapt/model.py:194-202: (1 - mask).bool());
 * NULL = attend to every key.  p_saved [B, heads, Nq, Nk] keeps the probabilities for backward. */
int mpg_attn_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* key_mask,
                 int B, int Nq, int Nk, int E, int heads, float* o, float* p_saved, void* stream);
/* dq/dk/dv are dense [B*N, E] and overwritten */
int mpg_attn_bwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* key_mask,
                 int B, int Nq, int Nk, int E, int heads, const float* p_saved, const float* dout, float* dq,
                 float* dk, float* dv, void* stream);
/* out = dropout(x + r) (r may be NULL): the residual + nn.Dropout steps of MAB.forward (:129-137) */
int mpg_residual_dropout_fwd(const float* x, const float* r, float* out, size_t rows, int cols, float p_drop,
                             uint64_t seed, const uint64_t* seed_dev, uint32_t rng_stream, void* stream);
int mpg_residual_dropout_bwd(const float* dout, float* dx, size_t rows, int cols, float p_drop, uint64_t seed,
                             const uint64_t* seed_dev, uint32_t rng_stream, void* stream);

/* ---- k-nearest-neighbour message passing (mpgan/model.py:319-381, `fully_connected=False`) ----------------------
 * mpg_knn_select: idx[b,i,m] (int32, [B,N,k]) = index inside jet b of the m-th nearest sender of receiver i, by the
 * Euclidean distance over the first nd features to the sender scaled by (1 - 1e4) * mask + 1e4 (masked particles
 * rank last, :336-340), 1e-12 added per component (:348); ascending, ties by index; self_loops = 0 skips the nearest
 * entry (:354-363).  mpg_edge_nbr_fwd / _bwd: the fused edge network of mpg_edge_fwd / _bwd with receiver i
 * aggregating over its k listed senders only (mask = the listed sender's mask, mean = 1/k); ef_mode 1 feeds the
 * distance (to the scaled sender when knn_scale != 0) as the pair feature.  fp32 SIMT kernels.  With nbr == NULL
 * both are the fully connected op on the fp32 kernels; dmask (optional, [B*N], overwritten) receives the gradient
 * w.r.t. the mask multiplier: dmask[b,j] = scale * sum_i <fe(x_i | x_j), dagg[b,i]>.  lc (optional, [B, H0]): the
 * first-layer contribution of the conditioning columns (clabels / mask_fne_np, :247-253), Lc = cond W0c^T; pair row r
 * of the [B*N*K] edge list adds lc[r % B] (the reference's `.repeat` pairs row r with jet r % B); dlc [B, H0]
 * (overwritten) receives its gradient. */
int mpg_knn_select(const float* x, int ldx, const float* mask, int B, int N, int nd, int k, int self_loops, int* idx,
                   void* stream);
int mpg_edge_nbr_fwd(const int* nbr, int K, int knn_scale, const float* lc, const float* x, int ldx, const float* mask,
                     const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                     int B, int N, int F, int H0, int H1, int H2, int ef_mode, int nd, int mean, float alpha, float p_drop,
                     uint64_t seed, const uint64_t* seed_dev, void* workspace, size_t workspace_bytes, float* agg,
                     void* stream);
int mpg_edge_nbr_bwd(const int* nbr, int K, int knn_scale, const float* lc, const float* x, int ldx, const float* mask,
                     const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                     int B, int N, int F, int H0, int H1, int H2, int ef_mode, int nd, int mean, float alpha, float p_drop,
                     uint64_t seed, const uint64_t* seed_dev, void* workspace, size_t workspace_bytes, const float* dagg,
                     float* dx, int lddx, float* dmask, float* dlc, float* dw0, float* db0, float* dw1, float* db1,
                     float* dw2, float* db2, void* stream);
/* out[r, :F] = x[r, :], out[r, F + t] = cond[r % B, t]: the node network's conditioning columns (labels / particle
 * count, mpgan/model.py:270-276; the reference's `.repeat(num_nodes, 1)` gives row r the entry r % B). */
int mpg_cond_columns(const float* x, int ldx, const float* cond, int C, float* out, size_t rows, int F, int B, void* stream);

/* ---- second-order products of the fused edge op: double backward for WGAN-GP (train.py:286-324) -----------------
 * fe is piecewise linear, so with u [B*N, F] the cotangent the double backward receives for mpg_edge_bwd's dx output,
 * s = <u, dx> depends on (dagg, mask, weights) through the tangent t = J_fe u taken along the slopes / dropout masks
 * of the primal pass:  tagg [B*N, H2] = ds/d(dagg) = scale * sum_j mask_j t_ij;  gmask [B*N] (optional) = ds/d(mask);
 * dw0 [H0, 2F], dw1, dw2 += ds/dW (optional, all or none; the biases and x have no second-order term).  Pair
 * features (ef_mode != 0) are not supported.  fp32 SIMT kernel. */
size_t mpg_edge_bwd2_workspace_bytes(int B, int N, int F, int H0, int H1, int H2);
int mpg_edge_bwd2(const float* x, int ldx, const float* u, int ldu, const float* mask, const float* w0, const float* b0,
                  const float* w1, const float* b1, const float* w2, const float* b2, int B, int N, int F, int H0, int H1,
                  int H2, int mean, float alpha, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* workspace,
                  size_t workspace_bytes, const float* dagg, float* tagg, float* gmask, float* dw0, float* dw1, float* dw2,
                  void* stream);

/* ---- receiver compaction (tcgen05 edge path; exact for the discriminator, mpgan/model.py:810-822,881-884: padded
 * particles are masked as senders and multiplied by the mask at the pooling, so nothing they receive is ever used).
 * mpg_compact_map writes the map (mpg_compact_map_ints(B, N) ints; `scratch`: 2B + 8 ints) that packs the rows with
 * mask != 0 into 128-row tiles; mpg_edge_set_compaction(map) makes THIS THREAD's following mpg_edge_fwd / mpg_edge_bwd /
 * mpg_edge_bwd_saved calls build their tiles from it (NULL: off).  Tensors keep their padded [B*N, .] layout; rows
 * outside every tile get agg = 0 and dx = 0.  Ignored on the fp32 path.  Do not use for a generator (its padded rows are
 * part of its output, mpgan/model.py:723-752). */
size_t mpg_compact_map_ints(int B, int N);
int mpg_compact_map(const float* mask, int B, int N, int* cmap, int* scratch, void* stream);
int mpg_edge_set_compaction(const int* cmap);

/* ---- mask-channel gradients (the discriminator's mask = x[..., -1] + 0.5 is differentiable, mpgan/model.py:881) --
 * mpg_split_mask_bwd: dx [rows, ldx] = 0 except dx[r, ldx-1] = dmask[r];  mpg_pool_dmask: dmask[b,i] = scale *
 * <h[b,i,:], dout[b,:]> (gradient of the masked sum pool w.r.t. the mask). */
int mpg_split_mask_bwd(const float* dmask, float* dx, int ldx, size_t rows, void* stream);
int mpg_pool_dmask(const float* h, const float* dout, float* dmask, int B, int N, int C, float scale, void* stream);

/* ---- generation post-processing (gen.py:126-141): out[r, i] = ((jets[r, i] - shift[i]) / norm[i]) * maxv[i] for
 * i < nfeat (shift / norm / maxv are HOST arrays; a NaN entry = gen.py's None: step skipped), rows whose mask channel
 * jets[r, ldj-1] < 0.5 zeroed (use_mask), feature 2 clamped at 0.  `out` [rows, ldo] may be pinned host memory. */
int mpg_gen_postprocess(const float* jets, int ldj, float* out, int ldo, size_t rows, int nfeat, const float* shift,
                        const float* norm, const float* maxv, int use_mask, void* stream);

/* ---- LayerNorm over the last dimension (gapt/model.py:116-118, 130-136: nn.LayerNorm(embed_dim), eps 1e-5) --------
 * mean / rstd [rows] are saved for backward; dw / db (optional, both or none) are ACCUMULATED. */
int mpg_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, size_t rows,
                      int C, float eps, void* stream);
int mpg_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, float* dx,
                      float* dw, float* db, size_t rows, int C, void* stream);

/* ---- fused GAPT attention block (MAB.forward, gapt/model.py:124-139, layer_norm = False) -------------------------
 * out = Dropout(h + Dropout(leaky_relu(h Wff^T + bff))),  h = Dropout(x + MultiheadAttention(x, y, y, key mask)) as
 * ONE kernel per direction: x [B*Nq, E] (row stride ldx) is the query side, y [B*Nk, E] the key/value side (pass the
 * same pointer for self attention); w_in [3E, E] / b_in [3E] = nn.MultiheadAttention.in_proj_*, w_out / b_out =
 * out_proj, w_ff / b_ff = ff.net.0 (ff_layers = []).  Supported (mpg_mab_supported): E = 64, 4 heads, Nq, Nk <= 32 --
 * SAB, ISAB (10 inducing points) and PMA (1 seed) on 30-particle jets.  p_res = MAB's dropout, p_ff = the LinearNet's.
 * precision 0: fp32 SIMT arithmetic; 1: the five projections (and their gradients) on TF32 mma.sync tiles, fp32
 * accumulate, attention core in fp32.
 * The forward leaves q [B*Nq,E], kv [B*Nk,2E], o, h, f [B*Nq,E] for the backward, which recomputes the attention
 * probabilities, overwrites dx (and dy when y != x) and ACCUMULATES the six parameter gradients (all or none). */
int mpg_mab_supported(int E, int heads, int Nq, int Nk);
size_t mpg_mab_workspace_bytes(int B);
int mpg_mab_fwd(const float* x, int ldx, const float* y, int ldy, const float* key_mask, const float* w_in,
                const float* b_in, const float* w_out, const float* b_out, const float* w_ff, const float* b_ff, int B,
                int Nq, int Nk, int E, int heads, float alpha, float p_res, float p_ff, uint64_t seed,
                const uint64_t* seed_dev, int precision, void* workspace, size_t workspace_bytes, float* q, float* kv,
                float* o, float* h, float* f, float* out, void* stream);
int mpg_mab_bwd(const float* x, int ldx, const float* y, int ldy, const float* key_mask, const float* w_in,
                const float* b_in, const float* w_out, const float* b_out, const float* w_ff, const float* b_ff, int B,
                int Nq, int Nk, int E, int heads, float alpha, float p_res, float p_ff, uint64_t seed,
                const uint64_t* seed_dev, int precision, void* workspace, size_t workspace_bytes, const float* q,
                const float* kv, const float* o, const float* h, const float* f, const float* dout, float* dx, float* dy,
                float* dw_in, float* db_in, float* dw_out, float* db_out, float* dw_ff, float* db_ff, void* stream);

/* ---- data-parallel update: one-shot gradient all-reduce fused with RMSprop over NVLink peer memory ------------------
 * (replaces DataParallel's gradient reduction, setup_training.py:1418-1421, + torch.optim.RMSprop, :1511-1513.)
 * peer_grads[r] / peer_flags[r] (HOST arrays of `world` device pointers) are rank r's flat gradient buffer and flag
 * block as mapped on THIS GPU (symmetric / peer memory, e.g. torch.distributed._symmetric_memory); the flag blocks hold
 * mpg_peer_flag_words(ctas, world) zero-initialised 32-bit words.  Every rank calls it once per step on its own
 * stream with the same n / ctas: the kernel waits (system-scope flags, bounded spin) until every peer's gradients
 * are complete, sums them in rank order, applies p -= lr * g / (sqrt(sq) + eps) with g = sum / world and
 * sq = alpha * sq + (1 - alpha) * g^2, and returns only after every peer has finished reading this rank's
 * gradients.  grads_multicast (optional): the NVLink-SHARP multicast address of the gradient buffers; the sum is then
 * ONE multimem.ld_reduce per 16 bytes, reduced inside the NVSwitch, instead of `world` peer loads.  ctas <= the SM
 * count (all CTAs must be co-resident).  Graph-capturable (the epoch lives on the device). */
size_t mpg_peer_flag_words(int ctas, int world);
int mpg_allreduce_rmsprop(float* p, float* sq, const void* const* peer_grads, void* const* peer_flags,
                          const void* grads_multicast, size_t n, int rank, int world, int ctas, float lr, float alpha,
                          float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPGAN_B200_H_ */
