// Round-2 additions: generation post-processing, mask-channel gradients, GAPT LayerNorm, kNN neighbour selection.
#pragma once
#include "common.cuh"
namespace mpg {
struct PostCfg {
  int nfeat;             // <= 8
  float shift[8], norm[8], maxv[8];
  unsigned has_shift, has_norm;   // bit i: apply to feature i
};
int launch_gen_postprocess(const float* jets, int ldj, float* out, int ldo, size_t rows, const PostCfg& c, int use_mask,
                           cudaStream_t s);
int launch_cond_columns(const float* x, int ldx, const float* cond, int C, float* out, size_t rows, int F, int B,
                        cudaStream_t s);
int launch_compact_map(const float* mask, int B, int N, int* cmap, int* scratch, cudaStream_t s);
int launch_split_mask_bwd(const float* dmask, float* dx, int ldx, size_t rows, cudaStream_t s);
int launch_pool_dmask(const float* h, const float* dout, float* dmask, int B, int N, int C, float scale, cudaStream_t s);
int launch_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, size_t rows,
                         int C, float eps, cudaStream_t s);
int launch_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, float* dx,
                         float* dw, float* db, size_t rows, int C, cudaStream_t s);
int launch_knn_select(const float* x, int ldx, const float* mask, int B, int N, int nd, int k, int skip, int* idx,
                      cudaStream_t s);
}  // namespace mpg
