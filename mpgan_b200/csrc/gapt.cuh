// GAPT set-attention core and residual+dropout kernels (gapt/model.py:124-139).
#pragma once
#include "common.cuh"
namespace mpg {
struct AttnArgs {
  const float* q; int ldq;      // [B*Nq, E] rows (row stride ldq: packed QKV projections need no split)
  const float* k; int ldk;      // [B*Nk, E]
  const float* v; int ldv;
  const float* key_mask;        // [B, Nk] or null; keys with mask != 1.0 are ignored
  int B, Nq, Nk, E, heads;
};
int launch_attn_fwd(const AttnArgs& a, float* o, float* P, cudaStream_t s);
int launch_attn_bwd(const AttnArgs& a, const float* P, const float* dO, float* dq, float* dk, float* dv,
                    cudaStream_t s);
int launch_resdrop(const float* x, const float* r, float* out, size_t rows, int cols, DropCfg dc, uint32_t stream,
                   bool bwd, cudaStream_t s);
}  // namespace mpg
