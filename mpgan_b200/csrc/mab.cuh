// Fused GAPT multihead-attention block (mab.cu).
#pragma once
#include "common.cuh"
namespace mpg {
struct MabArgs {
  const float* x; int ldx;        // query-side input [B*Nq, 64]
  const float* y; int ldy;        // key/value-side input [B*Nk, 64] (== x for self attention)
  const float* key_mask;          // [B*Nk] JetNet mask (keys with mask != 1 ignored) or null
  const float *w_in, *b_in, *w_out, *b_out, *w_ff, *b_ff;   // reference layouts: [192,64],[192],[64,64],[64],[64,64],[64]
  int B, Nq, Nk;
  float alpha;
  DropCfg drop_res, drop_ff;      // MAB dropout (streams 48, 49) and the feed-forward LinearNet's dropout (stream 16)
  float *q, *kv, *o, *h, *f, *out;   // saved activations + output (forward writes, backward reads)
};
struct MabGrads {
  const float* dout;              // [B*Nq, 64]
  float* dx;                      // [B*Nq, 64] (self attention: includes the key/value path)
  float* dy;                      // [B*Nk, 64] (cross attention only)
  float* slab;                    // filled in by the launcher
  float *dw_in, *db_in, *dw_out, *db_out, *dw_ff, *db_ff;   // accumulated; dw_in null = input gradients only
};
bool mab_supported(int E, int heads, int Nq, int Nk);
size_t mab_workspace_bytes(int B);
int launch_mab_fwd(const MabArgs& a, void* workspace, int precision, cudaStream_t s);
int launch_mab_bwd(const MabArgs& a, MabGrads g, void* workspace, int precision, cudaStream_t s);
}  // namespace mpg
