#pragma once
#include "common.cuh"
namespace mpg {
int launch_rank_mask(const float* x, int ldx, const float* labels, int ldl, int B, int N, float* mask, cudaStream_t s);
int launch_act_bwd(const float* dy, const float* y, float* dz, int M, int N, int act, float alpha, DropCfg dc,
                   uint32_t stream, cudaStream_t s);
int launch_gen_tail_fwd(const float* h, const float* mask, float* out, int rows, int Fo, int act, cudaStream_t s);
int launch_gen_tail_bwd(const float* dout, const float* out, float* dh, int rows, int Fo, int ldo, int act,
                        cudaStream_t s);
int launch_particle_order(const float* mask, int B, int N, int* pos, float* mask_sorted, cudaStream_t s);
int launch_batch_order(const float* key, int ldk, int B, int* pos, cudaStream_t s);
int launch_permute_rows(const float* src, int lds, float* dst, int ldd, const int* pos, int B, int N, int F, int mode,
                        cudaStream_t s);
int launch_ls_loss(const float* d, const float* gout, int n, int n0, float t0, float t1, float* out, bool bwd,
                   cudaStream_t s);
int launch_split_mask(const float* x, int ldx, int rows, float* mask, cudaStream_t s);
int launch_pool_fwd(const float* h, const float* mask, float* out, int B, int N, int C, int mean, cudaStream_t s);
int launch_pool_bwd(const float* dout, const float* mask, float* dh, int B, int N, int C, int mean, cudaStream_t s);
int launch_unary(const float* x, const float* dy, float* out, size_t n, int act, bool bwd, cudaStream_t s);
int launch_sn_fwd(const float* Wb, float* u, float* v, float* Wout, float* sigma, int H, int Wd, cudaStream_t s);
int launch_sn_bwd(const float* dW, const float* Wb, const float* u, const float* v, const float* sigma, float* dWb,
                  int H, int Wd, cudaStream_t s);
int launch_rmsprop(float* p, const float* g, float* sq, size_t n, float lr, float alpha, float eps, float gscale,
                   cudaStream_t s);
}  // namespace mpg
