// Node-level GEMM family (rows = particles, not pairs): C = epi(op(A) * op(B)).
//
// Used for everything that is O(B*N) rather than O(B*N^2): the factorised first edge layer
// (P = x*Wa^T + b0, Q = x*Wb^T), the node network fn, fnd, and all of their backward GEMMs.
// TF32 tensor-core mma.sync with fp32 accumulate; PRECISE=true runs the 3xTF32 error-compensated
// split (fp32-class accuracy) and is what the bit-tight parity tests use.
#pragma once
#include "common.cuh"

namespace mpg {

struct GemmEpi {
  const float* bias = nullptr;   // [N], added before activation
  int act = 0;                   // 1: leaky_relu(alpha)
  float alpha = 0.2f;
  int drop = 0;                  // 1: dropout on output element (row m, col n) of stream `stream`
  DropCfg dc{};
  uint32_t stream = 0;
  // multiply by d(act+dropout)/dz of a previous layer whose OUTPUT is gy (backward fusion)
  const float* gy = nullptr;
  int ldgy = 0;
  int g_act = 0;
  int g_drop = 0;
  DropCfg gdc{};
  uint32_t gstream = 0;
  float scale = 1.f;             // result *= scale (before everything else)
  int accumulate = 0;            // C += result
  int atomic = 0;                // use atomicAdd (split-K)
};

// A_K: A is [M,K] with K contiguous (lda = row stride); else stored [K,M] (M contiguous).
// B_K: B is [N,K] with K contiguous;                    else stored [K,N] (N contiguous).
int launch_gemm(bool a_k, bool b_k, bool precise, const float* A, int lda, const float* B, int ldb,
                float* C, int ldc, int M, int N, int K, const GemmEpi& epi, int split_k,
                cudaStream_t stream);

// column sums: out[n] (+)= sum_m X[m, n] * g(gy)   (bias gradients)
int launch_colsum(const float* X, int ldx, int M, int N, float* out, cudaStream_t stream);

}  // namespace mpg
