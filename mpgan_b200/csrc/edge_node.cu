// Node-level ends of the factorised first edge layer (W0 [x_i ; x_j ; ef] = Wa x_i + Wb x_j + Wef ef).
//
//   forward   P = x Wa^T + b0,  Q = x Wb^T                                    [B*N, H0] each
//   backward  dx = dP Wa + dQ Wb,  dWa += dP^T x,  dWb += dQ^T x,  db0 += sum_r dP
//
// These are O(B*N) x (F <= 64) x (H0 <= 128) problems: a few MFLOP that used to cost six TF32 GEMM
// launches + a column sum (~20 us each, latency-bound).  One fp32 SIMT kernel per direction does the
// same work from shared memory in a few microseconds, exactly (no TF32 rounding).
// Reference: first Linear of fe applied to cat(x_i, x_j), mpgan/model.py:77-83, 294-311.
#include "edge_node.cuh"

namespace mpg {
namespace {

constexpr int BWD_ROWS = 64;     // rows per block, backward (two blocks per SM)
constexpr int NT = 256;

// Ws[c][k] = W0[k][c] for c < 2F (column c of the weight, all H0 outputs), stride H0P
__device__ __forceinline__ void load_w_t(float* Ws, const float* __restrict__ W0, int ldw, int F, int H0, int H0P) {
  for (int idx = threadIdx.x; idx < 2 * F * H0; idx += NT) {
    const int k = idx / (2 * F), c = idx % (2 * F);     // consecutive threads: consecutive columns of one row (coalesced)
    Ws[c * H0P + k] = W0[(size_t)k * ldw + c];
  }
}

__global__ void __launch_bounds__(PQ_NT) pq_fwd_kernel(PqFwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  pq_fwd_tile(a, blockIdx.x, sm);
}

// dx, dW0[:, :2F] += [dP^T x | dQ^T x], db0 += column sums of dP.  Shared-memory bandwidth (one warp load per
// clock) is the limit, so both products are register-blocked around 16-byte loads: 4 rows x 4 k per W/d quad
// for dx, 24 gradient columns per x value for dW0.
__global__ void __launch_bounds__(NT) pq_bwd_kernel(const float* __restrict__ dP, const float* __restrict__ dQ,
                                                    const float* __restrict__ x, int ldx, const float* __restrict__ W0,
                                                    int ldw, float* __restrict__ dx, int lddx, float* __restrict__ dW0,
                                                    float* __restrict__ db0, int BN, int F, int H0, int p_tiled) {
  extern __shared__ __align__(16) float sm[];
  const int H0P = H0 + 4, FP = F + 1, DP = 2 * H0 + 4;   // H0 % 4 == 0 (launcher): rows stay 16-byte aligned
  float* Ws = sm;                          // [2F][H0P]   Ws[c][k] = W0[k][c]
  float* ds = Ws + 2 * F * H0P;            // [BWD_ROWS][DP]  row = [dP | dQ]
  float* xs = ds + BWD_ROWS * DP;          // [BWD_ROWS][FP]
  const int r0 = blockIdx.x * BWD_ROWS;
  load_w_t(Ws, W0, ldw, F, H0, H0P);
  for (int idx = threadIdx.x; idx < BWD_ROWS * F; idx += NT) {
    const int r = idx / F, f = idx % F;
    xs[r * FP + f] = r0 + r < BN ? x[(size_t)(r0 + r) * ldx + f] : 0.f;
  }
  if (p_tiled) {   // dP: 16-byte column groups, 64 consecutive rows of a group are contiguous; dQ: row-major
    const int ng = H0 / 4;
    for (int idx = threadIdx.x; idx < ng * BWD_ROWS; idx += NT) {
      const int g = idx / BWD_ROWS, r = idx % BWD_ROWS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < BN) v = *reinterpret_cast<const float4*>(dP + p_tiled_index(r0 + r, 4 * g, H0));
      *reinterpret_cast<float4*>(ds + r * DP + 4 * g) = v;
    }
    for (int idx = threadIdx.x; idx < BWD_ROWS * ng; idx += NT) {
      const int r = idx / ng, g = idx % ng;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < BN) v = *reinterpret_cast<const float4*>(dQ + (size_t)(r0 + r) * H0 + 4 * g);
      *reinterpret_cast<float4*>(ds + r * DP + H0 + 4 * g) = v;
    }
  } else {
    for (int idx = threadIdx.x; idx < BWD_ROWS * 2 * H0; idx += NT) {
      const int r = idx / (2 * H0), c = idx % (2 * H0);
      float v = 0.f;
      if (r0 + r < BN) v = c < H0 ? dP[(size_t)(r0 + r) * H0 + c] : dQ[(size_t)(r0 + r) * H0 + c - H0];
      ds[r * DP + c] = v;
    }
  }
  __syncthreads();
  // ---- dx[r][f] = sum_k dP[r][k] Wa[k][f] + dQ[r][k] Wb[k][f]: 8 rows (independent accumulators) per thread ----
  for (int idx = threadIdx.x; idx < (BWD_ROWS / 8) * F; idx += NT) {
    const int f = idx % F, rg = idx / F;
    const float4* wa = reinterpret_cast<const float4*>(Ws + f * H0P);
    const float4* wb = reinterpret_cast<const float4*>(Ws + (F + f) * H0P);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int k4 = 0; k4 < H0 / 4; ++k4) {
      const float4 a4 = wa[k4], b4 = wb[k4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* d = ds + (rg * 8 + i) * DP;
        const float4 p4 = *reinterpret_cast<const float4*>(d + 4 * k4);
        const float4 q4 = *reinterpret_cast<const float4*>(d + H0 + 4 * k4);
        acc[i] = fmaf(p4.x, a4.x, acc[i]); acc[i] = fmaf(q4.x, b4.x, acc[i]);
        acc[i] = fmaf(p4.y, a4.y, acc[i]); acc[i] = fmaf(q4.y, b4.y, acc[i]);
        acc[i] = fmaf(p4.z, a4.z, acc[i]); acc[i] = fmaf(q4.z, b4.z, acc[i]);
        acc[i] = fmaf(p4.w, a4.w, acc[i]); acc[i] = fmaf(q4.w, b4.w, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + rg * 8 + i;
      if (r < BN) dx[(size_t)r * lddx + f] = acc[i];
    }
  }
  if (dW0 == nullptr) return;   // input gradient only
  // ---- dW0[k][c] += sum_r d[r][k (+H0 for c >= F)] x[r][c mod F]: 8 consecutive gradient columns per thread --
  const int ngrp = 2 * H0 / 8;
  for (int idx = threadIdx.x; idx < ngrp * F; idx += NT) {
    const int f = idx % F, g = idx / F;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 8
    for (int r = 0; r < BWD_ROWS; ++r) {
      const float xv = xs[r * FP + f];
      const float4 d0 = *reinterpret_cast<const float4*>(ds + r * DP + 8 * g);
      const float4 d1 = *reinterpret_cast<const float4*>(ds + r * DP + 8 * g + 4);
      acc[0] = fmaf(d0.x, xv, acc[0]); acc[1] = fmaf(d0.y, xv, acc[1]);
      acc[2] = fmaf(d0.z, xv, acc[2]); acc[3] = fmaf(d0.w, xv, acc[3]);
      acc[4] = fmaf(d1.x, xv, acc[4]); acc[5] = fmaf(d1.y, xv, acc[5]);
      acc[6] = fmaf(d1.z, xv, acc[6]); acc[7] = fmaf(d1.w, xv, acc[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int dc = 8 * g + i;                       // column of [dP | dQ]
      const int k = dc < H0 ? dc : dc - H0;
      atomicAdd(dW0 + (size_t)k * ldw + (dc < H0 ? f : F + f), acc[i]);
    }
  }
  for (int k = threadIdx.x; k < H0; k += NT) {
    float acc = 0.f;
    for (int r = 0; r < BWD_ROWS; ++r) acc += ds[r * DP + k];
    atomicAdd(db0 + k, acc);
  }
}

// ---- the same products on tensor cores (precision 1: TF32 mma.sync.m16n8k8, fp32 accumulation) -------------------------
// The fp32 kernel above is FMA-bound (1.6 MFMA per 64-row block); here the block is bound by its loads.
//   dx[64 x F]    = [dP | dQ][64 x 2*H0] * [Wa ; Wb][2*H0 x F]        16-row x 8-column tiles, K = 2*H0
//   dW0[2*H0 x F] += [dP | dQ]^T[2*H0 x 64] * x[64 x F]               K = the block's 64 rows
// Operand tiles sit in shared memory with row pitches that keep the fragment loads conflict-free (or 2-way).
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__host__ __device__ inline int pq_f8(int F) { return (F + 7) & ~7; }
size_t bwd_tc_smem(int F, int H0) {
  const int F8 = pq_f8(F);
  return (size_t)(2 * F8 * (H0 + 4) + BWD_ROWS * (F8 + 8) + BWD_ROWS * (2 * H0 + 4)) * sizeof(float);
}
__global__ void __launch_bounds__(NT) pq_bwd_tc_kernel(const float* __restrict__ dP, const float* __restrict__ dQ,
                                                       const float* __restrict__ x, int ldx, const float* __restrict__ W0,
                                                       int ldw, float* __restrict__ dx, int lddx, float* __restrict__ dW0,
                                                       float* __restrict__ db0, int BN, int F, int H0, int p_tiled,
                                                       const int* __restrict__ cmap, int ctiles_max) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int rmap[BWD_ROWS];           // padded row of each of the block's rows (-1: none)
  {
    const int pos0 = blockIdx.x * BWD_ROWS;   // receiver compaction: positions in the compacted tile space
    if (cmap != nullptr && pos0 >= cmap[0] * 128) return;   // (block-uniform)
    for (int r = threadIdx.x; r < BWD_ROWS; r += NT)
      rmap[r] = cmap != nullptr ? cmap[2 + 2 * ctiles_max + pos0 + r] : (pos0 + r < BN ? pos0 + r : -1);
    __syncthreads();
  }
  const int F8 = pq_f8(F), H0P = H0 + 4, FP = F8 + 8, DP = 2 * H0 + 4;
  float* Ws = sm;                          // [2*F8][H0P]  Ws[h*F8 + f][k] = W0[k][h*F + f], zero rows for f >= F
  float* ds = Ws + 2 * F8 * H0P;           // [BWD_ROWS][DP]  row = [dP | dQ]
  float* xs = ds + BWD_ROWS * DP;          // [BWD_ROWS][FP]  zero columns for f >= F
  const int r0 = blockIdx.x * BWD_ROWS;
  const int ng = H0 / 4;
  {   // gradient rows first (the largest item), 16 bytes per load, every load of a thread in flight together
    constexpr int U = 6;
    for (int i0 = threadIdx.x; i0 < 2 * ng * BWD_ROWS; i0 += NT * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = i0 + u * NT;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < ng * BWD_ROWS) {                       // dP: tiled -> 64 consecutive rows of a column group are contiguous
          const int g = idx / BWD_ROWS, r = idx % BWD_ROWS;
          if (rmap[r] >= 0)   // tiled dP is indexed by the position in the (compacted) tile space
            v[u] = *reinterpret_cast<const float4*>(dP + (p_tiled ? p_tiled_index(r0 + r, 4 * g, H0) : (size_t)rmap[r] * H0 + 4 * g));
        } else if (idx < 2 * ng * BWD_ROWS) {            // dQ: row-major, padded rows
          const int j = idx - ng * BWD_ROWS, r = j / ng, g = j % ng;
          if (rmap[r] >= 0) v[u] = *reinterpret_cast<const float4*>(dQ + (size_t)rmap[r] * H0 + 4 * g);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = i0 + u * NT;
        if (idx < ng * BWD_ROWS) {
          const int g = idx / BWD_ROWS, r = idx % BWD_ROWS;
          *reinterpret_cast<float4*>(ds + r * DP + 4 * g) = v[u];
        } else if (idx < 2 * ng * BWD_ROWS) {
          const int j = idx - ng * BWD_ROWS, r = j / ng, g = j % ng;
          *reinterpret_cast<float4*>(ds + r * DP + H0 + 4 * g) = v[u];
        }
      }
    }
  }
  batched_fill<8, NT>(2 * F8 * H0,
      [&](int idx) {
        const int k = idx / (2 * F8), c = idx % (2 * F8), h = c >= F8 ? 1 : 0, f = c - h * F8;
        return f < F ? W0[(size_t)k * ldw + h * F + f] : 0.f;
      },
      [&](int idx, float v) { const int k = idx / (2 * F8), c = idx % (2 * F8); Ws[c * H0P + k] = v; });
  batched_fill<8, NT>(BWD_ROWS * F8,
      [&](int idx) { const int r = idx / F8, f = idx % F8; return (rmap[r] >= 0 && f < F) ? x[(size_t)rmap[r] * ldx + f] : 0.f; },
      [&](int idx, float v) { const int r = idx / F8, f = idx % F8; xs[r * FP + f] = v; });
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nnt = F8 / 8;                  // 8-column tiles of the F axis
  // ---- dx: (BWD_ROWS / 16) x nnt tiles over the warps ---------------------------------------------------------------
  for (int tile = warp; tile < (BWD_ROWS / 16) * nnt; tile += NT / 32) {
    const int m0 = (tile / nnt) * 16, n0 = (tile % nnt) * 8;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int h = 0; h < 2; ++h) {
      const float* A = ds + h * H0;                       // dP or dQ columns
      const float* B = Ws + (size_t)(h * F8 + n0 + g) * H0P;   // B[k][n] = Ws[n][k]
      for (int k0 = 0; k0 < H0; k0 += 8) {
        const uint32_t a[4] = {to_tf32(A[(m0 + g) * DP + k0 + t]), to_tf32(A[(m0 + g + 8) * DP + k0 + t]),
                               to_tf32(A[(m0 + g) * DP + k0 + t + 4]), to_tf32(A[(m0 + g + 8) * DP + k0 + t + 4])};
        const uint32_t b[2] = {to_tf32(B[k0 + t]), to_tf32(B[k0 + t + 4])};
        mma_tf32(c, a, b);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = rmap[m0 + g + 8 * (e >> 1)], f = n0 + 2 * t + (e & 1);
      if (r >= 0 && f < F) dx[(size_t)r * lddx + f] = c[e];
    }
  }
  if (dW0 == nullptr) return;   // input gradient only
  // ---- dW0: (2*H0 / 16) x nnt tiles; A[m = column of [dP | dQ]][k = row] is ds read transposed ---------------------------
  for (int tile = warp; tile < (2 * H0 / 16) * nnt; tile += NT / 32) {
    const int m0 = (tile / nnt) * 16, n0 = (tile % nnt) * 8;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int k0 = 0; k0 < BWD_ROWS; k0 += 8) {
      const uint32_t a[4] = {to_tf32(ds[(k0 + t) * DP + m0 + g]), to_tf32(ds[(k0 + t) * DP + m0 + g + 8]),
                             to_tf32(ds[(k0 + t + 4) * DP + m0 + g]), to_tf32(ds[(k0 + t + 4) * DP + m0 + g + 8])};
      const uint32_t b[2] = {to_tf32(xs[(k0 + t) * FP + n0 + g]), to_tf32(xs[(k0 + t + 4) * FP + n0 + g])};
      mma_tf32(c, a, b);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int dc = m0 + g + 8 * (e >> 1), f = n0 + 2 * t + (e & 1);   // column of [dP | dQ], input feature
      if (f < F) {
        const int k = dc < H0 ? dc : dc - H0;
        atomicAdd(dW0 + (size_t)k * ldw + (dc < H0 ? f : F + f), c[e]);
      }
    }
  }
  for (int k = threadIdx.x; k < H0; k += NT) {
    float acc = 0.f;
    for (int r = 0; r < BWD_ROWS; ++r) acc += ds[r * DP + k];
    atomicAdd(db0 + k, acc);
  }
}

size_t bwd_smem(int F, int H0) {
  return (size_t)(2 * F * (H0 + 4) + BWD_ROWS * (F + 1) + BWD_ROWS * (2 * H0 + 4)) * sizeof(float);
}

}  // namespace

bool pq_supported(int F, int H0) {
  return F >= 1 && F <= 64 && H0 >= 8 && H0 <= 128 && H0 % 8 == 0 && pq_fwd_smem(F, H0) <= 227 * 1024;
}

int launch_pq_fwd(const float* x, int ldx, const float* W0, int ldw, const float* b0, float* P, float* Q, int BN,
                  int F, int H0, cudaStream_t stream, bool p_tiled) {
  MPG_CHECK((reinterpret_cast<uintptr_t>(P) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0,
            "pq_fwd: P / Q must be 16-byte aligned");
  const size_t smem = pq_fwd_smem(F, H0);
  MPG_CUDA(cudaFuncSetAttribute(pq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PqFwdArgs a{x, ldx, W0, ldw, b0, P, Q, BN, F, H0, p_tiled ? 1 : 0, nullptr, 0};
  pq_fwd_kernel<<<2 * cdiv(BN, PQ_ROWS), PQ_NT, smem, stream>>>(a);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_pq_bwd(const float* dP, const float* dQ, const float* x, int ldx, const float* W0, int ldw, float* dx,
                  int lddx, float* dW0, float* db0, int BN, int F, int H0, cudaStream_t stream, bool p_tiled, bool tf32,
                  const int* cmap, int ctiles_max) {
  MPG_CHECK(cmap == nullptr || (tf32 && H0 % 16 == 0 && p_tiled), "pq_bwd: receiver compaction needs the tensor-core form");
  if (tf32 && H0 % 16 == 0) {
    const size_t smem = bwd_tc_smem(F, H0);
    MPG_CUDA(cudaFuncSetAttribute(pq_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = cmap != nullptr ? ctiles_max * (128 / BWD_ROWS) : cdiv(BN, BWD_ROWS);
    pq_bwd_tc_kernel<<<grid, NT, smem, stream>>>(dP, dQ, x, ldx, W0, ldw, dx, lddx, dW0, db0, BN, F, H0, p_tiled ? 1 : 0,
                                                 cmap, ctiles_max);
    MPG_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = bwd_smem(F, H0);
  MPG_CUDA(cudaFuncSetAttribute(pq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pq_bwd_kernel<<<cdiv(BN, BWD_ROWS), NT, smem, stream>>>(dP, dQ, x, ldx, W0, ldw, dx, lddx, dW0, db0, BN, F, H0, p_tiled ? 1 : 0);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
