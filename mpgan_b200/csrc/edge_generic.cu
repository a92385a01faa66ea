// Generic fused edge network (any layer widths <= 256, all edge-feature modes), fp32 SIMT.
//
// One CTA per receiver (b, i).  Pair rows (b, i, j) are produced 32 senders at a time in shared
// memory, pushed through fe's three layers and reduced over j on chip: the [B*N*N, H] edge tensors
// never reach HBM.  Backward recomputes the activations per chunk.  This is the reference-accuracy
// path (fp32 accumulate, fp32 operands) and the fallback for non-default architectures; the
// tcgen05 kernel in edge_tc.cu covers the default 96/160/192 edge network.
//
// Reference semantics: mpgan/model.py:256-267 (fe, mask on the sender axis, sum/mean over senders),
// :284-317 (pair features), LinearNet :77-83 (Linear -> leaky_relu -> Dropout).
#include "edge.cuh"

namespace mpg {
namespace {

constexpr int R = 32;     // senders per chunk
constexpr int RS = 36;    // smem row stride of a [channel][sender] tile
constexpr int NTHR = 256;

// out[r][c] = sum_k in[k][r] * W[k*ldw + c]  for r in this thread's 8 rows, c = tc + 64u
template <int U>
__device__ __forceinline__ void tile_matmul(float (&acc)[8][U], const float* __restrict__ in, int K,
                                            const float* __restrict__ W, int ldw, int C, int tr, int tc) {
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int u = 0; u < U; ++u) acc[r][u] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(in + k * RS + tr * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(in + k * RS + tr * 8 + 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = tc + 64 * u;
      w[u] = c < C ? __ldg(W + (size_t)k * ldw + c) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int u = 0; u < U; ++u) acc[r][u] = fmaf(a[r], w[u], acc[r][u]);
  }
}

// keep decision of pair-row element (layer, col): the packed one-draw-per-quarter-row layout the tcgen05
// kernels use when it applies (p == 0.5), else the generic per-element stream
__device__ __forceinline__ bool edge_keep(const EdgeArgs& a, uint32_t layer, uint64_t pair, int col) {
  if (a.drop.half && edge_drop_packed_ok(a.H0, a.H1, a.H2))
    return edge_drop_keep(a.drop.seed, pair, (int)layer, col, a.H0, a.H1, a.H2);
  return drop_keep(a.drop, layer, pair, (uint32_t)col);
}

struct ChunkCtx {
  int b, i, j0, nvalid;   // senders j0 .. j0+nvalid-1
  uint64_t pair0;         // (b*N + i)*N
};

// pair features for the chunk: efs[e*RS + r]
__device__ __forceinline__ void build_ef(const EdgeArgs& a, const ChunkCtx& c, float* efs, float* diffs) {
  if (a.n_ef == 0) return;
  for (int r = threadIdx.x; r < R; r += NTHR) {
    float d2 = 0.f;
    const float* xi = a.x + ((size_t)c.b * a.N + c.i) * a.ldx;
    const float* xj = a.x + ((size_t)c.b * a.N + min(c.j0 + r, a.N - 1)) * a.ldx;
    for (int e = 0; e < a.nd; ++e) {
      const float d = xj[e] - xi[e];
      diffs[e * RS + r] = d;
      const float de = d + 1e-12f;           // eps per component before the norm (model.py:304)
      d2 += de * de;
    }
    int col = 0;
    if (a.ef_mode & 2)
      for (int e = 0; e < a.nd; ++e) efs[(col++) * RS + r] = diffs[e * RS + r];
    if (a.ef_mode & 1) efs[(col++) * RS + r] = sqrtf(d2);
  }
}

// H0[k][r] = drop(lrelu(P_i[k] + Q_j[k] + sum_e ef_e * Wef[k][e]))
__device__ __forceinline__ void build_h0(const EdgeArgs& a, const ChunkCtx& c, const float* efs, float* H0s) {
  const float* Pi = a.P + ((size_t)c.b * a.N + c.i) * a.H0;
  for (int idx = threadIdx.x; idx < a.H0 * R; idx += NTHR) {
    const int r = idx % R, k = idx / R;
    float v = 0.f;
    if (r < c.nvalid) {
      v = Pi[k] + a.Q[((size_t)c.b * a.N + c.j0 + r) * a.H0 + k];
      for (int e = 0; e < a.n_ef; ++e) v = fmaf(efs[e * RS + r], a.Wef[(size_t)k * a.ldwef + e], v);
      v = lrelu(v, a.alpha);
      if (a.drop.p > 0.f) v = edge_keep(a, 0, c.pair0 + c.j0 + r, k) ? v * a.drop.scale : 0.f;
    }
    H0s[k * RS + r] = v;
  }
}

template <int U>
__device__ __forceinline__ void layer_fwd(const EdgeArgs& a, const ChunkCtx& c, const float* in, int K,
                                          const float* Wt, const float* bias, int C, uint32_t stream,
                                          float* out) {
  const int tr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  float acc[8][U];
  tile_matmul<U>(acc, in, K, Wt, C, C, tr, tc);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int col = tc + 64 * u;
    if (col >= C) continue;
    const float bv = bias[col];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = tr * 8 + r;
      float v = lrelu(acc[r][u] + bv, a.alpha);
      if (a.drop.p > 0.f) v = edge_keep(a, stream, c.pair0 + c.j0 + row, col) ? v * a.drop.scale : 0.f;
      out[col * RS + row] = row < c.nvalid ? v : 0.f;
    }
  }
}

__device__ __forceinline__ void chunk_forward(const EdgeArgs& a, const ChunkCtx& c, float* efs, float* diffs,
                                              float* H0s, float* H1s, float* H2s) {
  build_ef(a, c, efs, diffs);
  __syncthreads();
  build_h0(a, c, efs, H0s);
  __syncthreads();
  if (a.H1 <= 64) layer_fwd<1>(a, c, H0s, a.H0, a.W1t, a.b1, a.H1, 1, H1s);
  else if (a.H1 <= 128) layer_fwd<2>(a, c, H0s, a.H0, a.W1t, a.b1, a.H1, 1, H1s);
  else layer_fwd<4>(a, c, H0s, a.H0, a.W1t, a.b1, a.H1, 1, H1s);
  __syncthreads();
  if (a.H2 <= 64) layer_fwd<1>(a, c, H1s, a.H1, a.W2t, a.b2, a.H2, 2, H2s);
  else if (a.H2 <= 128) layer_fwd<2>(a, c, H1s, a.H1, a.W2t, a.b2, a.H2, 2, H2s);
  else layer_fwd<4>(a, c, H1s, a.H1, a.W2t, a.b2, a.H2, 2, H2s);
  __syncthreads();
}

__device__ __forceinline__ float sender_mask(const EdgeArgs& a, const ChunkCtx& c, int r) {
  if (r >= c.nvalid) return 0.f;
  return a.mask ? a.mask[(size_t)c.b * a.N + c.j0 + r] : 1.f;
}

__global__ void __launch_bounds__(NTHR) edge_fwd_generic(EdgeArgs a) {
  resolve_seed(a.drop);
  extern __shared__ __align__(16) float sm[];
  float* H0s = sm;
  float* H1s = H0s + a.H0 * RS;
  float* H2s = H1s + a.H1 * RS;
  float* efs = H2s + a.H2 * RS;
  float* diffs = efs + (a.n_ef + 1) * RS;
  float* aggs = diffs + (a.nd + 1) * RS;
  const int bi = blockIdx.x;
  ChunkCtx c;
  c.b = bi / a.N; c.i = bi % a.N;
  c.pair0 = (uint64_t)bi * a.N;
  for (int k = threadIdx.x; k < a.H2; k += NTHR) aggs[k] = 0.f;
  for (c.j0 = 0; c.j0 < a.N; c.j0 += R) {
    c.nvalid = min(R, a.N - c.j0);
    chunk_forward(a, c, efs, diffs, H0s, H1s, H2s);
    for (int k = threadIdx.x; k < a.H2; k += NTHR) {
      float s = 0.f;
      for (int r = 0; r < c.nvalid; ++r) s = fmaf(H2s[k * RS + r], sender_mask(a, c, r), s);
      aggs[k] += s;
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < a.H2; k += NTHR) a.agg[(size_t)bi * a.H2 + k] = aggs[k] * a.out_scale;
}

// d(act+dropout)/dz given the stored output y (see common.cuh) for pair-row elements
__device__ __forceinline__ float act_grad(const EdgeArgs& a, float y, uint32_t stream, uint64_t row, int col) {
  float gfac = lrelu_grad_from_out(y, a.alpha);
  if (a.drop.p > 0.f) gfac = edge_keep(a, stream, row, col) ? gfac * a.drop.scale : 0.f;
  return gfac;
}

// dW[c_out][c_in] += sum_r dOut[c_out][r] * In[c_in][r]   (atomic into global, ld = ldw)
__device__ __forceinline__ void wgrad(const float* dOut, int Cout, const float* In, int Cin, float* dW, int ldw) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (int o0 = 0; o0 < Cout; o0 += 64)
    for (int i0 = 0; i0 < Cin; i0 += 64) {
      float acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
      for (int r = 0; r < R; ++r) {
        float dv[4], iv[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int co = o0 + ty + 16 * p, ci = i0 + tx + 16 * p;
          dv[p] = co < Cout ? dOut[co * RS + r] : 0.f;
          iv[p] = ci < Cin ? In[ci * RS + r] : 0.f;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(dv[p], iv[q], acc[p][q]);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int co = o0 + ty + 16 * p, ci = i0 + tx + 16 * q;
          if (co < Cout && ci < Cin && acc[p][q] != 0.f) atomicAdd(dW + (size_t)co * ldw + ci, acc[p][q]);
        }
    }
}

// dIn[c_in][r] = (sum_co dOut[co][r] * W[co][c_in]) * act_grad(In[c_in][r])
template <int U>
__device__ __forceinline__ void dgrad(const EdgeArgs& a, const ChunkCtx& c, const float* dOut, int Cout,
                                      const float* W, int Cin, const float* In, uint32_t stream, float* dIn) {
  const int tr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  float acc[8][U];
  tile_matmul<U>(acc, dOut, Cout, W, Cin, Cin, tr, tc);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int col = tc + 64 * u;
    if (col >= Cin) continue;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = tr * 8 + r;
      float v = 0.f;
      if (row < c.nvalid) v = acc[r][u] * act_grad(a, In[col * RS + row], stream, c.pair0 + c.j0 + row, col);
      dIn[col * RS + row] = v;
    }
  }
}

__global__ void __launch_bounds__(NTHR) edge_bwd_generic(EdgeArgs a) {
  resolve_seed(a.drop);
  extern __shared__ __align__(16) float sm[];
  float* H0s = sm;
  float* H1s = H0s + a.H0 * RS;
  float* H2s = H1s + a.H1 * RS;   // becomes dH2 in place
  float* dH1s = H2s + a.H2 * RS;
  float* dH0s = dH1s + a.H1 * RS;
  float* efs = dH0s + a.H0 * RS;
  float* diffs = efs + (a.n_ef + 1) * RS;
  float* dPs = diffs + (a.nd + 1) * RS;   // [H0]
  float* defs = dPs + a.H0;               // [(n_ef+1)][RS]
  const int bi = blockIdx.x;
  ChunkCtx c;
  c.b = bi / a.N; c.i = bi % a.N;
  c.pair0 = (uint64_t)bi * a.N;
  for (int k = threadIdx.x; k < a.H0; k += NTHR) dPs[k] = 0.f;
  const float* dAgg = a.dagg + (size_t)bi * a.H2;
  for (c.j0 = 0; c.j0 < a.N; c.j0 += R) {
    c.nvalid = min(R, a.N - c.j0);
    chunk_forward(a, c, efs, diffs, H0s, H1s, H2s);
    // dH2 (pre-activation grads of layer 2), in place
    for (int idx = threadIdx.x; idx < a.H2 * R; idx += NTHR) {
      const int r = idx % R, k = idx / R;
      float v = 0.f;
      if (r < c.nvalid)
        v = dAgg[k] * a.out_scale * sender_mask(a, c, r) *
            act_grad(a, H2s[k * RS + r], 2, c.pair0 + c.j0 + r, k);
      H2s[k * RS + r] = v;
    }
    __syncthreads();
    wgrad(H2s, a.H2, H1s, a.H1, a.dW2, a.H1);
    for (int k = threadIdx.x; k < a.H2; k += NTHR) {
      float s = 0.f;
      for (int r = 0; r < c.nvalid; ++r) s += H2s[k * RS + r];
      if (s != 0.f) atomicAdd(a.db2 + k, s);
    }
    if (a.H1 <= 64) dgrad<1>(a, c, H2s, a.H2, a.W2, a.H1, H1s, 1, dH1s);
    else if (a.H1 <= 128) dgrad<2>(a, c, H2s, a.H2, a.W2, a.H1, H1s, 1, dH1s);
    else dgrad<4>(a, c, H2s, a.H2, a.W2, a.H1, H1s, 1, dH1s);
    __syncthreads();
    wgrad(dH1s, a.H1, H0s, a.H0, a.dW1, a.H0);
    for (int k = threadIdx.x; k < a.H1; k += NTHR) {
      float s = 0.f;
      for (int r = 0; r < c.nvalid; ++r) s += dH1s[k * RS + r];
      if (s != 0.f) atomicAdd(a.db1 + k, s);
    }
    if (a.H0 <= 64) dgrad<1>(a, c, dH1s, a.H1, a.W1, a.H0, H0s, 0, dH0s);
    else if (a.H0 <= 128) dgrad<2>(a, c, dH1s, a.H1, a.W1, a.H0, H0s, 0, dH0s);
    else dgrad<4>(a, c, dH1s, a.H1, a.W1, a.H0, H0s, 0, dH0s);
    __syncthreads();
    // dP_i (owned by this CTA), dQ_j (shared across receivers of the jet -> atomics)
    for (int k = threadIdx.x; k < a.H0; k += NTHR) {
      float s = 0.f;
      for (int r = 0; r < c.nvalid; ++r) s += dH0s[k * RS + r];
      dPs[k] += s;
    }
    for (int idx = threadIdx.x; idx < a.H0 * R; idx += NTHR) {
      const int k = idx % a.H0, r = idx / a.H0;
      if (r < c.nvalid) {
        const float v = dH0s[k * RS + r];
        if (v != 0.f) atomicAdd(a.dQ + ((size_t)c.b * a.N + c.j0 + r) * a.H0 + k, v);
      }
    }
    if (a.n_ef > 0) {
      // d ef_e(r) = sum_k dH0[k][r] * Wef[k][e];  dWef[k][e] += sum_r dH0[k][r] * ef_e(r)
      for (int idx = threadIdx.x; idx < a.n_ef * R; idx += NTHR) {
        const int r = idx % R, e = idx / R;
        float s = 0.f;
        for (int k = 0; k < a.H0; ++k) s = fmaf(dH0s[k * RS + r], a.Wef[(size_t)k * a.ldwef + e], s);
        defs[e * RS + r] = s;
      }
      for (int idx = threadIdx.x; idx < a.n_ef * a.H0; idx += NTHR) {
        const int e = idx % a.n_ef, k = idx / a.n_ef;
        float s = 0.f;
        for (int r = 0; r < c.nvalid; ++r) s = fmaf(dH0s[k * RS + r], efs[e * RS + r], s);
        if (s != 0.f) atomicAdd(a.dWef + (size_t)k * a.ldwef + e, s);
      }
      __syncthreads();
      // chain to x through diffs / dist
      for (int idx = threadIdx.x; idx < a.nd * R; idx += NTHR) {
        const int r = idx % R, e = idx / R;
        if (r >= c.nvalid) continue;
        float gd = 0.f;
        int col = 0;
        if (a.ef_mode & 2) { gd += defs[e * RS + r]; col = a.nd; }
        if (a.ef_mode & 1) {
          const float dist = efs[col * RS + r];
          gd += defs[col * RS + r] * (diffs[e * RS + r] + 1e-12f) / dist;
        }
        if (gd != 0.f) {
          atomicAdd(a.dx_ef + ((size_t)c.b * a.N + c.j0 + r) * a.F + e, gd);
          atomicAdd(a.dx_ef + ((size_t)c.b * a.N + c.i) * a.F + e, -gd);
        }
      }
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < a.H0; k += NTHR) a.dP[(size_t)bi * a.H0 + k] = dPs[k];
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  // out[c][r] = in[r][c]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * cols) {
    const int r = idx / cols, c = idx % cols;
    out[(size_t)c * rows + r] = in[idx];
  }
}

}  // namespace

size_t edge_generic_smem(const EdgeArgs& a, bool bwd) {
  size_t f = (size_t)(a.H0 + a.H1 + a.H2) * RS + (size_t)(a.n_ef + 1) * RS + (size_t)(a.nd + 1) * RS;
  if (bwd) f += (size_t)(a.H1 + a.H0) * RS + a.H0 + (size_t)(a.n_ef + 1) * RS;
  else f += a.H2;
  return f * sizeof(float);
}

int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t stream) {
  const int n = rows * cols;
  transpose_kernel<<<cdiv(n, 256), 256, 0, stream>>>(in, rows, cols, out);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_edge_generic(const EdgeArgs& a, bool bwd, cudaStream_t stream) {
  MPG_CHECK(a.H0 <= 256 && a.H1 <= 256 && a.H2 <= 256, "edge layer widths must be <= 256");
  const size_t smem = edge_generic_smem(a, bwd);
  MPG_CHECK(smem <= 227 * 1024, "edge network too wide for shared memory (%zu B)", smem);
  if (bwd) {
    MPG_CUDA(cudaFuncSetAttribute(edge_bwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_bwd_generic<<<a.B * a.N, NTHR, smem, stream>>>(a);
  } else {
    MPG_CUDA(cudaFuncSetAttribute(edge_fwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_fwd_generic<<<a.B * a.N, NTHR, smem, stream>>>(a);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
