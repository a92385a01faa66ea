// Generic fused edge network (any layer widths <= 256, all edge-feature modes), fp32 SIMT.
//
// One CTA per receiver (b, i).  Pair rows (b, i, j) are produced 32 senders at a time in shared
// memory, pushed through fe's three layers and reduced over j on chip: the [B*N*N, H] edge tensors
// never reach HBM.  Backward recomputes the activations per chunk.  This is the reference-accuracy
// path (fp32 accumulate, fp32 operands) and the fallback for non-default architectures; the
// tcgen05 kernel in edge_tc.cu covers the default 96/160/192 edge network.
//
// The senders of a receiver are either all N particles of its jet (fully connected) or the K listed neighbours
// (kNN message passing, mpgan/model.py:319-381): `nbr` carries the list, everything else is unchanged.
//
// The backward kernel has a second-order ("tangent") mode for double backward (WGAN-GP, train.py:286-324): fe is
// piecewise linear (leaky-relu + dropout), so the derivative of <u, dx> -- u the cotangent of the first backward's
// dx output -- w.r.t. dagg / mask / weights is a forward pass of the tangent t = J u through the SAME slopes the
// primal pass took, combined with the ordinary backward signals g: dW_l(2nd) = g_l (x) t_{l-1}, d/d(dagg) = sum_j
// mask_j t_2, d/d(mask_j) = sum_i <t_2, dagg_i>.
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
  int b, i, j0, nvalid;   // senders j0 .. j0+nvalid-1 of the receiver's sender list
  uint64_t pair0;         // (b*N + i) * K
  const int* sj;          // shared memory: particle index (inside the jet) of each sender of the chunk
};

// particle index of the chunk's senders: the list entry (kNN) or the position itself (fully connected)
__device__ __forceinline__ void load_senders(const EdgeArgs& a, const ChunkCtx& c, int* sj) {
  for (int r = threadIdx.x; r < R; r += NTHR) {
    int j = c.j0 + min(r, c.nvalid - 1);
    if (a.nbr != nullptr) j = a.nbr[((size_t)c.b * a.N + c.i) * a.K + j];
    sj[r] = j;
  }
}

// multiplier of a masked sender's coordinates in the kNN distance feature (model.py:336-340)
__device__ __forceinline__ float knn_sender_scale(const EdgeArgs& a, const ChunkCtx& c, int j) {
  if (!a.knn_scale || a.mask == nullptr) return 1.f;
  return fmaf(1.f - 1e4f, a.mask[(size_t)c.b * a.N + j], 1e4f);
}

// pair features for the chunk: efs[e*RS + r]
__device__ __forceinline__ void build_ef(const EdgeArgs& a, const ChunkCtx& c, float* efs, float* diffs) {
  if (a.n_ef == 0) return;
  for (int r = threadIdx.x; r < R; r += NTHR) {
    float d2 = 0.f;
    const int j = c.sj[r];
    const float sc = knn_sender_scale(a, c, j);
    const float* xi = a.x + ((size_t)c.b * a.N + c.i) * a.ldx;
    const float* xj = a.x + ((size_t)c.b * a.N + j) * a.ldx;
    for (int e = 0; e < a.nd; ++e) {
      const float d = sc * xj[e] - xi[e];
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
      v = Pi[k] + a.Q[((size_t)c.b * a.N + c.sj[r]) * a.H0 + k];
      for (int e = 0; e < a.n_ef; ++e) v = fmaf(efs[e * RS + r], a.Wef[(size_t)k * a.ldwef + e], v);
      if (a.Lc != nullptr) v += a.Lc[(size_t)((c.pair0 + c.j0 + r) % (uint64_t)a.B) * a.H0 + k];
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

__device__ __forceinline__ void chunk_forward(const EdgeArgs& a, const ChunkCtx& c, int* sj, float* efs, float* diffs,
                                              float* H0s, float* H1s, float* H2s) {
  load_senders(a, c, sj);
  __syncthreads();
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
  return a.mask ? a.mask[(size_t)c.b * a.N + c.sj[r]] : 1.f;
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
  int* sj = reinterpret_cast<int*>(aggs + a.H2);
  const int bi = blockIdx.x;
  ChunkCtx c;
  c.b = bi / a.N; c.i = bi % a.N;
  c.pair0 = (uint64_t)bi * a.K;
  c.sj = sj;
  for (int k = threadIdx.x; k < a.H2; k += NTHR) aggs[k] = 0.f;
  for (c.j0 = 0; c.j0 < a.K; c.j0 += R) {
    c.nvalid = min(R, a.K - c.j0);
    chunk_forward(a, c, sj, efs, diffs, H0s, H1s, H2s);
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

// tangent of one layer: T_out[c][r] = (sum_k T_in[k][r] * Wt[k][c]) * act_grad(Y[c][r])  (no bias: it has no tangent)
template <int U>
__device__ __forceinline__ void layer_tan(const EdgeArgs& a, const ChunkCtx& c, const float* Tin, int K, const float* Wt,
                                          int C, const float* Y, uint32_t stream, float* Tout) {
  const int tr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  float acc[8][U];
  tile_matmul<U>(acc, Tin, K, Wt, C, C, tr, tc);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int col = tc + 64 * u;
    if (col >= C) continue;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = tr * 8 + r;
      float v = 0.f;
      if (row < c.nvalid) v = acc[r][u] * act_grad(a, Y[col * RS + row], stream, c.pair0 + c.j0 + row, col);
      Tout[col * RS + row] = v;
    }
  }
}

template <bool TAN>
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
  float* defs = dPs + ((a.H0 + 3) & ~3);  // [(n_ef+1)][RS]  (tiles stay 16-byte aligned for any layer width)
  int* sj = reinterpret_cast<int*>(defs + (a.n_ef + 1) * RS);   // [R]
  float* T0s = reinterpret_cast<float*>(sj + R);                // tangent mode only
  float* T1s = T0s + (TAN ? a.H0 * RS : 0);
  float* T2s = T1s + (TAN ? a.H1 * RS : 0);
  float* taggs = T2s + (TAN ? a.H2 * RS : 0);                   // [H2]
  const int bi = blockIdx.x;
  ChunkCtx c;
  c.b = bi / a.N; c.i = bi % a.N;
  c.pair0 = (uint64_t)bi * a.K;
  c.sj = sj;
  for (int k = threadIdx.x; k < a.H0; k += NTHR) dPs[k] = 0.f;
  if (TAN)
    for (int k = threadIdx.x; k < a.H2; k += NTHR) taggs[k] = 0.f;
  const float* dAgg = a.dagg + (size_t)bi * a.H2;
  for (c.j0 = 0; c.j0 < a.K; c.j0 += R) {
    c.nvalid = min(R, a.K - c.j0);
    chunk_forward(a, c, sj, efs, diffs, H0s, H1s, H2s);
    if (TAN) {
      // tangent forward through the slopes of the primal pass
      const float* Pti = a.Pt + ((size_t)c.b * a.N + c.i) * a.H0;
      for (int idx = threadIdx.x; idx < a.H0 * R; idx += NTHR) {
        const int r = idx % R, k = idx / R;
        float v = 0.f;
        if (r < c.nvalid)
          v = (Pti[k] + a.Qt[((size_t)c.b * a.N + c.sj[r]) * a.H0 + k]) *
              act_grad(a, H0s[k * RS + r], 0, c.pair0 + c.j0 + r, k);
        T0s[k * RS + r] = v;
      }
      __syncthreads();
      if (a.H1 <= 64) layer_tan<1>(a, c, T0s, a.H0, a.W1t, a.H1, H1s, 1, T1s);
      else if (a.H1 <= 128) layer_tan<2>(a, c, T0s, a.H0, a.W1t, a.H1, H1s, 1, T1s);
      else layer_tan<4>(a, c, T0s, a.H0, a.W1t, a.H1, H1s, 1, T1s);
      __syncthreads();
      if (a.H2 <= 64) layer_tan<1>(a, c, T1s, a.H1, a.W2t, a.H2, H2s, 2, T2s);
      else if (a.H2 <= 128) layer_tan<2>(a, c, T1s, a.H1, a.W2t, a.H2, H2s, 2, T2s);
      else layer_tan<4>(a, c, T1s, a.H1, a.W2t, a.H2, H2s, 2, T2s);
      __syncthreads();
      // d<u,dx>/d(dagg_i) = scale * sum_j mask_j t2_ij;  d<u,dx>/d(mask_j) += scale * <t2_ij, dagg_i>
      for (int k = threadIdx.x; k < a.H2; k += NTHR) {
        float s = 0.f;
        for (int r = 0; r < c.nvalid; ++r) s = fmaf(T2s[k * RS + r], sender_mask(a, c, r), s);
        taggs[k] += s;
      }
      if (a.gmask != nullptr && threadIdx.x < c.nvalid) {
        const int r = threadIdx.x;
        float s = 0.f;
        for (int k = 0; k < a.H2; ++k) s = fmaf(T2s[k * RS + r], dAgg[k], s);
        atomicAdd(a.gmask + (size_t)c.b * a.N + c.sj[r], s * a.out_scale);
      }
    }
    // d agg / d mask_j = scale * <message_ij, dagg_i>  (only when the mask itself needs a gradient)
    if (a.dmask != nullptr && threadIdx.x < c.nvalid) {
      const int r = threadIdx.x;
      float s = 0.f;
      for (int k = 0; k < a.H2; ++k) s = fmaf(H2s[k * RS + r], dAgg[k], s);
      atomicAdd(a.dmask + (size_t)c.b * a.N + c.sj[r], s * a.out_scale);
    }
    __syncthreads();
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
    wgrad(H2s, a.H2, TAN ? T1s : H1s, a.H1, a.dW2, a.H1);
    if (!TAN)
      for (int k = threadIdx.x; k < a.H2; k += NTHR) {
        float s = 0.f;
        for (int r = 0; r < c.nvalid; ++r) s += H2s[k * RS + r];
        if (s != 0.f) atomicAdd(a.db2 + k, s);
      }
    if (a.H1 <= 64) dgrad<1>(a, c, H2s, a.H2, a.W2, a.H1, H1s, 1, dH1s);
    else if (a.H1 <= 128) dgrad<2>(a, c, H2s, a.H2, a.W2, a.H1, H1s, 1, dH1s);
    else dgrad<4>(a, c, H2s, a.H2, a.W2, a.H1, H1s, 1, dH1s);
    __syncthreads();
    wgrad(dH1s, a.H1, TAN ? T0s : H0s, a.H0, a.dW1, a.H0);
    if (!TAN)
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
        if (v != 0.f) {
          atomicAdd(a.dQ + ((size_t)c.b * a.N + c.sj[r]) * a.H0 + k, v);
          if (a.dLc != nullptr) atomicAdd(a.dLc + (size_t)((c.pair0 + c.j0 + r) % (uint64_t)a.B) * a.H0 + k, v);
        }
      }
    }
    if (!TAN && a.n_ef > 0) {
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
      // chain to x through diffs / dist (the sender side carries the kNN scale of a masked sender)
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
          const int j = c.sj[r];
          atomicAdd(a.dx_ef + ((size_t)c.b * a.N + j) * a.F + e, gd * knn_sender_scale(a, c, j));
          atomicAdd(a.dx_ef + ((size_t)c.b * a.N + c.i) * a.F + e, -gd);
        }
      }
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < a.H0; k += NTHR) a.dP[(size_t)bi * a.H0 + k] = dPs[k];
  if (TAN)
    for (int k = threadIdx.x; k < a.H2; k += NTHR) a.tagg[(size_t)bi * a.H2 + k] = taggs[k] * a.out_scale;
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

size_t edge_generic_smem(const EdgeArgs& a, bool bwd, bool tangent) {
  size_t f = (size_t)(a.H0 + a.H1 + a.H2) * RS + (size_t)(a.n_ef + 1) * RS + (size_t)(a.nd + 1) * RS + R;
  if (bwd) f += (size_t)(a.H1 + a.H0) * RS + ((a.H0 + 3) & ~3) + (size_t)(a.n_ef + 1) * RS;
  else f += a.H2;
  if (tangent) f += (size_t)(a.H0 + a.H1 + a.H2) * RS + a.H2;
  return f * sizeof(float);
}

int launch_transpose(const float* in, int rows, int cols, float* out, cudaStream_t stream) {
  const int n = rows * cols;
  transpose_kernel<<<cdiv(n, 256), 256, 0, stream>>>(in, rows, cols, out);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_edge_generic(const EdgeArgs& a0, bool bwd, cudaStream_t stream) {
  EdgeArgs a = a0;
  if (a.nbr == nullptr) a.K = a.N;
  MPG_CHECK(a.H0 <= 256 && a.H1 <= 256 && a.H2 <= 256, "edge layer widths must be <= 256");
  MPG_CHECK(a.K > 0, "edge network: no senders");
  const size_t smem = edge_generic_smem(a, bwd);
  MPG_CHECK(smem <= 227 * 1024, "edge network too wide for shared memory (%zu B)", smem);
  if (bwd) {
    MPG_CUDA(cudaFuncSetAttribute(edge_bwd_generic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_bwd_generic<false><<<a.B * a.N, NTHR, smem, stream>>>(a);
  } else {
    MPG_CUDA(cudaFuncSetAttribute(edge_fwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    edge_fwd_generic<<<a.B * a.N, NTHR, smem, stream>>>(a);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_edge_generic_tangent(const EdgeArgs& a0, cudaStream_t stream) {
  EdgeArgs a = a0;
  if (a.nbr == nullptr) a.K = a.N;
  MPG_CHECK(a.H0 <= 256 && a.H1 <= 256 && a.H2 <= 256, "edge layer widths must be <= 256");
  MPG_CHECK(a.n_ef == 0, "second-order edge products are implemented for networks without pair features (pos_diffs)");
  MPG_CHECK(a.Pt && a.Qt && a.tagg, "edge tangent: null operand");
  const size_t smem = edge_generic_smem(a, true, true);
  MPG_CHECK(smem <= 227 * 1024, "edge network too wide for the second-order kernel's shared memory (%zu B)", smem);
  MPG_CUDA(cudaFuncSetAttribute(edge_bwd_generic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edge_bwd_generic<true><<<a.B * a.N, NTHR, smem, stream>>>(a);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
