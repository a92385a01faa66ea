// Fused multihead-attention block of GAPT (reference MAB.forward, gapt/model.py:124-139, without LayerNorm):
//
//   a   = MultiheadAttention(x, y, y, key mask)       packed in_proj (rows 0:E = Q, E:2E = K, 2E:3E = V), heads of
//                                                      E / heads channels, softmax(q k^T / sqrt(d)), out_proj
//   h   = Dropout(x + a)
//   out = Dropout(h + Dropout(leaky_relu(h Wff^T + bff)))      ff = LinearNet([], E -> E, final_linear = False)
//
// as ONE kernel per direction instead of 4-6 GEMM launches + attention core + 2 residual kernels.  A jet has at most
// 32 rows of E = 64 channels: the whole block for one jet lives in shared memory next to the block's weights (83 KB
// fp32, loaded once per CTA); a persistent CTA walks over jets.  The arithmetic is fp32 SIMT with register tiles
// (8 rows x U columns per thread) -- the block is ~1.5 MFLOP per jet against ~100 KB of saved activations, so what is
// being removed is launch latency and HBM round trips, not tensor work.
//
// Saved for backward (per jet): q [Nq,E], kv [Nk,2E], o [Nq,E] (attention output before out_proj), h, f (ff output).
// The backward recomputes the attention probabilities from q / k, keeps the weight gradients of the whole launch in
// registers (81 accumulators per thread) and leaves them in a per-CTA slab that mab_reduce_kernel sums into the
// parameter gradients.  Dropout masks are functions of (seed, stream, row, column) as everywhere else: streams 48 / 49
// = the two residual dropouts, 16 = the feed-forward layer.
#include "mab.cuh"

namespace mpg {
namespace {

constexpr int E = 64, HEADS = 4, HD = 16, RMAX = 32;
constexpr int RS = 36;                       // row stride of a [channel][row] tile (16-byte aligned rows)
constexpr int NT = 256;
constexpr int W_IN = 3 * E * E, W_SQ = E * E;
constexpr int OFF_WIN = 0, OFF_WOUT = W_IN, OFF_WFF = W_IN + W_SQ, OFF_BIN = W_IN + 2 * W_SQ, OFF_BOUT = OFF_BIN + 3 * E,
              OFF_BFF = OFF_BOUT + E, W_FLOATS = OFF_BFF + E;      // 20800
constexpr int SC = RMAX + 1;                 // row stride of a score matrix

// acc[r][u] = sum_k At[k][8 tr + r] * W[k * ldw + tc + 64 u]
template <int U>
__device__ __forceinline__ void mm(float (&acc)[8][U], const float* __restrict__ At, int K, const float* __restrict__ W,
                                   int ldw) {
  const int tr = threadIdx.x >> 6, tc = threadIdx.x & 63;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int u = 0; u < U; ++u) acc[r][u] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(At + k * RS + tr * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(At + k * RS + tr * 8 + 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = W[k * ldw + tc + 64 * u];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int u = 0; u < U; ++u) acc[r][u] = fmaf(a[r], w[u], acc[r][u]);
  }
}

// tile[c][i] = src[(row0 + i) * ld + c] for i < n, zero for n <= i < 32
__device__ __forceinline__ void load_tile(float* tile, const float* __restrict__ src, size_t row0, int ld, int n, int C) {
  for (int idx = threadIdx.x; idx < RMAX * C; idx += NT) {
    const int i = idx / C, c = idx % C;
    tile[c * RS + i] = i < n ? src[(row0 + i) * ld + c] : 0.f;
  }
}

__device__ __forceinline__ bool key_ignored(const float* key_mask, size_t idx) {
  return key_mask != nullptr && (1.f - key_mask[idx]) != 0.f;   // (1 - mask).bool()  (gapt/model.py:194-202)
}

// warp = head, lane = query: scores -> probabilities (in P[h][i][:]), attention output (registers)
__device__ __forceinline__ void attn_probs(const float* qkvT, float* P, const float* ign, int h, int i, int Nk,
                                           float (&q)[HD]) {
  const float scale = 0.25f;   // 1 / sqrt(16)
#pragma unroll
  for (int c = 0; c < HD; ++c) q[c] = qkvT[(h * HD + c) * RS + i] * scale;
  float* row = P + ((size_t)h * RMAX + i) * SC;
  float mx = -INFINITY;
  for (int j = 0; j < Nk; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) s = fmaf(q[c], qkvT[(E + h * HD + c) * RS + j], s);
    if (ign[j] != 0.f) s = -INFINITY;
    row[j] = s;
    mx = fmaxf(mx, s);
  }
  float sum = 0.f;
  for (int j = 0; j < Nk; ++j) {
    const float p = (mx == -INFINITY) ? 0.f : expf(row[j] - mx);
    row[j] = p;
    sum += p;
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = 0; j < Nk; ++j) row[j] *= inv;
}

__global__ void mab_prepare_kernel(MabArgs a, float* __restrict__ wt) {
  // transposed copies for the forward: WinT[k][n] = Win[n][k] (ld 3E), WoutT, WffT; biases as they are
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < W_IN) {
    const int n = idx / E, k = idx % E;
    wt[OFF_WIN + k * 3 * E + n] = a.w_in[idx];
  } else if (idx < W_IN + W_SQ) {
    const int j = idx - W_IN, n = j / E, k = j % E;
    wt[OFF_WOUT + k * E + n] = a.w_out[j];
  } else if (idx < W_IN + 2 * W_SQ) {
    const int j = idx - W_IN - W_SQ, n = j / E, k = j % E;
    wt[OFF_WFF + k * E + n] = a.w_ff[j];
  } else if (idx < W_FLOATS) {
    const int j = idx - OFF_BIN;
    wt[idx] = j < 3 * E ? a.b_in[j] : (j < 4 * E ? a.b_out[j - 3 * E] : a.b_ff[j - 4 * E]);
  }
}

__global__ void __launch_bounds__(NT, 1) mab_fwd_kernel(MabArgs a, const float* __restrict__ wt) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;                          // transposed weights + biases
  float* xT = W + W_FLOATS;               // [E][RS]
  float* yT = xT + E * RS;
  float* qkvT = yT + E * RS;              // [3E][RS]
  float* oT = qkvT + 3 * E * RS;
  float* hT = oT + E * RS;
  float* P = hT + E * RS;                 // [HEADS][32][SC]
  float* ign = P + HEADS * RMAX * SC;     // [32]
  for (int i = threadIdx.x; i < W_FLOATS; i += NT) W[i] = wt[i];
  DropCfg dres = a.drop_res, dff = a.drop_ff;
  resolve_seed(dres);
  resolve_seed(dff);
  const int tr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  const int Nq = a.Nq, Nk = a.Nk;
  const bool self = a.x == a.y && a.ldx == a.ldy && Nq == Nk;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    load_tile(xT, a.x, (size_t)b * Nq, a.ldx, Nq, E);
    if (!self) load_tile(yT, a.y, (size_t)b * Nk, a.ldy, Nk, E);
    for (int j = threadIdx.x; j < RMAX; j += NT) ign[j] = (j < Nk && key_ignored(a.key_mask, (size_t)b * Nk + j)) ? 1.f : 0.f;
    __syncthreads();
    // ---- packed in_proj: Q from x, K | V from y ---------------------------------------------------------------
    if (self) {
      float acc[8][3];
      mm<3>(acc, xT, E, W + OFF_WIN, 3 * E);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int n = tc + 64 * u;
        const float bv = W[OFF_BIN + n];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = tr * 8 + r;
          const float v = i < Nq ? acc[r][u] + bv : 0.f;
          qkvT[n * RS + i] = v;
          if (i < Nq) {
            if (u == 0) a.q[((size_t)b * Nq + i) * E + n] = v;
            else a.kv[((size_t)b * Nk + i) * 2 * E + n - E] = v;
          }
        }
      }
    } else {
      float acc[8][1];
      mm<1>(acc, xT, E, W + OFF_WIN, 3 * E);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = tr * 8 + r;
        const float v = i < Nq ? acc[r][0] + W[OFF_BIN + tc] : 0.f;
        qkvT[tc * RS + i] = v;
        if (i < Nq) a.q[((size_t)b * Nq + i) * E + tc] = v;
      }
      float acc2[8][2];
      mm<2>(acc2, yT, E, W + OFF_WIN + E, 3 * E);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int n = E + tc + 64 * u;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = tr * 8 + r;
          const float v = i < Nk ? acc2[r][u] + W[OFF_BIN + n] : 0.f;
          qkvT[n * RS + i] = v;
          if (i < Nk) a.kv[((size_t)b * Nk + i) * 2 * E + n - E] = v;
        }
      }
    }
    __syncthreads();
    // ---- attention core: warp = head, lane = query -----------------------------------------------------------
    if (threadIdx.x < HEADS * 32) {
      const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
      float q[HD], acc[HD];
      attn_probs(qkvT, P, ign, h, i, Nk, q);
      const float* row = P + ((size_t)h * RMAX + i) * SC;
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] = 0.f;
      for (int j = 0; j < Nk; ++j) {
        const float p = row[j];
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = fmaf(p, qkvT[(2 * E + h * HD + c) * RS + j], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        const float v = i < Nq ? acc[c] : 0.f;
        oT[(h * HD + c) * RS + i] = v;
        if (i < Nq) a.o[((size_t)b * Nq + i) * E + h * HD + c] = v;
      }
    }
    __syncthreads();
    // ---- out_proj + residual + dropout -> h -----------------------------------------------------------------------
    {
      float acc[8][1];
      mm<1>(acc, oT, E, W + OFF_WOUT, E);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = tr * 8 + r;
        float v = 0.f;
        if (i < Nq) {
          v = xT[tc * RS + i] + acc[r][0] + W[OFF_BOUT + tc];
          if (dres.p > 0.f) v = drop_keep(dres, 48, (uint64_t)b * Nq + i, (uint32_t)tc) ? v * dres.scale : 0.f;
          a.h[((size_t)b * Nq + i) * E + tc] = v;
        }
        hT[tc * RS + i] = v;
      }
    }
    __syncthreads();
    // ---- feed-forward layer (leaky-relu + its dropout), residual, dropout -> out ----------------------------------------
    {
      float acc[8][1];
      mm<1>(acc, hT, E, W + OFF_WFF, E);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = tr * 8 + r;
        if (i < Nq) {
          float f = lrelu(acc[r][0] + W[OFF_BFF + tc], a.alpha);
          if (dff.p > 0.f) f = drop_keep(dff, 16, (uint64_t)b * Nq + i, (uint32_t)tc) ? f * dff.scale : 0.f;
          a.f[((size_t)b * Nq + i) * E + tc] = f;
          float v = hT[tc * RS + i] + f;
          if (dres.p > 0.f) v = drop_keep(dres, 49, (uint64_t)b * Nq + i, (uint32_t)tc) ? v * dres.scale : 0.f;
          a.out[((size_t)b * Nq + i) * E + tc] = v;
        }
      }
    }
  }
}

// dW[n][k] += sum_i Zt[n][i] * At[k][i] for this thread's 4 x 4 strided tile: n = (t >> 4) + 16 a, k = (t & 15) + 16 b
__device__ __forceinline__ void wgrad16(float (&acc)[16], const float* __restrict__ Zt, const float* __restrict__ At) {
  const int n0 = threadIdx.x >> 4, k0 = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < RMAX; i += 4) {
    float4 z[4], v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      z[p] = *reinterpret_cast<const float4*>(Zt + (n0 + 16 * p) * RS + i);
      v[p] = *reinterpret_cast<const float4*>(At + (k0 + 16 * p) * RS + i);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int s = 0; s < 4; ++s)
        acc[4 * p + s] += z[p].x * v[s].x + z[p].y * v[s].y + z[p].z * v[s].z + z[p].w * v[s].w;
  }
}
__device__ __forceinline__ float rowsum32(const float* t) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < RMAX; i += 4) {
    const float4 v = *reinterpret_cast<const float4*>(t + i);
    s += v.x + v.y + v.z + v.w;
  }
  return s;
}

__global__ void __launch_bounds__(NT, 1) mab_bwd_kernel(MabArgs a, MabGrads g) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;                          // weights in the reference layout [out][in]
  float* A = W + W_FLOATS;                // dout -> dh accumulator -> da
  float* Hh = A + E * RS;                 // h
  float* Ff = Hh + E * RS;                // f -> dz -> do
  float* Oo = Ff + E * RS;                // o
  float* xT = Oo + E * RS;
  float* yT = xT + E * RS;
  float* qkvT = yT + E * RS;              // [3E][RS]
  float* dqkvT = qkvT + 3 * E * RS;       // [3E][RS]
  float* P = dqkvT + 3 * E * RS;          // [HEADS][32][SC]
  float* dS = P + HEADS * RMAX * SC;
  float* ign = dS + HEADS * RMAX * SC;
  for (int i = threadIdx.x; i < W_IN; i += NT) W[OFF_WIN + i] = a.w_in[i];
  for (int i = threadIdx.x; i < W_SQ; i += NT) { W[OFF_WOUT + i] = a.w_out[i]; W[OFF_WFF + i] = a.w_ff[i]; }
  DropCfg dres = a.drop_res, dff = a.drop_ff;
  resolve_seed(dres);
  resolve_seed(dff);
  const int tr = threadIdx.x >> 6, tc = threadIdx.x & 63;
  const int Nq = a.Nq, Nk = a.Nk;
  const bool self = a.x == a.y && a.ldx == a.ldy && Nq == Nk;
  float gWq[16], gWk[16], gWv[16], gWo[16], gWf[16], gbin = 0.f, gbo = 0.f, gbf = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) gWq[i] = gWk[i] = gWv[i] = gWo[i] = gWf[i] = 0.f;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    const size_t rq = (size_t)b * Nq, rk = (size_t)b * Nk;
    load_tile(Hh, a.h, rq, E, Nq, E);
    load_tile(Oo, a.o, rq, E, Nq, E);
    load_tile(xT, a.x, rq, a.ldx, Nq, E);
    load_tile(self ? xT : yT, a.y, rk, a.ldy, Nk, E);
    load_tile(qkvT, a.q, rq, E, Nq, E);
    load_tile(qkvT + E * RS, a.kv, rk, 2 * E, Nk, 2 * E);
    for (int j = threadIdx.x; j < RMAX; j += NT) ign[j] = (j < Nk && key_ignored(a.key_mask, rk + j)) ? 1.f : 0.f;
    // dout through the last dropout: A = g = dout * keep49; Ff = dz = g * lrelu'(f) * keep16
    for (int idx = threadIdx.x; idx < RMAX * E; idx += NT) {
      const int i = idx / E, c = idx % E;
      float gv = 0.f, dz = 0.f;
      if (i < Nq) {
        gv = g.dout[(rq + i) * E + c];
        if (dres.p > 0.f) gv = drop_keep(dres, 49, rq + i, (uint32_t)c) ? gv * dres.scale : 0.f;
        const float fv = a.f[(rq + i) * E + c];
        dz = gv * lrelu_grad_from_out(fv, a.alpha);
        if (dff.p > 0.f) dz = drop_keep(dff, 16, rq + i, (uint32_t)c) ? dz * dff.scale : 0.f;
      }
      A[c * RS + i] = gv;
      Ff[c * RS + i] = dz;
    }
    __syncthreads();
    // ---- feed-forward layer: dWff += dz^T h, dbff += sum dz, dh = g + dz Wff ----------------------------------------------
    wgrad16(gWf, Ff, Hh);
    if (threadIdx.x < E) gbf += rowsum32(Ff + threadIdx.x * RS);
    {
      float acc[8][1];
      mm<1>(acc, Ff, E, W + OFF_WFF, E);
      __syncthreads();   // everyone has read Ff (dz) and A before they change
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = tr * 8 + r;
        float v = 0.f;
        if (i < Nq) {   // through the first residual dropout: da = dx_res = (g + dz Wff) * keep48
          v = A[tc * RS + i] + acc[r][0];
          if (dres.p > 0.f) v = drop_keep(dres, 48, rq + i, (uint32_t)tc) ? v * dres.scale : 0.f;
        }
        A[tc * RS + i] = v;
      }
    }
    __syncthreads();
    // ---- out_proj: dWout += da^T o, dbout += sum da, do = da Wout -> Ff ---------------------------------------------------
    wgrad16(gWo, A, Oo);
    if (threadIdx.x < E) gbo += rowsum32(A + threadIdx.x * RS);
    {
      float acc[8][1];
      mm<1>(acc, A, E, W + OFF_WOUT, E);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = tr * 8 + r;
        Ff[tc * RS + i] = i < Nq ? acc[r][0] : 0.f;
      }
    }
    // zero the gradient tile of the projections (rows beyond Nq / Nk must stay zero for the weight gradients)
    for (int idx = threadIdx.x; idx < 3 * E * RS; idx += NT) dqkvT[idx] = 0.f;
    __syncthreads();
    // ---- attention backward -------------------------------------------------------------------------------------
    if (threadIdx.x < HEADS * 32) {   // pass A: warp = head, lane = query
      const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
      float q[HD], dor[HD];
      attn_probs(qkvT, P, ign, h, i, Nk, q);
#pragma unroll
      for (int c = 0; c < HD; ++c) dor[c] = Ff[(h * HD + c) * RS + i];
      float* prow = P + ((size_t)h * RMAX + i) * SC;
      float* drow = dS + ((size_t)h * RMAX + i) * SC;
      float D = 0.f;
      for (int j = 0; j < Nk; ++j) {
        float dp = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) dp = fmaf(dor[c], qkvT[(2 * E + h * HD + c) * RS + j], dp);
        drow[j] = dp;
        D = fmaf(prow[j], dp, D);
      }
      float dq[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = 0.f;
      for (int j = 0; j < Nk; ++j) {
        const float ds = (i < Nq) ? prow[j] * (drow[j] - D) : 0.f;
        drow[j] = ds;
        if (i >= Nq) prow[j] = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) dq[c] = fmaf(ds, qkvT[(E + h * HD + c) * RS + j], dq[c]);
      }
      if (i < Nq)
#pragma unroll
        for (int c = 0; c < HD; ++c) dqkvT[(h * HD + c) * RS + i] = dq[c] * 0.25f;
    }
    __syncthreads();
    if (threadIdx.x < HEADS * 32) {   // pass B: warp = head, lane = key
      const int h = threadIdx.x >> 5, j = threadIdx.x & 31;
      if (j < Nk) {
        float dk[HD], dv[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) dk[c] = dv[c] = 0.f;
        for (int i = 0; i < Nq; ++i) {
          const float ds = dS[((size_t)h * RMAX + i) * SC + j], p = P[((size_t)h * RMAX + i) * SC + j];
#pragma unroll
          for (int c = 0; c < HD; ++c) {
            dk[c] = fmaf(ds, qkvT[(h * HD + c) * RS + i], dk[c]);
            dv[c] = fmaf(p, Ff[(h * HD + c) * RS + i], dv[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < HD; ++c) {
          dqkvT[(E + h * HD + c) * RS + j] = dk[c] * 0.25f;
          dqkvT[(2 * E + h * HD + c) * RS + j] = dv[c];
        }
      }
    }
    __syncthreads();
    // ---- in_proj: dWq += dq^T x, dWk | dWv += dk | dv ^T y, db_in, dx = da + dq Wq, dy = dk Wk + dv Wv -------------------------
    const float* yTt = self ? xT : yT;
    wgrad16(gWq, dqkvT, xT);
    wgrad16(gWk, dqkvT + E * RS, yTt);
    wgrad16(gWv, dqkvT + 2 * E * RS, yTt);
    if (threadIdx.x < 3 * E) gbin += rowsum32(dqkvT + threadIdx.x * RS);
    {
      float accx[8][1], accy[8][1];
      mm<1>(accx, dqkvT, E, W + OFF_WIN, E);
      mm<1>(accy, dqkvT + E * RS, 2 * E, W + OFF_WIN + E * E, E);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = tr * 8 + r;
        const float dxv = A[tc * RS + i] + accx[r][0];
        if (self) {
          if (i < Nq) g.dx[(rq + i) * E + tc] = dxv + accy[r][0];
        } else {
          if (i < Nq) g.dx[(rq + i) * E + tc] = dxv;
          if (i < Nk) g.dy[(rk + i) * E + tc] = accy[r][0];
        }
      }
    }
  }
  // ---- this CTA's weight-gradient partials -> its slab ------------------------------------------------------------------
  if (g.slab != nullptr) {
    float* s = g.slab + (size_t)blockIdx.x * W_FLOATS;
    const int n0 = threadIdx.x >> 4, k0 = threadIdx.x & 15;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = n0 + 16 * p, k = k0 + 16 * q;
        s[OFF_WIN + n * E + k] = gWq[4 * p + q];
        s[OFF_WIN + (E + n) * E + k] = gWk[4 * p + q];
        s[OFF_WIN + (2 * E + n) * E + k] = gWv[4 * p + q];
        s[OFF_WOUT + n * E + k] = gWo[4 * p + q];
        s[OFF_WFF + n * E + k] = gWf[4 * p + q];
      }
    if (threadIdx.x < 3 * E) s[OFF_BIN + threadIdx.x] = gbin;
    if (threadIdx.x < E) { s[OFF_BOUT + threadIdx.x] = gbo; s[OFF_BFF + threadIdx.x] = gbf; }
  }
}

// parameter gradients += sum of the CTAs' slabs
__global__ void mab_reduce_kernel(const float* __restrict__ slabs, int nslabs, MabGrads g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W_FLOATS) return;
  float s = 0.f;
  for (int c = 0; c < nslabs; ++c) s += slabs[(size_t)c * W_FLOATS + i];
  float* dst;
  if (i < OFF_WOUT) dst = g.dw_in + i;
  else if (i < OFF_WFF) dst = g.dw_out + (i - OFF_WOUT);
  else if (i < OFF_BIN) dst = g.dw_ff + (i - OFF_WFF);
  else if (i < OFF_BOUT) dst = g.db_in + (i - OFF_BIN);
  else if (i < OFF_BFF) dst = g.db_out + (i - OFF_BOUT);
  else dst = g.db_ff + (i - OFF_BFF);
  *dst += s;
}


// =====================================================================================================================
// precision 1: the same block with its GEMMs on tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate).  Tiles are
// row-major [row][channel] with a stride of 68 floats and the weights keep the reference layout [out][in] (same
// stride): every fragment load (row = g or t, column = t or g, g = lane / 4, t = lane % 4) then touches 32 distinct
// banks (A / B^T operands) or at worst two per bank (transposed operands).  A [32 x K] x [K x 64] product is 2 x 8
// MMA tiles x K/8 steps = 16 MMAs per warp for K = 64 instead of ~900 FMA-loop instructions per thread.
// =====================================================================================================================
constexpr int LDT = 68;
constexpr int TILE_F = RMAX * LDT;          // floats of one [32][68] tile
constexpr int WT_ROWS = 5 * E;              // Win (192) | Wout (64) | Wff (64) rows of 68

__device__ __forceinline__ uint32_t tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C[32 x 8*NTW*8] = A[32 x K] * W^T,  W [n][k] (B operand column-major = rows of W).  Warp w owns n-tiles
// w*NTW .. w*NTW+NTW-1; acc[mi][nt][..] follows the m16n8 accumulator layout (row g / g+8, columns 2t, 2t+1).
template <int NTW>
__device__ __forceinline__ void mma_AWt(float (&acc)[2][NTW][4], const float* __restrict__ A, const float* __restrict__ W,
                                        int K) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][nt][e] = 0.f;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t a[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const float* ap = A + (mi * 16 + g) * LDT + k0 + t;
      a[mi][0] = tf32(ap[0]); a[mi][1] = tf32(ap[8 * LDT]); a[mi][2] = tf32(ap[4]); a[mi][3] = tf32(ap[8 * LDT + 4]);
    }
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      const float* bp = W + ((warp * NTW + nt) * 8 + g) * LDT + k0 + t;
      const uint32_t b[2] = {__float_as_uint(bp[0]), __float_as_uint(bp[4])};   // weights were rounded at load time
      mma8(acc[0][nt], a[0], b);
      mma8(acc[1][nt], a[1], b);
    }
  }
}
// C[32 x 64] = A[32 x K] * W,  W [k][n] (reduction index = row of W): dX = dZ W.  Warp w owns n-tile w.
__device__ __forceinline__ void mma_AW(float (&acc)[2][1][4], const float* __restrict__ A, const float* __restrict__ W,
                                       int K) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[mi][0][e] = 0.f;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 8) {
    const float* bp = W + (k0 + t) * LDT + warp * 8 + g;
    const uint32_t b[2] = {__float_as_uint(bp[0]), __float_as_uint(bp[4 * LDT])};
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const float* ap = A + (mi * 16 + g) * LDT + k0 + t;
      const uint32_t a[4] = {tf32(ap[0]), tf32(ap[8 * LDT]), tf32(ap[4]), tf32(ap[8 * LDT + 4])};
      mma8(acc[mi][0], a, b);
    }
  }
}
// dW[64 x 64] += Z^T X over the 32 rows: dW[m][n] = sum_i Z[i][m] X[i][n].  Warp w owns n-tile w, all four m-tiles.
__device__ __forceinline__ void mma_ZtX(float (&acc)[4][4], const float* __restrict__ Z, const float* __restrict__ X) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k0 = 0; k0 < RMAX; k0 += 8) {
    const float* bp = X + (k0 + t) * LDT + warp * 8 + g;
    const uint32_t b[2] = {tf32(bp[0]), tf32(bp[4 * LDT])};
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const float* ap = Z + (k0 + t) * LDT + mi * 16 + g;
      const uint32_t a[4] = {tf32(ap[0]), tf32(ap[8]), tf32(ap[4 * LDT]), tf32(ap[4 * LDT + 8])};
      mma8(acc[mi], a, b);
    }
  }
}
// visit the accumulator elements of warp-owned tiles: f(row, col, value&)
template <int NTW, class F>
__device__ __forceinline__ void for_acc(float (&acc)[2][NTW][4], F&& f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        f(mi * 16 + g + 8 * (e >> 1), (warp * NTW + nt) * 8 + 2 * t + (e & 1), acc[mi][nt][e]);
}

// tile[i][c] = src[(row0 + i) * ld + c] for i < n, zero rows beyond
__device__ __forceinline__ void load_rows(float* tile, const float* __restrict__ src, size_t row0, int ld, int n, int C,
                                          int col0 = 0) {
  for (int idx = threadIdx.x; idx < RMAX * C; idx += NT) {
    const int i = idx / C, c = idx % C;
    tile[i * LDT + c] = i < n ? src[(row0 + i) * ld + col0 + c] : 0.f;
  }
}
// the same through cp.async (16 bytes per copy, zero fill for the rows beyond n): every tile of a jet is in flight at
// once and no register holds the data; needs 16-byte aligned rows (ld % 4 == 0, aligned base)
__device__ __forceinline__ void load_rows_async(float* tile, const float* __restrict__ src, size_t row0, int ld, int n,
                                                int col0 = 0) {
  for (int idx = threadIdx.x; idx < RMAX * (E / 4); idx += NT) {
    const int i = idx >> 4, c = (idx & 15) * 4;
    const float* sp = src + (row0 + (i < n ? i : n - 1)) * ld + col0 + c;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + i * LDT + c);
    const int bytes = i < n ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(sp), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ bool rows_aligned(const float* p, int ld) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0;
}

__device__ __forceinline__ void load_weights_tc(float* W, const MabArgs& a) {   // [320][68], TF32-rounded
  const bool vec = ((reinterpret_cast<uintptr_t>(a.w_in) | reinterpret_cast<uintptr_t>(a.w_out) |
                     reinterpret_cast<uintptr_t>(a.w_ff)) & 15) == 0;
  for (int idx = threadIdx.x; idx < WT_ROWS * (E / 4); idx += NT) {
    const int n = idx >> 4, k = (idx & 15) * 4;
    const float* src = n < 3 * E ? a.w_in + n * E + k : (n < 4 * E ? a.w_out + (n - 3 * E) * E + k : a.w_ff + (n - 4 * E) * E + k);
    float4 v;
    if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
    else v = make_float4(src[0], src[1], src[2], src[3]);
    *reinterpret_cast<float4*>(W + n * LDT + k) = make_float4(__uint_as_float(tf32(v.x)), __uint_as_float(tf32(v.y)),
                                                              __uint_as_float(tf32(v.z)), __uint_as_float(tf32(v.w)));
  }
}

// Dropout keep bits of one jet, p == 0.5: ONE Philox draw per (site, row) -- the 64 columns of a row are bits of the
// words x, y of drop_bits128(seed, stream, row, 0) (common.cuh: drop_keep) -- computed by 96 threads into shared memory
// instead of one draw per element in every epilogue.  Sites: 0 = stream 48 (first residual), 1 = stream 16 (ff), 2 = 49.
__device__ __forceinline__ void jet_drop_words(uint2* words, const DropCfg& dres, const DropCfg& dff, size_t row0, int n) {
  if (threadIdx.x < 96) {
    const int site = threadIdx.x >> 5, r = threadIdx.x & 31;
    const DropCfg& d = site == 1 ? dff : dres;
    uint2 w = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    if (d.p > 0.f && d.half && r < n) {
      const u4 b = drop_bits128(d.seed, site == 0 ? 48u : (site == 1 ? 16u : 49u), row0 + r, 0);
      w = make_uint2(b.x, b.y);
    }
    words[site * RMAX + r] = w;
  }
}
__device__ __forceinline__ bool jet_keep(const uint2* words, const DropCfg& d, int site, uint32_t stream, size_t row0, int i,
                                         int c) {
  if (d.half) {
    const uint2 w = words[site * RMAX + i];
    return ((c < 32 ? w.x : w.y) >> (c & 31)) & 1u;
  }
  return drop_keep(d, stream, row0 + i, (uint32_t)c);
}

// attention probabilities for (head h, query i) from row-major Q / K tiles; q scaled
__device__ __forceinline__ void attn_probs_rm(const float* Q, const float* Kt, float* P, const float* ign, int h, int i,
                                              int Nk, float (&q)[HD]) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(Q + i * LDT + h * HD + c);
    q[c] = v.x * 0.25f; q[c + 1] = v.y * 0.25f; q[c + 2] = v.z * 0.25f; q[c + 3] = v.w * 0.25f;
  }
  float* row = P + ((size_t)h * RMAX + i) * SC;
  float mx = -INFINITY;
  for (int j = 0; j < Nk; ++j) {
    const float* kr = Kt + j * LDT + h * HD;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 2) { s0 = fmaf(q[c], kr[c], s0); s1 = fmaf(q[c + 1], kr[c + 1], s1); }
    float sv = s0 + s1;
    if (ign[j] != 0.f) sv = -INFINITY;
    row[j] = sv;
    mx = fmaxf(mx, sv);
  }
  float sum = 0.f;
  for (int j = 0; j < Nk; ++j) {
    const float p = (mx == -INFINITY) ? 0.f : __expf(row[j] - mx);
    row[j] = p;
    sum += p;
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = 0; j < Nk; ++j) row[j] *= inv;
}

// the same with the keys of one (head, query) split over a lane pair (half = lane & 1 takes j = half, half + 2, ...):
// all eight warps work on the attention core and every dependent chain is half as long
__device__ __forceinline__ void attn_probs_rm2(const float* Q, const float* Kt, float* P, const float* ign, int h, int i,
                                               int half, int Nk, float (&q)[HD]) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(Q + i * LDT + h * HD + c);
    q[c] = v.x * 0.25f; q[c + 1] = v.y * 0.25f; q[c + 2] = v.z * 0.25f; q[c + 3] = v.w * 0.25f;
  }
  float* row = P + ((size_t)h * RMAX + i) * SC;
  float mx = -INFINITY;
  for (int j = half; j < Nk; j += 2) {
    const float* kr = Kt + j * LDT + h * HD;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 2) { s0 = fmaf(q[c], kr[c], s0); s1 = fmaf(q[c + 1], kr[c + 1], s1); }
    float sv = s0 + s1;
    if (ign[j] != 0.f) sv = -INFINITY;
    row[j] = sv;
    mx = fmaxf(mx, sv);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
  for (int j = half; j < Nk; j += 2) {
    const float p = (mx == -INFINITY) ? 0.f : __expf(row[j] - mx);
    row[j] = p;
    sum += p;
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = half; j < Nk; j += 2) row[j] *= inv;
}

__global__ void __launch_bounds__(NT, 1) mab_fwd_tc_kernel(MabArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;                           // [320][68]
  float* bias = W + WT_ROWS * LDT;         // [320]
  float* Xb = bias + 5 * E;                // x / y tiles, double buffered: the next jet's rows arrive (cp.async) under
  float* Yb = Xb + 2 * TILE_F;             // this jet's arithmetic
  float* Q = Yb + 2 * TILE_F;
  float* Kt = Q + TILE_F;
  float* V = Kt + TILE_F;
  float* O = V + TILE_F;
  float* Hh = O + TILE_F;
  float* P = Hh + TILE_F;                  // [HEADS][32][SC]
  float* ign = P + HEADS * RMAX * SC;
  uint2* dwords = reinterpret_cast<uint2*>(ign + RMAX);   // [3][32]
  load_weights_tc(W, a);
  for (int i = threadIdx.x; i < 5 * E; i += NT) bias[i] = i < 3 * E ? a.b_in[i] : (i < 4 * E ? a.b_out[i - 3 * E] : a.b_ff[i - 4 * E]);
  DropCfg dres = a.drop_res, dff = a.drop_ff;
  resolve_seed(dres);
  resolve_seed(dff);
  const int Nq = a.Nq, Nk = a.Nk;
  const bool self = a.x == a.y && a.ldx == a.ldy && Nq == Nk;
  const bool async = rows_aligned(a.x, a.ldx) && rows_aligned(a.y, a.ldy);
  auto fetch = [&](int b, int buf) {
    if (async) {
      load_rows_async(Xb + buf * TILE_F, a.x, (size_t)b * Nq, a.ldx, Nq);
      if (!self) load_rows_async(Yb + buf * TILE_F, a.y, (size_t)b * Nk, a.ldy, Nk);
    } else {
      load_rows(Xb + buf * TILE_F, a.x, (size_t)b * Nq, a.ldx, Nq, E);
      if (!self) load_rows(Yb + buf * TILE_F, a.y, (size_t)b * Nk, a.ldy, Nk, E);
    }
  };
  int buf = 0;
  fetch(blockIdx.x, 0);
  for (int b = blockIdx.x; b < a.B; b += gridDim.x, buf ^= 1) {
    const size_t rq = (size_t)b * Nq, rk = (size_t)b * Nk;
    cp_async_wait_all();
    __syncthreads();                      // this jet's rows have landed; the previous jet is finished everywhere
    if (b + (int)gridDim.x < a.B) fetch(b + gridDim.x, buf ^ 1);
    for (int j = threadIdx.x; j < RMAX; j += NT) ign[j] = (j < Nk && key_ignored(a.key_mask, rk + j)) ? 1.f : 0.f;
    jet_drop_words(dwords, dres, dff, rq, Nq);
    float* X = Xb + buf * TILE_F;
    const float* Ys = self ? X : Yb + buf * TILE_F;
    {   // Q = x Wq^T + bq
      float acc[2][1][4];
      mma_AWt<1>(acc, X, W, E);
      for_acc<1>(acc, [&](int i, int n, float& v) {
        const float o = i < Nq ? v + bias[n] : 0.f;
        Q[i * LDT + n] = o;
        if (i < Nq) a.q[(rq + i) * E + n] = o;
      });
    }
    {   // K | V = y Wkv^T + bkv
      float acc[2][2][4];
      mma_AWt<2>(acc, Ys, W + E * LDT, E);
      for_acc<2>(acc, [&](int i, int n, float& v) {
        const float o = i < Nk ? v + bias[E + n] : 0.f;
        (n < E ? Kt : V)[i * LDT + (n & (E - 1))] = o;
        if (i < Nk) a.kv[(rk + i) * 2 * E + n] = o;
      });
    }
    __syncthreads();
    {   // attention core: 64 threads per head, a lane pair per query (keys split even / odd)
      const int h = threadIdx.x >> 6, i = (threadIdx.x >> 1) & 31, half = threadIdx.x & 1;
      float q[HD], acc[HD];
      attn_probs_rm2(Q, Kt, P, ign, h, i, half, Nk, q);
      const float* row = P + ((size_t)h * RMAX + i) * SC;
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] = 0.f;
      for (int j = half; j < Nk; j += 2) {
        const float p = row[j];
        const float* vr = V + j * LDT + h * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = fmaf(p, vr[c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
#pragma unroll
      for (int c = 0; c < HD / 2; ++c) {   // each lane of the pair stores half of the head's channels
        const int cc = half * (HD / 2) + c;
        O[i * LDT + h * HD + cc] = i < Nq ? acc[cc] : 0.f;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < Nq * E; idx += NT) a.o[rq * E + idx] = O[(idx >> 6) * LDT + (idx & 63)];
    {   // h = Dropout(x + o Wout^T + bout)
      float acc[2][1][4];
      mma_AWt<1>(acc, O, W + 3 * E * LDT, E);
      for_acc<1>(acc, [&](int i, int n, float& v) {
        float o = 0.f;
        if (i < Nq) {
          o = X[i * LDT + n] + v + bias[3 * E + n];
          if (dres.p > 0.f) o = jet_keep(dwords, dres, 0, 48, rq, i, n) ? o * dres.scale : 0.f;
          a.h[(rq + i) * E + n] = o;
        }
        Hh[i * LDT + n] = o;
      });
    }
    __syncthreads();
    {   // out = Dropout(h + Dropout(lrelu(h Wff^T + bff)))
      float acc[2][1][4];
      mma_AWt<1>(acc, Hh, W + 4 * E * LDT, E);
      for_acc<1>(acc, [&](int i, int n, float& v) {
        if (i < Nq) {
          float f = lrelu(v + bias[4 * E + n], a.alpha);
          if (dff.p > 0.f) f = jet_keep(dwords, dff, 1, 16, rq, i, n) ? f * dff.scale : 0.f;
          a.f[(rq + i) * E + n] = f;
          float o = Hh[i * LDT + n] + f;
          if (dres.p > 0.f) o = jet_keep(dwords, dres, 2, 49, rq, i, n) ? o * dres.scale : 0.f;
          a.out[(rq + i) * E + n] = o;
        }
      });
    }
  }
}

__global__ void __launch_bounds__(NT, 1) mab_bwd_tc_kernel(MabArgs a, MabGrads g) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;                           // [320][68] reference layout, TF32-rounded
  float* A = W + WT_ROWS * LDT;            // dout -> dh accumulator -> da
  float* Hh = A + TILE_F;
  float* Ff = Hh + TILE_F;                 // dz -> do
  float* Oo = Ff + TILE_F;
  float* X = Oo + TILE_F;
  float* Y = X + TILE_F;
  float* Q = Y + TILE_F;
  float* Kt = Q + TILE_F;
  float* V = Kt + TILE_F;
  float* dQ = V + TILE_F;
  float* dK = dQ + TILE_F;
  float* dV = dK + TILE_F;
  float* P = dV + TILE_F;
  float* dS = P + HEADS * RMAX * SC;
  float* ign = dS + HEADS * RMAX * SC;
  uint2* dwords = reinterpret_cast<uint2*>(ign + RMAX);   // [3][32]
  load_weights_tc(W, a);
  DropCfg dres = a.drop_res, dff = a.drop_ff;
  resolve_seed(dres);
  resolve_seed(dff);
  const int Nq = a.Nq, Nk = a.Nk;
  const bool self = a.x == a.y && a.ldx == a.ldy && Nq == Nk;
  const bool async = rows_aligned(a.x, a.ldx) && rows_aligned(a.y, a.ldy) && rows_aligned(g.dout, E) &&
                     rows_aligned(a.q, E) && rows_aligned(a.kv, E) && rows_aligned(a.o, E) && rows_aligned(a.h, E) &&
                     rows_aligned(a.f, E);
  float gWq[4][4], gWk[4][4], gWv[4][4], gWo[4][4], gWf[4][4], gbin = 0.f, gbo = 0.f, gbf = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) gWq[i][e] = gWk[i][e] = gWv[i][e] = gWo[i][e] = gWf[i][e] = 0.f;
  auto colsum = [&](const float* T, int c) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < RMAX; ++i) s += T[i * LDT + c];
    return s;
  };

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    const size_t rq = (size_t)b * Nq, rk = (size_t)b * Nk;
    if (async) {   // every tile of the jet in flight at once (saved activations are dense, 256-byte rows)
      load_rows_async(Hh, a.h, rq, E, Nq);
      load_rows_async(Oo, a.o, rq, E, Nq);
      load_rows_async(X, a.x, rq, a.ldx, Nq);
      if (!self) load_rows_async(Y, a.y, rk, a.ldy, Nk);
      load_rows_async(Q, a.q, rq, E, Nq);
      load_rows_async(Kt, a.kv, rk, 2 * E, Nk, 0);
      load_rows_async(V, a.kv, rk, 2 * E, Nk, E);
      load_rows_async(A, g.dout, rq, E, Nq);      // raw dout and f: transformed in place below
      load_rows_async(Ff, a.f, rq, E, Nq);
    } else {
      load_rows(Hh, a.h, rq, E, Nq, E);
      load_rows(Oo, a.o, rq, E, Nq, E);
      load_rows(X, a.x, rq, a.ldx, Nq, E);
      if (!self) load_rows(Y, a.y, rk, a.ldy, Nk, E);
      load_rows(Q, a.q, rq, E, Nq, E);
      load_rows(Kt, a.kv, rk, 2 * E, Nk, E, 0);
      load_rows(V, a.kv, rk, 2 * E, Nk, E, E);
      load_rows(A, g.dout, rq, E, Nq, E);
      load_rows(Ff, a.f, rq, E, Nq, E);
    }
    for (int j = threadIdx.x; j < RMAX; j += NT) ign[j] = (j < Nk && key_ignored(a.key_mask, rk + j)) ? 1.f : 0.f;
    jet_drop_words(dwords, dres, dff, rq, Nq);
    cp_async_wait_all();
    __syncthreads();
    for (int idx = threadIdx.x; idx < RMAX * E; idx += NT) {   // A = dout * keep49, Ff = dz = A * lrelu'(f) * keep16
      const int i = idx >> 6, c = idx & 63;
      float gv = 0.f, dz = 0.f;
      if (i < Nq) {
        gv = A[i * LDT + c];
        if (dres.p > 0.f) gv = jet_keep(dwords, dres, 2, 49, rq, i, c) ? gv * dres.scale : 0.f;
        dz = gv * lrelu_grad_from_out(Ff[i * LDT + c], a.alpha);
        if (dff.p > 0.f) dz = jet_keep(dwords, dff, 1, 16, rq, i, c) ? dz * dff.scale : 0.f;
      }
      A[i * LDT + c] = gv;
      Ff[i * LDT + c] = dz;
    }
    __syncthreads();
    const float* Ys = self ? X : Y;
    // ---- feed-forward: dWff += dz^T h, dbff, da = (g + dz Wff) * keep48 -----------------------------------------------
    mma_ZtX(gWf, Ff, Hh);
    if (threadIdx.x < E) gbf += colsum(Ff, threadIdx.x);
    {
      float acc[2][1][4];
      mma_AW(acc, Ff, W + 4 * E * LDT, E);
      for_acc<1>(acc, [&](int i, int n, float& v) {
        float o = 0.f;
        if (i < Nq) {
          o = A[i * LDT + n] + v;
          if (dres.p > 0.f) o = jet_keep(dwords, dres, 0, 48, rq, i, n) ? o * dres.scale : 0.f;
        }
        A[i * LDT + n] = o;
      });
    }
    __syncthreads();
    // ---- out_proj: dWout += da^T o, dbout, do = da Wout -> Ff --------------------------------------------------------------
    mma_ZtX(gWo, A, Oo);
    if (threadIdx.x < E) gbo += colsum(A, threadIdx.x);
    {
      float acc[2][1][4];
      mma_AW(acc, A, W + 3 * E * LDT, E);
      __syncthreads();   // dz (Ff) has been read by everyone (dWff product above, this product's operand is A)
      for_acc<1>(acc, [&](int i, int n, float& v) { Ff[i * LDT + n] = i < Nq ? v : 0.f; });
    }
    for (int idx = threadIdx.x; idx < 3 * TILE_F; idx += NT) dQ[idx] = 0.f;   // dQ | dK | dV are adjacent
    __syncthreads();
    // ---- attention backward -------------------------------------------------------------------------------------------
    {   // pass A: a lane pair per (head, query), keys split even / odd
      const int h = threadIdx.x >> 6, i = (threadIdx.x >> 1) & 31, half = threadIdx.x & 1;
      float q[HD], dor[HD];
      attn_probs_rm2(Q, Kt, P, ign, h, i, half, Nk, q);
#pragma unroll
      for (int c = 0; c < HD; ++c) dor[c] = Ff[i * LDT + h * HD + c];
      float* prow = P + ((size_t)h * RMAX + i) * SC;
      float* drow = dS + ((size_t)h * RMAX + i) * SC;
      float D = 0.f;
      for (int j = half; j < Nk; j += 2) {
        const float* vr = V + j * LDT + h * HD;
        float dp = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) dp = fmaf(dor[c], vr[c], dp);
        drow[j] = dp;
        D = fmaf(prow[j], dp, D);
      }
      D += __shfl_xor_sync(0xffffffffu, D, 1);
      float dq[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = 0.f;
      for (int j = half; j < Nk; j += 2) {
        const float ds = (i < Nq) ? prow[j] * (drow[j] - D) : 0.f;
        drow[j] = ds;
        if (i >= Nq) prow[j] = 0.f;
        const float* kr = Kt + j * LDT + h * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) dq[c] = fmaf(ds, kr[c], dq[c]);
      }
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] += __shfl_xor_sync(0xffffffffu, dq[c], 1);
      if (i < Nq)
#pragma unroll
        for (int c = 0; c < HD / 2; ++c) {
          const int cc = half * (HD / 2) + c;
          dQ[i * LDT + h * HD + cc] = dq[cc] * 0.25f;
        }
    }
    __syncthreads();
    {   // pass B: a lane pair per (head, key), queries split even / odd
      const int h = threadIdx.x >> 6, j = (threadIdx.x >> 1) & 31, half = threadIdx.x & 1;
      float dk[HD], dv[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) dk[c] = dv[c] = 0.f;
      if (j < Nk)
        for (int i = half; i < Nq; i += 2) {
          const float ds = dS[((size_t)h * RMAX + i) * SC + j], p = P[((size_t)h * RMAX + i) * SC + j];
          const float* qr = Q + i * LDT + h * HD;
          const float* dr = Ff + i * LDT + h * HD;
#pragma unroll
          for (int c = 0; c < HD; ++c) {
            dk[c] = fmaf(ds, qr[c], dk[c]);
            dv[c] = fmaf(p, dr[c], dv[c]);
          }
        }
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        dk[c] += __shfl_xor_sync(0xffffffffu, dk[c], 1);
        dv[c] += __shfl_xor_sync(0xffffffffu, dv[c], 1);
      }
      if (j < Nk)
#pragma unroll
        for (int c = 0; c < HD / 2; ++c) {
          const int cc = half * (HD / 2) + c;
          dK[j * LDT + h * HD + cc] = dk[cc] * 0.25f;
          dV[j * LDT + h * HD + cc] = dv[cc];
        }
    }
    __syncthreads();
    // ---- in_proj: weight / bias gradients, dx = da + dq Wq, dy = dk Wk + dv Wv ----------------------------------------------
    mma_ZtX(gWq, dQ, X);
    mma_ZtX(gWk, dK, Ys);
    mma_ZtX(gWv, dV, Ys);
    if (threadIdx.x < 3 * E) gbin += colsum(threadIdx.x < E ? dQ : (threadIdx.x < 2 * E ? dK : dV), threadIdx.x & (E - 1));
    {
      float ax[2][1][4], ak[2][1][4], av[2][1][4];
      mma_AW(ax, dQ, W, E);
      mma_AW(ak, dK, W + E * LDT, E);
      mma_AW(av, dV, W + 2 * E * LDT, E);
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gg = lane >> 2, t = lane & 3;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = mi * 16 + gg + 8 * (e >> 1), n = warp * 8 + 2 * t + (e & 1);
          const float dxv = A[i * LDT + n] + ax[mi][0][e], dyv = ak[mi][0][e] + av[mi][0][e];
          if (self) {
            if (i < Nq) g.dx[(rq + i) * E + n] = dxv + dyv;
          } else {
            if (i < Nq) g.dx[(rq + i) * E + n] = dxv;
            if (i < Nk) g.dy[(rk + i) * E + n] = dyv;
          }
        }
    }
  }
  // ---- this CTA's weight-gradient partials -> its slab (accumulator layout: rows mi*16 + g (+8), columns warp*8 + 2t (+1))
  if (g.slab != nullptr) {
    float* s = g.slab + (size_t)blockIdx.x * W_FLOATS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gg = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = mi * 16 + gg + 8 * (e >> 1), k = warp * 8 + 2 * t + (e & 1);
        s[OFF_WIN + n * E + k] = gWq[mi][e];
        s[OFF_WIN + (E + n) * E + k] = gWk[mi][e];
        s[OFF_WIN + (2 * E + n) * E + k] = gWv[mi][e];
        s[OFF_WOUT + n * E + k] = gWo[mi][e];
        s[OFF_WFF + n * E + k] = gWf[mi][e];
      }
    if (threadIdx.x < 3 * E) s[OFF_BIN + threadIdx.x] = gbin;
    if (threadIdx.x < E) { s[OFF_BOUT + threadIdx.x] = gbo; s[OFF_BFF + threadIdx.x] = gbf; }
  }
}

size_t fwd_tc_smem() { return (size_t)(WT_ROWS * LDT + 5 * E + 9 * TILE_F + HEADS * RMAX * SC + RMAX + 6 * RMAX) * sizeof(float); }
size_t bwd_tc_smem() { return (size_t)(WT_ROWS * LDT + 12 * TILE_F + 2 * HEADS * RMAX * SC + RMAX + 6 * RMAX) * sizeof(float); }

size_t fwd_smem() { return (size_t)(W_FLOATS + 7 * E * RS + HEADS * RMAX * SC + RMAX) * sizeof(float); }
size_t bwd_smem() { return (size_t)(W_FLOATS + 12 * E * RS + 2 * HEADS * RMAX * SC + RMAX) * sizeof(float); }

int grid_for(int B) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return B < sms ? B : sms;
}

}  // namespace

bool mab_supported(int E_, int heads, int Nq, int Nk) {
  return E_ == E && heads == HEADS && Nq >= 1 && Nq <= RMAX && Nk >= 1 && Nk <= RMAX;
}
size_t mab_workspace_bytes(int B) {   // transposed weights (forward) / per-CTA slabs (backward)
  return (size_t)W_FLOATS * sizeof(float) * (size_t)(1 + grid_for(B > 0 ? B : 1));
}

int launch_mab_fwd(const MabArgs& a, void* workspace, int precision, cudaStream_t s) {
  if (a.B <= 0) return 0;
  if (precision != 0) {   // TF32 tensor-core GEMMs
    const size_t smem = fwd_tc_smem();
    MPG_CUDA(cudaFuncSetAttribute(mab_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mab_fwd_tc_kernel<<<grid_for(a.B), NT, smem, s>>>(a);
    MPG_LAUNCH_CHECK();
    return 0;
  }
  float* wt = reinterpret_cast<float*>(workspace);
  mab_prepare_kernel<<<cdiv(W_FLOATS, 256), 256, 0, s>>>(a, wt);
  MPG_LAUNCH_CHECK();
  const size_t smem = fwd_smem();
  MPG_CUDA(cudaFuncSetAttribute(mab_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mab_fwd_kernel<<<grid_for(a.B), NT, smem, s>>>(a, wt);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_mab_bwd(const MabArgs& a, MabGrads g, void* workspace, int precision, cudaStream_t s) {
  if (a.B <= 0) return 0;
  const int grid = grid_for(a.B);
  g.slab = g.dw_in != nullptr ? reinterpret_cast<float*>(workspace) + W_FLOATS : nullptr;
  if (precision != 0) {
    const size_t smem = bwd_tc_smem();
    MPG_CHECK(smem <= 227 * 1024, "mab_bwd: shared memory %zu", smem);
    MPG_CUDA(cudaFuncSetAttribute(mab_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mab_bwd_tc_kernel<<<grid, NT, smem, s>>>(a, g);
  } else {
    const size_t smem = bwd_smem();
    MPG_CHECK(smem <= 227 * 1024, "mab_bwd: shared memory %zu", smem);
    MPG_CUDA(cudaFuncSetAttribute(mab_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mab_bwd_kernel<<<grid, NT, smem, s>>>(a, g);
  }
  MPG_LAUNCH_CHECK();
  if (g.slab != nullptr) {
    mab_reduce_kernel<<<cdiv(W_FLOATS, 256), 256, 0, s>>>(g.slab, grid, g);
    MPG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mpg
