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

int launch_mab_fwd(const MabArgs& a, void* workspace, cudaStream_t s) {
  if (a.B <= 0) return 0;
  float* wt = reinterpret_cast<float*>(workspace);
  mab_prepare_kernel<<<cdiv(W_FLOATS, 256), 256, 0, s>>>(a, wt);
  MPG_LAUNCH_CHECK();
  const size_t smem = fwd_smem();
  MPG_CUDA(cudaFuncSetAttribute(mab_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mab_fwd_kernel<<<grid_for(a.B), NT, smem, s>>>(a, wt);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_mab_bwd(const MabArgs& a, MabGrads g, void* workspace, cudaStream_t s) {
  if (a.B <= 0) return 0;
  const int grid = grid_for(a.B);
  g.slab = g.dw_in != nullptr ? reinterpret_cast<float*>(workspace) + W_FLOATS : nullptr;
  const size_t smem = bwd_smem();
  MPG_CHECK(smem <= 227 * 1024, "mab_bwd: shared memory %zu", smem);
  MPG_CUDA(cudaFuncSetAttribute(mab_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mab_bwd_kernel<<<grid, NT, smem, s>>>(a, g);
  MPG_LAUNCH_CHECK();
  if (g.slab != nullptr) {
    mab_reduce_kernel<<<cdiv(W_FLOATS, 256), 256, 0, s>>>(g.slab, grid, g);
    MPG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace mpg
