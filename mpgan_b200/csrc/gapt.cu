// GAPT set attention core (gapt/model.py:124-139 via nn.MultiheadAttention): per (jet, head)
// masked softmax(q k^T / sqrt(d)) v with warp-per-head, lane-per-query mapping; K/V tiles live in
// shared memory, probabilities are kept for the backward.  The projections around it are the
// node-level GEMMs (gemm.cu); residual + dropout is one elementwise kernel.
#include "gapt.cuh"

namespace mpg {
namespace {

constexpr int MAXD = 32;   // head dim limit (registers per lane)

// smem: ks[Nk*E], vs[Nk*E], ign[Nk], sc[warps][32][Nk+1]
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnArgs a, float* __restrict__ o, float* __restrict__ P) {
  extern __shared__ __align__(16) float sm[];
  const int E = a.E, Nk = a.Nk, Nq = a.Nq, d = E / a.heads;
  float* ks = sm;
  float* vs = ks + (size_t)Nk * E;
  float* ign = vs + (size_t)Nk * E;
  float* sc = ign + Nk;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int idx = threadIdx.x; idx < Nk * E; idx += blockDim.x) {
    const int j = idx / E, c = idx % E;
    ks[idx] = a.k[((size_t)b * Nk + j) * a.ldk + c];
    vs[idx] = a.v[((size_t)b * Nk + j) * a.ldv + c];
  }
  // keys whose mask is not exactly 1.0 are ignored: (1 - mask).bool()  (gapt/model.py:194-202)
  for (int j = threadIdx.x; j < Nk; j += blockDim.x)
    ign[j] = (a.key_mask != nullptr && (1.f - a.key_mask[(size_t)b * Nk + j]) != 0.f) ? 1.f : 0.f;
  __syncthreads();
  const float scale = rsqrtf((float)d);
  float* myrow = sc + ((size_t)warp * 32 + lane) * (Nk + 1);
  for (int h = warp; h < a.heads; h += nw) {
    for (int i0 = 0; i0 < Nq; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < Nq;
      float q[MAXD];
#pragma unroll
      for (int c = 0; c < MAXD; ++c)
        q[c] = (valid && c < d) ? a.q[((size_t)b * Nq + i) * a.ldq + h * d + c] * scale : 0.f;
      float mx = -INFINITY;
      for (int j = 0; j < Nk; ++j) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < MAXD; ++c)
          if (c < d) s = fmaf(q[c], ks[j * E + h * d + c], s);
        if (ign[j] != 0.f) s = -INFINITY;
        myrow[j] = s;
        mx = fmaxf(mx, s);
      }
      float sum = 0.f;
      for (int j = 0; j < Nk; ++j) {
        const float p = (mx == -INFINITY) ? 0.f : expf(myrow[j] - mx);
        myrow[j] = p;
        sum += p;
      }
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      float acc[MAXD];
#pragma unroll
      for (int c = 0; c < MAXD; ++c) acc[c] = 0.f;
      for (int j = 0; j < Nk; ++j) {
        const float p = myrow[j] * inv;
        myrow[j] = p;
#pragma unroll
        for (int c = 0; c < MAXD; ++c)
          if (c < d) acc[c] = fmaf(p, vs[j * E + h * d + c], acc[c]);
      }
      // probabilities of this (head, query chunk): the warp's score rows -> P, contiguous in memory
      __syncwarp();
      {
        const int ni = min(32, Nq - i0);
        float* Pd = P + (((size_t)b * a.heads + h) * Nq + i0) * Nk;
        const float* rows = sc + (size_t)warp * 32 * (Nk + 1);
        for (int idx = lane; idx < ni * Nk; idx += 32) Pd[idx] = rows[(idx / Nk) * (Nk + 1) + idx % Nk];
      }
      __syncwarp();
      if (valid)
#pragma unroll
        for (int c = 0; c < MAXD; ++c)
          if (c < d) o[((size_t)b * Nq + i) * E + h * d + c] = acc[c];
    }
  }
}

// smem: ks, vs [Nk*E]; qs, dos [Nq*E]; dSs, Ps [heads? no: per warp 32 x (Nk+1)] x2
__global__ void __launch_bounds__(128) attn_bwd_kernel(AttnArgs a, const float* __restrict__ P,
                                                       const float* __restrict__ dO, float* __restrict__ dq,
                                                       float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  const int E = a.E, Nk = a.Nk, Nq = a.Nq, d = E / a.heads;
  float* ks = sm;
  float* vs = ks + (size_t)Nk * E;
  const int EP = E + 1;                                   // padded row: lane = query reads are conflict-free
  float* qs = vs + (size_t)Nk * E;
  float* dos = qs + (size_t)Nq * EP;
  float* dSs = dos + (size_t)Nq * EP;                     // [warps][32][Nk+1]
  float* Ps = dSs + (size_t)(blockDim.x >> 5) * 32 * (Nk + 1);
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int idx = threadIdx.x; idx < Nk * E; idx += blockDim.x) {
    const int j = idx / E, c = idx % E;
    ks[idx] = a.k[((size_t)b * Nk + j) * a.ldk + c];
    vs[idx] = a.v[((size_t)b * Nk + j) * a.ldv + c];
  }
  for (int idx = threadIdx.x; idx < Nq * E; idx += blockDim.x) {
    const int i = idx / E, c = idx % E;
    qs[i * EP + c] = a.q[((size_t)b * Nq + i) * a.ldq + c];
    dos[i * EP + c] = dO[((size_t)b * Nq + i) * E + c];
  }
  __syncthreads();
  const float scale = rsqrtf((float)d);
  float* mydS = dSs + (size_t)warp * 32 * (Nk + 1);
  float* myP = Ps + (size_t)warp * 32 * (Nk + 1);
  for (int h = warp; h < a.heads; h += nw) {
    // dk/dv accumulate over query chunks: lane = key within a 32-key block, kept in registers per block
    for (int j0 = 0; j0 < Nk; j0 += 32) {
      float dkacc[MAXD], dvacc[MAXD];
#pragma unroll
      for (int c = 0; c < MAXD; ++c) dkacc[c] = dvacc[c] = 0.f;
      for (int i0 = 0; i0 < Nq; i0 += 32) {
        // pass A (lane = query): dS rows for this query chunk (all keys), dq on the first key block only
        const int i = i0 + lane;
        const bool valid = i < Nq;
        __syncwarp();
        {   // probabilities of this (head, query chunk): contiguous in memory -> the warp's P rows
          const int ni = min(32, Nq - i0);
          const float* Ps_g = P + (((size_t)b * a.heads + h) * Nq + i0) * Nk;
          for (int idx = lane; idx < ni * Nk; idx += 32) myP[(idx / Nk) * (Nk + 1) + idx % Nk] = Ps_g[idx];
        }
        float dor[MAXD];   // this query's dO row for head h (was re-read from shared memory for every key)
#pragma unroll
        for (int c = 0; c < MAXD; ++c) dor[c] = (valid && c < d) ? dos[i * EP + h * d + c] : 0.f;
        __syncwarp();
        float D = 0.f;
        for (int j = 0; j < Nk; ++j) {
          float p = 0.f, dp = 0.f;
          if (valid) {
            p = myP[lane * (Nk + 1) + j];
#pragma unroll
            for (int c = 0; c < MAXD; ++c)
              if (c < d) dp = fmaf(dor[c], vs[j * E + h * d + c], dp);
          } else {
            myP[lane * (Nk + 1) + j] = 0.f;
          }
          mydS[lane * (Nk + 1) + j] = dp;
          D = fmaf(p, dp, D);
        }
        float dqacc[MAXD];
#pragma unroll
        for (int c = 0; c < MAXD; ++c) dqacc[c] = 0.f;
        for (int j = 0; j < Nk; ++j) {
          const float ds = myP[lane * (Nk + 1) + j] * (mydS[lane * (Nk + 1) + j] - D);
          mydS[lane * (Nk + 1) + j] = ds;
          if (j0 == 0) {
#pragma unroll
            for (int c = 0; c < MAXD; ++c)
              if (c < d) dqacc[c] = fmaf(ds, ks[j * E + h * d + c], dqacc[c]);
          }
        }
        if (j0 == 0 && valid)
#pragma unroll
          for (int c = 0; c < MAXD; ++c)
            if (c < d) dq[((size_t)b * Nq + i) * E + h * d + c] = dqacc[c] * scale;
        __syncwarp();
        // pass B (lane = key): reduce over the queries of this chunk
        const int j = j0 + lane;
        if (j < Nk) {
          const int ni = min(32, Nq - i0);
          for (int ii = 0; ii < ni; ++ii) {
            const float ds = mydS[ii * (Nk + 1) + j];
            const float p = myP[ii * (Nk + 1) + j];
#pragma unroll
            for (int c = 0; c < MAXD; ++c)
              if (c < d) {
                dkacc[c] = fmaf(ds, qs[(i0 + ii) * EP + h * d + c], dkacc[c]);
                dvacc[c] = fmaf(p, dos[(i0 + ii) * EP + h * d + c], dvacc[c]);
              }
          }
        }
      }
      const int j = j0 + lane;
      if (j < Nk)
#pragma unroll
        for (int c = 0; c < MAXD; ++c)
          if (c < d) {
            dk[((size_t)b * Nk + j) * E + h * d + c] = dkacc[c] * scale;
            dv[((size_t)b * Nk + j) * E + h * d + c] = dvacc[c];
          }
    }
  }
}

// thread -> 8 consecutive columns of one row: one Philox draw per thread on the p == 0.5 path (8 of its 32 bits)
// instead of one per element; cols % 8 == 0 takes the float4 path
template <bool BWD>
__global__ void resdrop_kernel(const float* __restrict__ x, const float* __restrict__ r, float* __restrict__ out,
                               size_t rows, int cols, DropCfg dc, uint32_t stream) {
  resolve_seed(dc);
  const int ng = (cols + 7) >> 3;
  const size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= rows * ng) return;
  const size_t row = gidx / ng;
  const int c0 = (int)(gidx % ng) * 8, nn = min(8, cols - c0);
  const bool drop = dc.p > 0.f;
  uint32_t kw = 0xFFFFFFFFu;
  if (drop && dc.half) kw = drop_word32(dc, stream, row, (uint32_t)(c0 >> 5)) >> (c0 & 31);
  auto one = [&](float v, int j) {
    if (!drop) return v;
    const bool keep = dc.half ? ((kw >> j) & 1u) : drop_keep(dc, stream, row, (uint32_t)(c0 + j));
    return keep ? v * dc.scale : 0.f;
  };
  const size_t base = row * cols + c0;
  if (nn == 8 && (cols & 3) == 0) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      float4 v = *reinterpret_cast<const float4*>(x + base + j);
      if (!BWD && r != nullptr) {
        const float4 rv = *reinterpret_cast<const float4*>(r + base + j);
        v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
      }
      *reinterpret_cast<float4*>(out + base + j) = make_float4(one(v.x, j), one(v.y, j + 1), one(v.z, j + 2), one(v.w, j + 3));
    }
  } else {
    for (int j = 0; j < nn; ++j) out[base + j] = one(x[base + j] + ((!BWD && r) ? r[base + j] : 0.f), j);
  }
}

}  // namespace

static int attn_check(const AttnArgs& a) {
  MPG_CHECK(a.heads > 0 && a.E % a.heads == 0 && a.E / a.heads <= MAXD, "attention: head dim must be <= %d", MAXD);
  MPG_CHECK(a.B >= 0 && a.Nq > 0 && a.Nk > 0, "attention: bad sizes");
  return 0;
}

int launch_attn_fwd(const AttnArgs& a, float* o, float* P, cudaStream_t s) {
  if (attn_check(a)) return 1;
  if (a.B == 0) return 0;
  const size_t smem = ((size_t)2 * a.Nk * a.E + a.Nk + (size_t)4 * 32 * (a.Nk + 1)) * sizeof(float);
  MPG_CHECK(smem <= 227 * 1024, "attention: too many keys for shared memory (%zu B)", smem);
  MPG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_kernel<<<a.B, 128, smem, s>>>(a, o, P);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_attn_bwd(const AttnArgs& a, const float* P, const float* dO, float* dq, float* dk, float* dv,
                    cudaStream_t s) {
  if (attn_check(a)) return 1;
  if (a.B == 0) return 0;
  const size_t smem =
      ((size_t)2 * a.Nk * a.E + (size_t)2 * a.Nq * (a.E + 1) + (size_t)2 * 4 * 32 * (a.Nk + 1)) * sizeof(float);
  MPG_CHECK(smem <= 227 * 1024, "attention backward: set too large for shared memory (%zu B)", smem);
  MPG_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_kernel<<<a.B, 128, smem, s>>>(a, P, dO, dq, dk, dv);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_resdrop(const float* x, const float* r, float* out, size_t rows, int cols, DropCfg dc, uint32_t stream,
                   bool bwd, cudaStream_t s) {
  const size_t n = rows * cols;
  if (n == 0) return 0;
  const size_t ngroups = rows * (size_t)((cols + 7) / 8);
  if (bwd) resdrop_kernel<true><<<cdiv(ngroups, 256), 256, 0, s>>>(x, nullptr, out, rows, cols, dc, stream);
  else resdrop_kernel<false><<<cdiv(ngroups, 256), 256, 0, s>>>(x, r, out, rows, cols, dc, stream);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
