// Small HBM / latency-bound kernels added in round 2 (see extra.cuh).
#include "extra.cuh"
#include "edge.cuh"

namespace mpg {
namespace {

// ---- generation post-processing (gen.py:126-141): un-normalise the features, zero the masked particles, clamp the
// third feature at 0, drop the mask channel.  One thread per particle; `out` may be pinned host memory (the writes
// are 12 contiguous bytes per thread, contiguous across the warp).
__global__ void gen_postprocess_kernel(const float* __restrict__ jets, int ldj, float* __restrict__ out, int ldo,
                                       size_t rows, PostCfg c, int use_mask) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* j = jets + r * ldj;
  const bool keep = !use_mask || j[ldj - 1] >= 0.5f;   // gen.py:136 (the mask channel itself is not shifted)
#pragma unroll 4
  for (int i = 0; i < c.nfeat; ++i) {
    float v = j[i];
    if ((c.has_shift >> i) & 1u) v = v - c.shift[i];                               // :128
    if ((c.has_norm >> i) & 1u) v = __fmul_rn(__fdiv_rn(v, c.norm[i]), c.maxv[i]);  // :131-132
    if (!keep) v = 0.f;                                                            // :137
    if (i == 2 && v < 0.f) v = 0.f;                                                // :139
    out[r * ldo + i] = v;
  }
}

// ---- node-network conditioning columns: out[r] = (x[r] | cond[r % B]) -------------------------------------------------
__global__ void cond_columns_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ cond, int C,
                                    float* __restrict__ out, size_t rows, int F, int B) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ldo = F + C;
  if (idx >= rows * ldo) return;
  const size_t r = idx / ldo;
  const int c = (int)(idx % ldo);
  out[idx] = c < F ? x[r * ldx + c] : cond[(r % B) * C + (c - F)];
}

// ---- d(mask = x[..., -1] + 0.5)/dx: zero everywhere but the last column ------------------------------------------
__global__ void split_mask_bwd_kernel(const float* __restrict__ dmask, float* __restrict__ dx, int ldx, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t r = i / ldx;
  const int c = (int)(i % ldx);
  dx[i] = c == ldx - 1 ? dmask[r] : 0.f;
}

// ---- dmask[b,i] = scale * <h[b,i,:], dout[b,:]>  (gradient of the masked sum pool w.r.t. the mask) -----------------
__global__ void pool_dmask_kernel(const float* __restrict__ h, const float* __restrict__ dout, float* __restrict__ dmask,
                                  int N, int C, float scale, size_t rows) {
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per particle
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const size_t b = r / N;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(h[r * C + c], dout[b * C + c], s);
  s = warp_sum(s);
  if (lane == 0) dmask[r] = s * scale;
}

// ---- LayerNorm over the last dimension (gapt/model.py:116-118,130-136; nn.LayerNorm, eps 1e-5, biased variance) ----
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                     float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd,
                                     size_t rows, int C, float eps) {
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + r * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mu = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mu; v = fmaf(d, d, v); }
  const float rs = rsqrtf(warp_sum(v) / (float)C + eps);
  for (int c = lane; c < C; c += 32) y[r * C + c] = fmaf((xr[c] - mu) * rs, w[c], b[c]);
  if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w;  dw += sum_r dy * xhat, db += sum_r dy
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                                     float* __restrict__ dw, float* __restrict__ db, size_t rows, int C) {
  extern __shared__ float sm[];   // [2][C] per-CTA partial sums of dw, db
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (size_t r = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (size_t)gridDim.x * wpb) {
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[r * C + c] - mu) * rs, g = dy[r * C + c] * w[c];
      s1 += g;
      s2 = fmaf(g, xh, s2);
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[r * C + c] - mu) * rs, d = dy[r * C + c];
      dx[r * C + c] = rs * (d * w[c] - s1 - xh * s2);
      if (dw != nullptr) {
        atomicAdd(&sm[c], d * xh);
        atomicAdd(&sm[C + c], d);
      }
    }
  }
  __syncthreads();
  if (dw != nullptr)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      atomicAdd(dw + c, sm[c]);
      atomicAdd(db + c, sm[C + c]);
    }
}

// ---- k nearest neighbours (mpgan/model.py:336-363): distances from every receiver i to every sender j of its jet,
// taken to the sender scaled by 1e4 where masked ((1 - 1e4) * mask + 1e4: so padded particles rank last), over the
// first nd features with 1e-12 added per component; the k smallest (after skipping `skip` = 0/1 for self loops),
// ascending, ties by index (a stable sort's order).  One warp per receiver; N <= 1024.
__global__ void knn_select_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mask, int N, int nd,
                                  int k, int skip, int* __restrict__ idx, size_t rows) {
  extern __shared__ float sm[];   // [warps][N] distances
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + wl;
  if (r >= rows) return;
  float* d = sm + (size_t)wl * N;
  const size_t b = r / N;
  const float* xi = x + r * ldx;
  for (int j = lane; j < N; j += 32) {
    const float* xj = x + (b * N + j) * ldx;
    const float sc = mask ? __fmaf_rn(1.f - 1e4f, mask[b * N + j], 1e4f) : 1.f;
    float d2 = 0.f;
    for (int e = 0; e < nd; ++e) {
      const float de = __fadd_rn(__fsub_rn(mask ? __fmul_rn(sc, xj[e]) : xj[e], xi[e]), 1e-12f);
      d2 = __fmaf_rn(de, de, d2);
    }
    d[j] = sqrtf(d2);
  }
  __syncwarp();
  // rank-based selection: rank(j) = #{l : d_l < d_j or (d_l == d_j and l < j)}; the element of rank skip + m is
  // neighbour m.  O(N^2 / 32) per warp: N <= 150 on this path
  for (int j = lane; j < N; j += 32) {
    const float dj = d[j];
    int rank = 0;
    for (int l = 0; l < N; ++l) {
      const float dl = d[l];
      rank += (dl < dj) || (dl == dj && l < j);
    }
    const int m = rank - skip;
    if (m >= 0 && m < k) idx[r * k + m] = j;
  }
}

}  // namespace

int launch_gen_postprocess(const float* jets, int ldj, float* out, int ldo, size_t rows, const PostCfg& c, int use_mask,
                           cudaStream_t s) {
  if (rows == 0) return 0;
  gen_postprocess_kernel<<<cdiv((long long)rows, 256), 256, 0, s>>>(jets, ldj, out, ldo, rows, c, use_mask);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_cond_columns(const float* x, int ldx, const float* cond, int C, float* out, size_t rows, int F, int B,
                        cudaStream_t s) {
  if (rows == 0) return 0;
  cond_columns_kernel<<<cdiv((long long)(rows * (F + C)), 256), 256, 0, s>>>(x, ldx, cond, C, out, rows, F, B);
  MPG_LAUNCH_CHECK();
  return 0;
}
// ---- receiver compaction map (EdgeArgs::cmap, edge.cuh) -------------------------------------------------------------
// One block, everything parallel.  Jet j gets max(n_j, CMAP_MINW) consecutive positions of the compacted row space (its
// n_j unmasked particles first, in order; the rest stays empty): with every jet at least 15 positions wide a 128-row
// tile touches at most floor(126 / 15) + 2 = 10 jets -- the bound of the edge kernels' Q ring -- and the positions are
// an exclusive prefix sum (a sequential "open a new tile at the 11th jet" rule packs a little tighter for batches of
// tiny jets but costs ~50 us on one thread).  Tile t's jet span comes from two binary searches over the positions.
constexpr int CMAP_MINW = 15;
constexpr size_t CMAP_SMEM_MAX = 160 * 1024;   // positions, counts and the mask bits live in shared memory up to here
__host__ __device__ inline size_t cmap_smem(long long B, long long N) {
  return (size_t)(2 * B + 8) * sizeof(int) + (size_t)((B * N + 3) / 4 + 15) / 16 * 16;
}
__global__ void __launch_bounds__(256) compact_map_kernel(const float* __restrict__ mask, int B, int N, int* __restrict__ cmap,
                                                          int tmax, int* __restrict__ scratch_g /* [2B + 8] */, int in_smem) {
  extern __shared__ int cm_sm[];
  int* start = in_smem ? cm_sm : scratch_g;   // first position of every jet
  int* cnt = start + B;                       // unmasked particles of every jet
  uint8_t* nib = reinterpret_cast<uint8_t*>(cm_sm + 2 * B + 8);   // in_smem: bit (i & 3) of nib[i >> 2] = mask[i] != 0
  __shared__ int wsum[8], s_total;
  int* tile_j0 = cmap + 2;
  int* tile_nj = cmap + 2 + tmax;
  int* rowmap = cmap + 2 + 2 * tmax;
  const int BN = B * N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < tmax * 128; i += blockDim.x) rowmap[i] = -1;
  if (in_smem) {   // one coalesced pass over the mask, eight loads in flight per thread
    const int n4 = (BN + 3) >> 2;
    for (int i0 = threadIdx.x; i0 < n4; i0 += blockDim.x * 8) {
      uint8_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = 0;
        if (i < n4) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (4 * i + e < BN) v[u] |= (uint8_t)((mask[4 * (size_t)i + e] != 0.f) << e);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n4) nib[i] = v[u];
      }
    }
    __syncthreads();
  }
  auto live = [&](int e) { return in_smem ? ((nib[e >> 2] >> (e & 3)) & 1) != 0 : mask[e] != 0.f; };
  // counts and widths: thread t owns jets [t*K, t*K + K)
  const int K = (B + (int)blockDim.x - 1) / (int)blockDim.x;
  const int ja = min(B, (int)threadIdx.x * K), jb = min(B, ja + K);
  int mine = 0;
  for (int j = ja; j < jb; ++j) {
    int n = 0;
    for (int i = 0; i < N; ++i) n += live(j * N + i);
    cnt[j] = n;
    mine += max(n, CMAP_MINW);
  }
  // block-wide exclusive scan of the per-thread widths
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { const int v = wsum[w]; wsum[w] = run; run += v; }
    s_total = run;
    cmap[0] = (run + 127) >> 7;
    cmap[1] = run;
  }
  __syncthreads();
  int pos = wsum[warp] + inc - mine;
  for (int j = ja; j < jb; ++j) {
    start[j] = pos;
    int p = pos;
    for (int i = 0; i < N; ++i)
      if (live(j * N + i)) rowmap[p++] = j * N + i;
    pos += max(cnt[j], CMAP_MINW);
  }
  __syncthreads();
  // jets of every tile: those with an unmasked particle at a position inside [128 t, 128 t + 127]
  const int ntile = (s_total + 127) >> 7;
  for (int t = threadIdx.x; t < tmax; t += blockDim.x) {
    int j0 = 0, nj = 0;
    if (t < ntile) {
      const int a = t * 128, b = a + 127;
      auto last_le = [&](int v) {   // last jet whose first position is <= v (start is non-decreasing, start[0] = 0)
        int lo = 0, hi = B - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (start[mid] <= v) lo = mid; else hi = mid - 1; }
        return lo;
      };
      int lo = last_le(a), hi = last_le(b);
      if (start[lo] + cnt[lo] <= a) ++lo;                 // its particles end before the tile begins
      while (lo <= hi && cnt[lo] == 0) ++lo;
      while (hi >= lo && cnt[hi] == 0) --hi;
      if (lo <= hi) { j0 = lo; nj = hi - lo + 1; }
    }
    tile_j0[t] = j0;
    tile_nj[t] = nj;
  }
}

int launch_compact_map(const float* mask, int B, int N, int* cmap, int* scratch, cudaStream_t s) {
  // (`scratch`, 2B + 8 ints, is only touched by very large batches whose positions do not fit shared memory)
  const bool in_smem = cmap_smem(B, N) <= CMAP_SMEM_MAX;
  const size_t smem = in_smem ? cmap_smem(B, N) : 0;
  MPG_CUDA(cudaFuncSetAttribute(compact_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CMAP_SMEM_MAX));
  compact_map_kernel<<<1, 256, smem, s>>>(mask, B, N, cmap, compact_tiles_max(B, N), scratch, in_smem ? 1 : 0);
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_split_mask_bwd(const float* dmask, float* dx, int ldx, size_t rows, cudaStream_t s) {
  if (rows == 0) return 0;
  const size_t n = rows * (size_t)ldx;
  split_mask_bwd_kernel<<<cdiv((long long)n, 256), 256, 0, s>>>(dmask, dx, ldx, n);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_pool_dmask(const float* h, const float* dout, float* dmask, int B, int N, int C, float scale, cudaStream_t s) {
  const size_t rows = (size_t)B * N;
  if (rows == 0) return 0;
  pool_dmask_kernel<<<cdiv((long long)rows, 8), 256, 0, s>>>(h, dout, dmask, N, C, scale, rows);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, size_t rows,
                         int C, float eps, cudaStream_t s) {
  if (rows == 0) return 0;
  layernorm_fwd_kernel<<<cdiv((long long)rows, 8), 256, 0, s>>>(x, w, b, y, mean, rstd, rows, C, eps);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, float* dx,
                         float* dw, float* db, size_t rows, int C, cudaStream_t s) {
  if (rows == 0) return 0;
  int grid = cdiv((long long)rows, 8 * 16);
  if (grid > 296) grid = 296;
  layernorm_bwd_kernel<<<grid, 256, 2 * C * sizeof(float), s>>>(dy, x, w, mean, rstd, dx, dw, db, rows, C);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_knn_select(const float* x, int ldx, const float* mask, int B, int N, int nd, int k, int skip, int* idx,
                      cudaStream_t s) {
  const size_t rows = (size_t)B * N;
  if (rows == 0) return 0;
  const int wpb = 4;
  knn_select_kernel<<<cdiv((long long)rows, wpb), wpb * 32, (size_t)wpb * N * sizeof(float), s>>>(x, ldx, mask, N, nd, k,
                                                                                               skip, idx, rows);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
