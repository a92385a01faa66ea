// Small HBM-bound kernels of the hot path: rank mask, activation backward, generator tail,
// discriminator pooling, spectral-norm power iteration, fused RMSprop.  All vectorised /
// warp-shuffle, one pass over their operands.
#include "misc.cuh"

namespace mpg {
namespace {

// ---- rank mask (mpgan/model.py:692-699, gapt/model.py:255-262) -------------------------------------
// n = int(fp32(label) * fp32(N)) - 1 (truncation toward zero); mask_i = rank(x0_i) <= n.
// rank = position in an ascending sort = #{k : x_k < x_i or (x_k == x_i and k < i)}.
__global__ void rank_mask_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ labels,
                                 int ldl, int N, float* __restrict__ mask) {
  extern __shared__ float row[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += blockDim.x) row[i] = x[((size_t)b * N + i) * ldx];
  __syncthreads();
  const int n = (int)__fmul_rn(labels[(size_t)b * ldl], (float)N) - 1;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float v = row[i];
    int rank = 0;
    for (int k = 0; k < N; ++k) {
      const float u = row[k];
      rank += (u < v) || (u == v && k < i);
    }
    mask[(size_t)b * N + i] = rank <= n ? 1.f : 0.f;
  }
}

// ---- dz = dy * d(act, dropout)/dz evaluated from the layer OUTPUT y --------------------------------
// One thread per (row, 32-column group): the p == 0.5 keep bits of the group come from ONE Philox draw
// (an element-wise kernel spent ~80 instructions of RNG per element).
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                               int M, int N, int act, float alpha, DropCfg dc, uint32_t stream) {
  resolve_seed(dc);
  const int ng = (N + 31) >> 5;
  const size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= (size_t)M * ng) return;
  const int m = (int)(gidx / ng), c32 = (int)(gidx % ng);
  const int n0 = c32 * 32, nn = min(32, N - n0);
  const bool drop = dc.p > 0.f;
  uint32_t kw = 0xFFFFFFFFu;
  if (drop && dc.half) kw = drop_word32(dc, stream, (uint64_t)m, (uint32_t)c32);
  const size_t base = (size_t)m * N + n0;
  const float keep_scale = drop ? dc.scale : 1.f;
  auto one = [&](float yv, float dyv, int j) {
    float g = act ? lrelu_grad_from_out(yv, alpha) : 1.f;
    bool keep = true;
    if (drop) keep = dc.half ? ((kw >> j) & 1u) : drop_keep(dc, stream, (uint64_t)m, (uint32_t)(n0 + j));
    return keep ? dyv * g * keep_scale : 0.f;
  };
  if (nn == 32 && (N & 3) == 0) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 yv = *reinterpret_cast<const float4*>(y + base + j);
      const float4 dv = *reinterpret_cast<const float4*>(dy + base + j);
      *reinterpret_cast<float4*>(dz + base + j) =
          make_float4(one(yv.x, dv.x, j), one(yv.y, dv.y, j + 1), one(yv.z, dv.z, j + 2), one(yv.w, dv.w, j + 3));
    }
  } else {
    for (int j = 0; j < nn; ++j) dz[base + j] = one(y[base + j], dy[base + j], j);
  }
}

// ---- particle order: real particles first inside every jet (stable) ----------------------------------
// pos[b, i] = new index of particle i: its rank among the real ones (mask != 0) if real, else n_real + its rank
// among the padded ones.  One warp per jet; also writes the permuted mask.  The message-passing layers are
// permutation-equivariant over particles, so this is a layout choice (model.py: MPNet.forward).
__global__ void particle_order_kernel(const float* __restrict__ mask, int B, int N, int* __restrict__ pos,
                                      float* __restrict__ mask_sorted) {
  const int jet = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (jet >= B) return;
  const float* m = mask + (size_t)jet * N;
  int n_real = 0;
  for (int i0 = 0; i0 < N; i0 += 32) {
    const bool real = i0 + lane < N && m[i0 + lane] != 0.f;
    n_real += __popc(__ballot_sync(0xffffffffu, real));
  }
  int seen_real = 0, seen_pad = 0;
  for (int i0 = 0; i0 < N; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i < N;
    const float mv = in ? m[i] : 0.f;
    const bool real = in && mv != 0.f;
    const uint32_t br = __ballot_sync(0xffffffffu, real), bp = __ballot_sync(0xffffffffu, in && !real);
    const uint32_t below = (1u << lane) - 1u;
    if (in) {
      const int p = real ? seen_real + __popc(br & below) : n_real + seen_pad + __popc(bp & below);
      pos[(size_t)jet * N + i] = p;
      mask_sorted[(size_t)jet * N + p] = mv;
    }
    seen_real += __popc(br);
    seen_pad += __popc(bp);
  }
}

// ---- batch order: jets by descending particle-count key, stable (train.py has no counterpart: layout only) ----
// pos[b] = #{b' : key[b'] > key[b]} + #{b' < b : key[b'] == key[b]}: one thread per jet, the keys streamed through
// shared memory in tiles of 2048 (any B; B = 256: one CTA, one tile).
__global__ void batch_order_kernel(const float* __restrict__ key, int ldk, int B, int* __restrict__ pos) {
  constexpr int TILE_K = 2048;
  __shared__ float keys[TILE_K];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const float k = b < B ? key[(size_t)b * ldk] : 0.f;
  int p = 0;
  for (int j0 = 0; j0 < B; j0 += TILE_K) {
    const int nj = min(TILE_K, B - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nj; i += blockDim.x) keys[i] = key[(size_t)(j0 + i) * ldk];
    __syncthreads();
    if (b < B)
      for (int j = 0; j < nj; ++j) {
        const float kj = keys[j];
        p += (kj > k) || (kj == k && j0 + j < b);
      }
  }
  if (b < B) pos[b] = p;
}

// scatter (mode 0): dst[b, pos[b,i], :] = src[b, i, :];  gather (mode 1): dst[b, i, :] = src[b, pos[b,i], :]
__global__ void permute_rows_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                    const int* __restrict__ pos, int N, int F, size_t total, int mode) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t r = idx / F;
  const int f = (int)(idx % F);
  const size_t jet0 = (r / N) * N;
  const size_t pr = jet0 + pos[r];
  if (mode == 0) dst[pr * ldd + f] = src[r * lds + f];
  else dst[r * ldd + f] = src[pr * lds + f];
}

// ---- generator tail: out[..., :Fo] = act(h), out[..., Fo] = mask - 0.5 (model.py:535-536,752) ------
__global__ void gen_tail_fwd_kernel(const float* __restrict__ h, const float* __restrict__ mask,
                                    float* __restrict__ out, int rows, int Fo, int act) {
  const int ldo = Fo + (mask != nullptr);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * ldo) return;
  const int r = (int)(idx / ldo), c = (int)(idx % ldo);
  float v;
  if (c < Fo) {
    v = h[(size_t)r * Fo + c];
    if (act == 1) v = tanhf(v);
    else if (act == 2) v = 1.f / (1.f + expf(-v));
  } else {
    v = mask[r] - 0.5f;
  }
  out[idx] = v;
}

__global__ void gen_tail_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                    float* __restrict__ dh, int rows, int Fo, int ldo, int act) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * Fo) return;
  const int r = (int)(idx / Fo), c = (int)(idx % Fo);
  const float y = out[(size_t)r * ldo + c];
  float g = dout[(size_t)r * ldo + c];
  if (act == 1) g *= 1.f - y * y;
  else if (act == 2) g *= y * (1.f - y);
  dh[idx] = g;
}

// ---- discriminator input split: mask = x[..., -1] + 0.5 (model.py:881) ----------------------------
__global__ void split_mask_kernel(const float* __restrict__ x, int ldx, int rows, float* __restrict__ mask) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) mask[r] = x[(size_t)r * ldx + ldx - 1] + 0.5f;
}

// ---- masked pooling over particles (model.py:810-822) ----------------------------------------------
// pooled[b][c] = sum_i h[b,i,c]*mask[b,i]  (/ (sum_i mask + 1e-12) if mean);  no mask: sum or mean.
__global__ void pool_fwd_kernel(const float* __restrict__ h, const float* __restrict__ mask, float* __restrict__ out,
                                int N, int C, int mean) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, ms = 0.f;
    for (int i = 0; i < N; ++i) {
      const float m = mask ? mask[(size_t)b * N + i] : 1.f;
      s = fmaf(h[((size_t)b * N + i) * C + c], m, s);
      ms += m;
    }
    if (mean) s = mask ? s / (ms + 1e-12f) : s / (float)N;
    out[(size_t)b * C + c] = s;
  }
}

__global__ void pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ mask,
                                float* __restrict__ dh, int N, int C, int mean) {
  const int b = blockIdx.x;
  float ms = 0.f;
  if (mean && mask)
    for (int i = 0; i < N; ++i) ms += mask[(size_t)b * N + i];
  const float inv = mean ? (mask ? 1.f / (ms + 1e-12f) : 1.f / (float)N) : 1.f;
  for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) {
    const int i = idx / C, c = idx % C;
    const float m = mask ? mask[(size_t)b * N + i] : 1.f;
    dh[((size_t)b * N + i) * C + c] = dout[(size_t)b * C + c] * m * inv;
  }
}

// ---- least-squares GAN loss (train.py:357-358,369-370,378,467,472: MSELoss against 1 / 0 targets) --------
// loss = mean_{i < n0} (d_i - t0)^2 + [n0 < n] mean_{i >= n0} (d_i - t1)^2: one CTA, one launch instead of the
// sub / pow / mean / add chain (and its autograd mirror) on [B, 1] tensors.
__global__ void ls_loss_fwd_kernel(const float* __restrict__ d, int n, int n0, float t0, float t1, float* __restrict__ loss) {
  __shared__ float red[2][8];
  float s0 = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float e = d[i] - (i < n0 ? t0 : t1);
    if (i < n0) s0 += e * e; else s1 += e * e;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
    *loss = a / (float)n0 + (n > n0 ? b / (float)(n - n0) : 0.f);
  }
}
__global__ void ls_loss_bwd_kernel(const float* __restrict__ d, const float* __restrict__ gout, int n, int n0, float t0,
                                   float t1, float* __restrict__ dd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = *gout;
  dd[i] = i < n0 ? g * 2.f * (d[i] - t0) / (float)n0 : g * 2.f * (d[i] - t1) / (float)(n - n0);
}

// ---- elementwise unary with backward from output ---------------------------------------------------
__global__ void unary_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int act) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float v = x[idx];
  y[idx] = act == 1 ? tanhf(v) : (act == 2 ? 1.f / (1.f + expf(-v)) : v);
}
__global__ void unary_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                 size_t n, int act) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float v = y[idx];
  dx[idx] = dy[idx] * (act == 1 ? 1.f - v * v : (act == 2 ? v * (1.f - v) : 1.f));
}

// ---- spectral norm (spectral_normalization.py:21-33) -----------------------------------------------
// one CTA per weight: v = l2n(W^T u); u = l2n(W v); sigma = u.(W v); W_out = W / (sigma + 1e-12)
__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void sn_fwd_kernel(const float* __restrict__ Wb, float* __restrict__ u, float* __restrict__ v,
                              float* __restrict__ Wout, float* __restrict__ sigma_out, int H, int Wd) {
  extern __shared__ float sm[];
  float* us = sm;            // [H]
  float* vs = us + H;        // [Wd]
  float* red = vs + Wd;      // [32]
  for (int i = threadIdx.x; i < H; i += blockDim.x) us[i] = u[i];
  __syncthreads();
  // v = W^T u
  float part = 0.f;
  for (int j = threadIdx.x; j < Wd; j += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < H; ++i) s = fmaf(Wb[(size_t)i * Wd + j], us[i], s);
    vs[j] = s;
    part = fmaf(s, s, part);
  }
  float nrm = sqrtf(block_sum(part, red));
  for (int j = threadIdx.x; j < Wd; j += blockDim.x) vs[j] = vs[j] / (nrm + 1e-12f);
  __syncthreads();
  // u = W v  (one warp per row)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = w; i < H; i += nw) {
    float s = 0.f;
    for (int j = lane; j < Wd; j += 32) s = fmaf(Wb[(size_t)i * Wd + j], vs[j], s);
    s = warp_sum(s);
    if (lane == 0) us[i] = s;   // holds W v
  }
  __syncthreads();
  part = 0.f;
  for (int i = threadIdx.x; i < H; i += blockDim.x) part = fmaf(us[i], us[i], part);
  nrm = sqrtf(block_sum(part, red));
  // u_new = Wv / (|Wv| + eps);  sigma = u_new . (W v)
  part = 0.f;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float wv = us[i];
    const float un = wv / (nrm + 1e-12f);
    part = fmaf(un, wv, part);
    u[i] = un;
  }
  const float sigma = block_sum(part, red);
  for (int j = threadIdx.x; j < Wd; j += blockDim.x) v[j] = vs[j];
  if (threadIdx.x == 0) *sigma_out = sigma;
  const float inv = 1.f / (sigma + 1e-12f);
  for (int idx = threadIdx.x; idx < H * Wd; idx += blockDim.x) Wout[idx] = Wb[idx] * inv;
}

// dWb += dW/(s+eps) - (sum(dW*Wb)/(s+eps)^2) * u v^T      (u, v constants; sigma = u^T Wb v)
__global__ void sn_bwd_kernel(const float* __restrict__ dW, const float* __restrict__ Wb,
                              const float* __restrict__ u, const float* __restrict__ v,
                              const float* __restrict__ sigma, float* __restrict__ dWb, int H, int Wd) {
  __shared__ float red[32];
  float part = 0.f;
  for (int idx = threadIdx.x; idx < H * Wd; idx += blockDim.x) part = fmaf(dW[idx], Wb[idx], part);
  const float dot = block_sum(part, red);
  const float inv = 1.f / (*sigma + 1e-12f);
  const float coef = dot * inv * inv;
  for (int idx = threadIdx.x; idx < H * Wd; idx += blockDim.x) {
    const int i = idx / Wd, j = idx % Wd;
    dWb[idx] += dW[idx] * inv - coef * u[i] * v[j];
  }
}

// ---- fused RMSprop over a flat parameter buffer (torch.optim.RMSprop defaults) ---------------------
__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq, size_t n,
                               float lr, float alpha, float eps, float gscale) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float gr = g[idx] * gscale;
  const float s = alpha * sq[idx] + (1.f - alpha) * gr * gr;
  sq[idx] = s;
  p[idx] -= lr * gr / (sqrtf(s) + eps);
}

}  // namespace

int launch_rank_mask(const float* x, int ldx, const float* labels, int ldl, int B, int N, float* mask,
                     cudaStream_t s) {
  if (B <= 0) return 0;
  rank_mask_kernel<<<B, 128, N * sizeof(float), s>>>(x, ldx, labels, ldl, N, mask);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_act_bwd(const float* dy, const float* y, float* dz, int M, int N, int act, float alpha, DropCfg dc,
                   uint32_t stream, cudaStream_t s) {
  const size_t n = (size_t)M * N;
  if (n == 0) return 0;
  act_bwd_kernel<<<cdiv((size_t)M * ((N + 31) / 32), 128), 128, 0, s>>>(dy, y, dz, M, N, act, alpha, dc, stream);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_gen_tail_fwd(const float* h, const float* mask, float* out, int rows, int Fo, int act, cudaStream_t s) {
  const size_t n = (size_t)rows * (Fo + (mask != nullptr));
  if (n == 0) return 0;
  gen_tail_fwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(h, mask, out, rows, Fo, act);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_gen_tail_bwd(const float* dout, const float* out, float* dh, int rows, int Fo, int ldo, int act,
                        cudaStream_t s) {
  const size_t n = (size_t)rows * Fo;
  if (n == 0) return 0;
  gen_tail_bwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(dout, out, dh, rows, Fo, ldo, act);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_particle_order(const float* mask, int B, int N, int* pos, float* mask_sorted, cudaStream_t s) {
  if (B <= 0 || N <= 0) return 0;
  particle_order_kernel<<<cdiv(B, 8), 256, 0, s>>>(mask, B, N, pos, mask_sorted);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_batch_order(const float* key, int ldk, int B, int* pos, cudaStream_t s) {
  if (B <= 0) return 0;
  const int nt = B < 256 ? ((B + 31) / 32) * 32 : 256;
  batch_order_kernel<<<cdiv(B, nt), nt, 0, s>>>(key, ldk, B, pos);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_permute_rows(const float* src, int lds, float* dst, int ldd, const int* pos, int B, int N, int F, int mode,
                        cudaStream_t s) {
  const size_t total = (size_t)B * N * F;
  if (total == 0) return 0;
  permute_rows_kernel<<<cdiv(total, 256), 256, 0, s>>>(src, lds, dst, ldd, pos, N, F, total, mode);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_ls_loss(const float* d, const float* gout, int n, int n0, float t0, float t1, float* out, bool bwd,
                   cudaStream_t s) {
  if (n <= 0) return 0;
  if (bwd) ls_loss_bwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, gout, n, n0, t0, t1, out);
  else ls_loss_fwd_kernel<<<1, 256, 0, s>>>(d, n, n0, t0, t1, out);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_split_mask(const float* x, int ldx, int rows, float* mask, cudaStream_t s) {
  if (rows <= 0) return 0;
  split_mask_kernel<<<cdiv(rows, 256), 256, 0, s>>>(x, ldx, rows, mask);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_pool_fwd(const float* h, const float* mask, float* out, int B, int N, int C, int mean, cudaStream_t s) {
  if (B <= 0) return 0;
  pool_fwd_kernel<<<B, 64, 0, s>>>(h, mask, out, N, C, mean);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_pool_bwd(const float* dout, const float* mask, float* dh, int B, int N, int C, int mean, cudaStream_t s) {
  if (B <= 0) return 0;
  pool_bwd_kernel<<<B, 256, 0, s>>>(dout, mask, dh, N, C, mean);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_unary(const float* x, const float* dy, float* out, size_t n, int act, bool bwd, cudaStream_t s) {
  if (n == 0) return 0;
  if (bwd) unary_bwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(dy, x, out, n, act);
  else unary_fwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(x, out, n, act);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_sn_fwd(const float* Wb, float* u, float* v, float* Wout, float* sigma, int H, int Wd, cudaStream_t s) {
  const size_t smem = (size_t)(H + Wd + 32) * sizeof(float);
  sn_fwd_kernel<<<1, 256, smem, s>>>(Wb, u, v, Wout, sigma, H, Wd);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_sn_bwd(const float* dW, const float* Wb, const float* u, const float* v, const float* sigma,
                  float* dWb, int H, int Wd, cudaStream_t s) {
  sn_bwd_kernel<<<1, 256, 0, s>>>(dW, Wb, u, v, sigma, dWb, H, Wd);
  MPG_LAUNCH_CHECK();
  return 0;
}
int launch_rmsprop(float* p, const float* g, float* sq, size_t n, float lr, float alpha, float eps, float gscale,
                   cudaStream_t s) {
  if (n == 0) return 0;
  rmsprop_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, g, sq, n, lr, alpha, eps, gscale);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
