// TF32 mma.sync GEMM for node-level (O(B*N) rows) matrices.  See gemm.cuh.
#include "gemm.cuh"

namespace mpg {

namespace {

constexpr int BM = 64, BN = 64, BK = 32, NT = 128;
constexpr int KS = BK + 8;   // stride of a [rows][k] tile  (conflict-free 64-bit fragment reads)
constexpr int MS = BM + 8;   // stride of a [k][rows] tile

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Loads one BMxBK (or BNxBK) operand tile into registers: 16 floats per thread.
// KMAJ: source element (r, k) at src[r*ld + k]; else at src[k*ld + r].
template <bool KMAJ>
__device__ __forceinline__ void load_tile(float (&reg)[16], const float* __restrict__ src, int ld, int r0,
                                          int k0, int R, int K, int tid, bool vec_ok) {
  if (KMAJ) {
    // thread -> (row = tid/4 + 32*i, k = (tid%4)*8 .. +7): 8 consecutive k per row, so that the store can
    // interleave k and k+4 (the two halves of an mma fragment pair)
    const int kk = k0 + (tid & 3) * 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + (tid >> 2) + 32 * i;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (r < R) {
        const float* p = src + (size_t)r * ld + kk;
        if (vec_ok && kk + 7 < K) {
          const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (kk + e < K) v[e] = p[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) reg[8 * i + e] = v[e];
    }
  } else {
    // thread -> (k = tid/16 + 8*i, row = (tid%16)*4 .. +3)
    const int rr = r0 + (tid & 15) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + (tid >> 4) + 8 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const float* p = src + (size_t)k * ld + rr;
        if (vec_ok && rr + 3 < R) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (rr + 0 < R) v.x = p[0];
          if (rr + 1 < R) v.y = p[1];
          if (rr + 2 < R) v.z = p[2];
          if (rr + 3 < R) v.w = p[3];
        }
      }
      reg[4 * i + 0] = v.x; reg[4 * i + 1] = v.y; reg[4 * i + 2] = v.z; reg[4 * i + 3] = v.w;
    }
  }
}

// CVT: round to TF32 here, once per element (fast path); the 3xTF32 path keeps fp32 and splits at use.
// K-major tiles store k and k+4 of every 8-group next to each other: position 2*(k%4) + (k/4)%2.
template <bool KMAJ, bool CVT>
__device__ __forceinline__ void store_tile(float* __restrict__ s, const float (&reg)[16], int tid) {
  auto cv = [](float x) { return CVT ? __uint_as_float(to_tf32(x)) : x; };
  if (KMAJ) {  // s[row][k'], stride KS
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float* p = s + ((tid >> 2) + 32 * i) * KS + (tid & 3) * 8;
      const float* v = reg + 8 * i;
      *reinterpret_cast<float4*>(p) = make_float4(cv(v[0]), cv(v[4]), cv(v[1]), cv(v[5]));
      *reinterpret_cast<float4*>(p + 4) = make_float4(cv(v[2]), cv(v[6]), cv(v[3]), cv(v[7]));
    }
  } else {     // s[k][row], stride MS
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* p = s + ((tid >> 4) + 8 * i) * MS + (tid & 15) * 4;
      *reinterpret_cast<float4*>(p) = make_float4(cv(reg[4 * i]), cv(reg[4 * i + 1]), cv(reg[4 * i + 2]), cv(reg[4 * i + 3]));
    }
  }
}

// (element (r, k), element (r, k + 4)) of a tile, k = kk + t with kk a multiple of 8
template <bool KMAJ>
__device__ __forceinline__ float2 tile_pair(const float* __restrict__ s, int r, int kk, int t) {
  if (KMAJ) return *reinterpret_cast<const float2*>(s + r * KS + kk + 2 * t);
  return make_float2(s[(kk + t) * MS + r], s[(kk + t + 4) * MS + r]);
}

constexpr int TILE_FLOATS = (BM * KS > BK * MS) ? BM * KS : BK * MS;

// The epilogue is unrolled over 32 outputs per thread; inlining the Philox-based keep decision twice per
// output made the kernel 15k instructions (245 KB) and the short node-level GEMMs instruction-cache bound.
__device__ __noinline__ bool drop_keep_call(const DropCfg& d, uint32_t stream, uint64_t row, uint32_t col) {
  return drop_keep(d, stream, row, col);
}

template <bool A_K, bool B_K, bool PRECISE>
__global__ void __launch_bounds__(NT) gemm_kernel(const float* __restrict__ A, int lda,
                                                  const float* __restrict__ B, int ldb,
                                                  float* __restrict__ C, int ldc, int M, int N, int K,
                                                  int k_per_split, GemmEpi epi, bool vecA, bool vecB, bool vecC) {
  if (epi.drop) resolve_seed(epi.dc);
  if (epi.g_drop) resolve_seed(epi.gdc);
  __shared__ __align__(16) float As[2][TILE_FLOATS];
  __shared__ __align__(16) float Bs[2][TILE_FLOATS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

  float ra[16], rb[16];
  const int nchunks = (kend - kbeg + BK - 1) / BK;
  if (nchunks > 0) {
    load_tile<A_K>(ra, A, lda, m0, kbeg, M, kend, tid, vecA);
    load_tile<B_K>(rb, B, ldb, n0, kbeg, N, kend, tid, vecB);
    store_tile<A_K, !PRECISE>(As[0], ra, tid);
    store_tile<B_K, !PRECISE>(Bs[0], rb, tid);
  }
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int cur = c & 1;
    if (c + 1 < nchunks) {
      load_tile<A_K>(ra, A, lda, m0, kbeg + (c + 1) * BK, M, kend, tid, vecA);
      load_tile<B_K>(rb, B, ldb, n0, kbeg + (c + 1) * BK, N, kend, tid, vecB);
    }
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = wm + mi * 16 + g;
        const float2 p0 = tile_pair<A_K>(as, r, kk, t), p1 = tile_pair<A_K>(as, r + 8, kk, t);
        const float v[4] = {p0.x, p1.x, p0.y, p1.y};   // (r,t) (r+8,t) (r,t+4) (r+8,t+4)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (PRECISE) {
            ah[mi][q] = to_tf32(v[q]);
            al[mi][q] = to_tf32(v[q] - __uint_as_float(ah[mi][q]));
          } else {
            ah[mi][q] = __float_as_uint(v[q]);   // rounded to TF32 when the tile was stored
          }
        }
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const float2 p = tile_pair<B_K>(bs, wn + ni * 8 + g, kk, t);
        const float v[2] = {p.x, p.y};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (PRECISE) {
            bh[ni][q] = to_tf32(v[q]);
            bl[ni][q] = to_tf32(v[q] - __uint_as_float(bh[ni][q]));
          } else {
            bh[ni][q] = __float_as_uint(v[q]);
          }
        }
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          if (PRECISE) {
            mma_tf32(acc[mi][ni], al[mi], bh[ni]);
            mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
          }
          mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
        }
    }
    if (c + 1 < nchunks) {
      store_tile<A_K, !PRECISE>(As[cur ^ 1], ra, tid);
      store_tile<B_K, !PRECISE>(Bs[cur ^ 1], rb, tid);
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------------
  // Accumulator fragments go through shared memory (the operand tiles are dead) so that the per-element
  // epilogue is a short ROLLED loop with coalesced rows: unrolled over the 32 fragment values it was
  // thousands of instructions, and these short GEMMs ran instruction-cache bound.
  float* Cs = &As[0][0];                 // [BM][BN + 1] floats = 16.6 KB <= sizeof(As)
  constexpr int CS = BN + 1;
  static_assert(BM * CS <= 2 * TILE_FLOATS, "C staging tile must fit the A operand buffers");
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        Cs[(wm + mi * 16 + g + ((q & 2) ? 8 : 0)) * CS + wn + ni * 8 + 2 * t + (q & 1)] = acc[mi][ni][q];
  __syncthreads();
  const bool add_bias = epi.bias != nullptr && blockIdx.z == 0;
  // thread -> (row r = tid / 2, 32-column half h = tid % 2): the p == 0.5 keep bits of its 32 outputs (and of the
  // fused backward factor) come from ONE Philox draw each instead of one per element
  {
    const int r = tid >> 1, h = tid & 1;
    const int m = m0 + r, nb = n0 + 32 * h;
    if (m < M && nb < N) {
      uint32_t kw = 0xFFFFFFFFu, gkw = 0xFFFFFFFFu;
      if (epi.drop && epi.dc.half) kw = drop_word32(epi.dc, epi.stream, (uint64_t)m, (uint32_t)(nb >> 5));
      if (epi.gy != nullptr && epi.g_drop && epi.gdc.half) gkw = drop_word32(epi.gdc, epi.gstream, (uint64_t)m, (uint32_t)(nb >> 5));
      auto finish = [&](float v, int j) {   // everything after the GEMM for output element (m, nb + j)
        const int n = nb + j;
        v *= epi.scale;
        if (add_bias) v += epi.bias[n];
        if (epi.act) v = lrelu(v, epi.alpha);
        if (epi.drop) {
          const bool keep = epi.dc.half ? ((kw >> j) & 1u) : drop_keep_call(epi.dc, epi.stream, (uint64_t)m, (uint32_t)n);
          v = keep ? v * epi.dc.scale : 0.f;
        }
        if (epi.gy != nullptr) {
          const float y = epi.gy[(size_t)m * epi.ldgy + n];
          float gfac = epi.g_act ? lrelu_grad_from_out(y, epi.alpha) : 1.f;
          if (epi.g_drop) {
            const bool keep = epi.gdc.half ? ((gkw >> j) & 1u) : drop_keep_call(epi.gdc, epi.gstream, (uint64_t)m, (uint32_t)n);
            gfac = keep ? gfac * epi.gdc.scale : 0.f;
          }
          v *= gfac;
        }
        return v;
      };
      const float* cs = Cs + r * CS + 32 * h;
      float* crow = C + (size_t)m * ldc + nb;
      if (vecC && !epi.atomic) {   // N % 4 == 0, 16-byte aligned rows
#pragma unroll 2
        for (int j = 0; j < 32 && nb + j < N; j += 4) {
          float4 o = make_float4(finish(cs[j], j), finish(cs[j + 1], j + 1), finish(cs[j + 2], j + 2), finish(cs[j + 3], j + 3));
          float4* dst = reinterpret_cast<float4*>(crow + j);
          if (epi.accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      } else {
        for (int j = 0; j < 32 && nb + j < N; ++j) {
          const float v = finish(cs[j], j);
          if (epi.atomic) atomicAdd(crow + j, v);
          else if (epi.accumulate) crow[j] += v;
          else crow[j] = v;
        }
      }
    }
  }
}

__global__ void colsum_kernel(const float* __restrict__ X, int ldx, int M, int N, float* __restrict__ out,
                              int rows_per_block) {
  // block (32 x 8): 32 consecutive columns, 8 row lanes
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int mbeg = blockIdx.y * rows_per_block;
  const int mend = min(M, mbeg + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = mbeg + threadIdx.y; m < mend; m += 8) s += X[(size_t)m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
    atomicAdd(out + n, tot);
  }
}

}  // namespace

int launch_gemm(bool a_k, bool b_k, bool precise, const float* A, int lda, const float* B, int ldb, float* C,
                int ldc, int M, int N, int K, const GemmEpi& epi_in, int split_k, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  GemmEpi epi = epi_in;
  if (split_k < 1) split_k = 1;
  int k_per_split = ((K + split_k - 1) / split_k + BK - 1) / BK * BK;
  if (k_per_split < BK) k_per_split = BK;
  split_k = (K + k_per_split - 1) / k_per_split;
  if (split_k < 1) split_k = 1;
  if (split_k > 1) {
    MPG_CHECK(epi.accumulate && !epi.act && !epi.drop && epi.gy == nullptr,
              "split-K GEMM needs a pure accumulate epilogue");
    epi.atomic = 1;
  }
  const bool vecA = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vecB = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const bool vecC = (N % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  dim3 grid(cdiv(N, BN), cdiv(M, BM), split_k), block(NT);
#define MPG_GEMM_CASE(AK, BK_, PR)                                                                    \
  if (a_k == AK && b_k == BK_ && precise == PR)                                                       \
    gemm_kernel<AK, BK_, PR><<<grid, block, 0, stream>>>(A, lda, B, ldb, C, ldc, M, N, K, k_per_split, \
                                                         epi, vecA, vecB, vecC);
  MPG_GEMM_CASE(true, true, false)
  MPG_GEMM_CASE(true, true, true)
  MPG_GEMM_CASE(true, false, false)
  MPG_GEMM_CASE(true, false, true)
  MPG_GEMM_CASE(false, false, false)
  MPG_GEMM_CASE(false, false, true)
  MPG_GEMM_CASE(false, true, false)
  MPG_GEMM_CASE(false, true, true)
#undef MPG_GEMM_CASE
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_colsum(const float* X, int ldx, int M, int N, float* out, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  const int rows_per_block = 256;
  dim3 grid(cdiv(N, 32), cdiv(M, rows_per_block)), block(32, 8);
  colsum_kernel<<<grid, block, 0, stream>>>(X, ldx, M, N, out, rows_per_block);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
