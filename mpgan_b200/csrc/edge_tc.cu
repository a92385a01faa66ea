// tcgen05 / TMEM edge network for the default MPGAN architecture (fe = 96 -> 160 -> 192).
//
// Work unit ("step"): 128 receivers (rows r = b*N + i of the flattened node list) x one sender
// index s of each row's own jet.  A CTA owns a contiguous range of (tile, s) steps and keeps the
// running neighbour sum of its 128 receivers in registers, so the reduction over senders is a
// plain in-thread add and the [B*N*N, hidden] tensors exist only as bf16 tiles in shared memory:
//
//   H0[r,:]  = lrelu(P[r,:] + Q[jet(r)*N + s,:])               fp32 add, bf16 tile (A operand)
//   D1       = H0 * W1^T   (tcgen05.mma, M=128 N=160 K=96+16)   fp32 in TMEM cols [0,160)
//   H1       = lrelu(D1)                                        bf16 tile (A operand)
//   D2       = H1 * W2^T   (tcgen05.mma, M=128 N=192 K=160+16)  fp32 in TMEM cols [256,448)
//   acc[r,:] += mask[jet(r), s] * lrelu(D2)                     fp32 registers
//
// The biases ride in an extra K=16 step: the A tile carries two constant 1.0 columns and the weight
// image carries bias_hi / bias_lo (bf16 split of the fp32 bias), so the epilogues are pure lrelu.
// Dropout (p = 0.5 only on this path) uses the same Philox bits as the generic kernel (one call per
// quarter-row, common.cuh: edge_drop_*); its 2x scale is folded into the next layer's weight image.
//
// 16 warps: thread <-> (TMEM lane = tile row, column quarter).  Four warps per scheduler hide the
// TMEM-load / shared-store latencies of the epilogues; thread 0 additionally issues every
// tcgen05.mma and warp 0 owns the TMEM allocation.  Weights reach shared memory once per CTA with two
// bulk-async (TMA) copies of a pre-swizzled bf16 image.  mbarrier pipeline per step:
//   h0_full -> MMA1 -> d1_full -> e1 -> h1_full -> MMA2 -> d2_full -> e2,
// with H0(s+1) built while MMA2(s) runs and e2(s) overlapping MMA1(s+1).
#include "edge.cuh"

namespace mpg {
namespace {

constexpr int K0 = 96, N1 = 160, N2 = 192;
constexpr int TILE = 128;
constexpr int NQ = 4;                            // column quarters (warps sharing a TMEM lane group)
constexpr int Q0 = K0 / NQ, Q1 = N1 / NQ, Q2 = N2 / NQ;   // 24, 40, 48 columns per thread
constexpr int KSTEPS1 = K0 / 16 + 1;             // + bias step
constexpr int KSTEPS2 = N1 / 16 + 1;
constexpr uint32_t W1_BLK = N1 * 128;            // bytes of one 64-wide K block of W1 (rows = out features)
constexpr uint32_t W2_BLK = N2 * 128;
constexpr uint32_t A_BLK = TILE * 128;           // one 64-wide K block of an activation tile
constexpr uint32_t W1_BYTES = 2 * W1_BLK;        // K = 128 (96 + bias step, padded)
constexpr uint32_t W2_BYTES = 3 * W2_BLK;        // K = 192 (160 + bias step, padded)
constexpr uint32_t H0_BYTES = 2 * A_BLK;
constexpr uint32_t H1_BYTES = 3 * A_BLK;
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = OFF_W1 + W1_BYTES;   // 40960
constexpr uint32_t OFF_H0 = OFF_W2 + W2_BYTES;   // 114688
constexpr uint32_t OFF_H1 = OFF_H0 + H0_BYTES;   // 147456
constexpr uint32_t OFF_BAR = OFF_H1 + H1_BYTES;  // 196608
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128 + 1024;   // + barriers + alignment slack
constexpr int NTHREADS = 128 * NQ;
constexpr uint32_t TMEM_COLS = 512, D1_COL = 0, D2_COL = 256;

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle: 8-row groups 1024 B apart, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc(int N) {
  // c = f32 (1 << 4), a = b = bf16 (1 << 7, 1 << 10), K-major A and B, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}
// 16 / 8 consecutive fp32 columns of this thread's TMEM lane, into v[o..]
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// NC (multiple of 8) columns starting at taddr
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
#pragma unroll
  for (int c = 0; c + 16 <= NC; c += 16) tmem_ld16(taddr + c, v + c);
  if (NC % 16) tmem_ld8(taddr + (NC / 16) * 16, v + (NC / 16) * 16);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
// byte offset of the 16-byte chunk holding columns [k, k+8) of `row` inside a swizzled tile whose
// 64-wide K blocks are blk_bytes apart
__device__ __forceinline__ uint32_t swz_chunk(uint32_t row, uint32_t k, uint32_t blk_bytes) {
  const uint32_t blk = k >> 6, chunk = (k & 63) >> 3;
  return blk * blk_bytes + row * 128 + ((chunk ^ (row & 7)) << 4);
}
// store 8 consecutive columns (one 16-byte chunk) of a row as bf16
__device__ __forceinline__ void st_chunk(uint32_t addr, const float* v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16(v[0], v[1])),
               "r"(pack_bf16(v[2], v[3])), "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])));
}
__device__ __forceinline__ void st_ones_chunk(uint32_t addr) {   // {1, 1, 0, 0, 0, 0, 0, 0}
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%2,%2};" ::"r"(addr), "r"(0x3F803F80u), "r"(0u));
}
__device__ __forceinline__ void st_zero_chunk(uint32_t addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u));
}
// keep bit `b` (0..127) of a 128-bit Philox draw as an all-ones / all-zeros word
__device__ __forceinline__ uint32_t keep_mask(const u4& bits, int b) {
  const uint32_t w = (b >> 5) == 0 ? bits.x : ((b >> 5) == 1 ? bits.y : ((b >> 5) == 2 ? bits.z : bits.w));
  int32_t m;
  asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"(w), "r"(b & 31));
  return (uint32_t)m;
}
__device__ __forceinline__ float apply_keep(float x, uint32_t mask) { return __uint_as_float(__float_as_uint(x) & mask); }

struct TcArgs {
  EdgeArgs a;
  const uint8_t* w1img;   // pre-swizzled bf16 images (see prep kernel)
  const uint8_t* w2img;
  int num_tiles;
  long long total_steps;
};

// ---------------------------------------------------------------------------------------------------
// weight image: img[n][k] (K-major, SW128) = bf16(scale * W[n][k]) for k < K, bias_hi / bias_lo at
// k = K, K+1, zero elsewhere.
// ---------------------------------------------------------------------------------------------------
__global__ void weight_image_kernel(const float* __restrict__ W, const float* __restrict__ bias, int Nout, int K,
                                    int Kpad, float scale, uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Nout * Kpad) return;
  const int n = idx / Kpad, k = idx % Kpad;
  float v = 0.f;
  if (k < K) v = W[(size_t)n * K + k] * scale;
  else if (k == K) v = bias[n];
  else if (k == K + 1) v = bias[n] - __bfloat162float(__float2bfloat16_rn(bias[n]));
  const uint32_t off = (uint32_t)(k >> 6) * (uint32_t)Nout * 128u + (uint32_t)n * 128u +
                       ((((uint32_t)(k & 63) >> 3) ^ ((uint32_t)n & 7u)) << 4) + (uint32_t)(k & 7) * 2u;
  *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------
template <bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1) edge_tc_fwd_kernel(TcArgs t) {
  extern __shared__ uint8_t smem_raw[];
  const EdgeArgs& a = t.a;
  // 1024-byte alignment for the 128B-swizzled tiles
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW1 = base + OFF_W1, sW2 = base + OFF_W2, sH0 = base + OFF_H0, sH1 = base + OFF_H1;
  const uint32_t bar_w = base + OFF_BAR, bar_h0 = bar_w + 8, bar_d1 = bar_w + 16, bar_h1 = bar_w + 24,
                 bar_d2 = bar_w + 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DropCfg drop = a.drop;
  if (DROP) resolve_seed(drop);

  // contiguous range of (tile, sender) steps for this CTA
  const long long g0 = t.total_steps * blockIdx.x / gridDim.x;
  const long long g1 = t.total_steps * (blockIdx.x + 1) / gridDim.x;
  const int nsteps = (int)(g1 - g0);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_h0, NTHREADS);
    mbar_init(bar_d1, 1);
    mbar_init(bar_h1, NTHREADS);
    mbar_init(bar_d2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool issuer = threadIdx.x == 0;   // this thread also issues every tcgen05.mma

  constexpr uint32_t idesc1 = umma_idesc(N1), idesc2 = umma_idesc(N2);
  auto issue1 = [&](int parity) {       // D1 = H0 * W1^T once every thread has published its H0 columns
    mbar_wait(bar_h0, parity);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < KSTEPS1; ++ks) {
      const uint32_t blk = ks >> 2, j = ks & 3;
      umma_bf16(tmem + D1_COL, umma_desc(sH0 + blk * A_BLK + j * 32), umma_desc(sW1 + blk * W1_BLK + j * 32),
                idesc1, ks > 0);
    }
    umma_commit(bar_d1);
  };
  auto issue2 = [&](int parity) {       // D2 = H1 * W2^T
    mbar_wait(bar_h1, parity);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < KSTEPS2; ++ks) {
      const uint32_t blk = ks >> 2, j = ks & 3;
      umma_bf16(tmem + D2_COL, umma_desc(sH1 + blk * A_BLK + j * 32), umma_desc(sW2 + blk * W2_BLK + j * 32),
                idesc2, ks > 0);
    }
    umma_commit(bar_d2);
  };
  if (issuer && nsteps > 0) {
    mbar_expect_tx(bar_w, W1_BYTES + W2_BYTES);
    bulk_g2s(sW1, t.w1img, W1_BYTES, bar_w);
    bulk_g2s(sW2, t.w2img, W2_BYTES, bar_w);
  }

  const int q = warp >> 2;                        // column quarter
  const int row = (warp & 3) * 32 + lane;         // tile row == TMEM lane
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
  const int BN = a.B * a.N;
  const float dscale = DROP ? 2.f : 1.f;          // scale of layer-2 output (layers 0/1: folded into weights)

  // constant 1.0 columns of the bias K-step (cols 96,97 of H0; 160,161 of H1), zeros after them
  if (q == 0) {
    st_ones_chunk(sH0 + swz_chunk(row, 96, A_BLK));
    st_zero_chunk(sH0 + swz_chunk(row, 104, A_BLK));
  } else if (q == 1) {
    st_ones_chunk(sH1 + swz_chunk(row, 160, A_BLK));
    st_zero_chunk(sH1 + swz_chunk(row, 168, A_BLK));
  }

  float acc[Q2];
  float Preg[Q0];
  int cur_tile = -1;      // tile whose P row is in Preg
  int acc_tile = -1;      // tile the accumulators belong to
  int r = 0, jet = 0;
  bool valid = false;
  u4 bits_next{0, 0, 0, 0};   // Philox keep bits of the step whose H0 was built last

  auto load_tile = [&](int tile) {
    cur_tile = tile;
    r = tile * TILE + row;
    valid = r < BN;
    const int rc = valid ? r : BN - 1;
    jet = rc / a.N;
    const float4* p = reinterpret_cast<const float4*>(a.P + (size_t)rc * K0 + q * Q0);
#pragma unroll
    for (int c = 0; c < Q0 / 4; ++c) {
      const float4 v = __ldg(p + c);
      Preg[4 * c] = v.x; Preg[4 * c + 1] = v.y; Preg[4 * c + 2] = v.z; Preg[4 * c + 3] = v.w;
    }
  };
  auto build_h0 = [&](long long g) {
    const int tile = (int)(g / a.N), s = (int)(g % a.N);
    if (tile != cur_tile) load_tile(tile);
    const float4* qp = reinterpret_cast<const float4*>(a.Q + ((size_t)jet * a.N + s) * K0 + q * Q0);
    if (DROP) bits_next = edge_drop_bits(drop.seed, (uint64_t)(valid ? r : 0) * a.N + s, q);
#pragma unroll
    for (int c8 = 0; c8 < Q0 / 8; ++c8) {
      const float4 q0 = __ldg(qp + 2 * c8), q1 = __ldg(qp + 2 * c8 + 1);
      float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float x = v[e] + Preg[8 * c8 + e];
        x = fmaxf(x, a.alpha * x);
        if (DROP) x = apply_keep(x, keep_mask(bits_next, c8 * 8 + e));
        v[e] = x;
      }
      st_chunk(sH0 + swz_chunk(row, q * Q0 + c8 * 8, A_BLK), v);
    }
  };
  auto flush = [&]() {
    if (acc_tile >= 0) {
      const int fr = acc_tile * TILE + row;
      if (fr < BN) {
        float* dst = a.agg + (size_t)fr * N2 + q * Q2;
#pragma unroll
        for (int c = 0; c < Q2; ++c) atomicAdd(dst + c, acc[c] * a.out_scale * dscale);
      }
    }
  };

  if (nsteps > 0) {
    build_h0(g0);
    fence_async_smem();
    mbar_arrive(bar_h0);
    if (issuer) {
      mbar_wait(bar_w, 0);   // weight images landed
      issue1(0);
    }
  }
  for (int it = 0; it < nsteps; ++it) {
    const long long g = g0 + it;
    const int tile = (int)(g / a.N), s = (int)(g % a.N);
    // rows of the tile this step belongs to (P registers may already hold the next tile's rows)
    const int er = tile * TILE + row;
    const bool evalid = er < BN;
    const int ejet = (evalid ? er : BN - 1) / a.N;
    const u4 bits = bits_next;   // keep bits of THIS step (build_h0 below overwrites bits_next)
    if (tile != acc_tile) {
      flush();
      acc_tile = tile;
#pragma unroll
      for (int c = 0; c < Q2; ++c) acc[c] = 0.f;
    }
    // ---- e1: D1 -> H1 --------------------------------------------------------------------------
    mbar_wait(bar_d1, it & 1);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < Q1; c0 += 16) {
      const int n = (Q1 - c0) >= 16 ? 16 : 8;   // static after unrolling
      float v[16];
      if (n == 16) tmem_ld16(tmem + tlane + D1_COL + q * Q1 + c0, v);
      else tmem_ld8(tmem + tlane + D1_COL + q * Q1 + c0, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (e < n) {
          float x = fmaxf(v[e], a.alpha * v[e]);
          if (DROP) x = apply_keep(x, keep_mask(bits, Q0 + c0 + e));
          v[e] = x;
        }
      }
      st_chunk(sH1 + swz_chunk(row, q * Q1 + c0, A_BLK), v);
      if (n == 16) st_chunk(sH1 + swz_chunk(row, q * Q1 + c0 + 8, A_BLK), v + 8);
    }
    fence_async_smem();
    tc_fence_before();
    mbar_arrive(bar_h1);
    if (issuer) issue2(it & 1);
    // ---- H0 of the next step while MMA2 runs ------------------------------------------------------
    if (it + 1 < nsteps) {
      build_h0(g + 1);
      fence_async_smem();
      mbar_arrive(bar_h0);
      if (issuer) issue1((it + 1) & 1);
    }
    // ---- e2: D2 -> masked accumulate -------------------------------------------------------------
    const float m = (evalid ? (a.mask ? a.mask[(size_t)ejet * a.N + s] : 1.f) : 0.f);
    mbar_wait(bar_d2, it & 1);
    tc_fence_after();
#pragma unroll
    for (int c16 = 0; c16 < Q2 / 16; ++c16) {
      float v[16];
      tmem_ld16(tmem + tlane + D2_COL + q * Q2 + c16 * 16, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float x = fmaxf(v[e], a.alpha * v[e]);
        if (DROP) x = apply_keep(x, keep_mask(bits, Q0 + Q1 + c16 * 16 + e));
        acc[c16 * 16 + e] = fmaf(x, m, acc[c16 * 16 + e]);
      }
    }
    tc_fence_before();
  }
  flush();

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

#include "edge_tc_bwd.cuh"

}  // namespace

int edge_tc_features() { return 3; }

// one-shot timing probes: bench.py arms a pair of CUDA events per kernel id (1 = forward,
// 2 = backward CHAIN, 3 = backward DW2); the next launch of that kernel is bracketed by them
static thread_local cudaEvent_t g_probe[4][2] = {};
void edge_tc_arm_probe(int id, cudaEvent_t e0, cudaEvent_t e1) {
  if (id >= 1 && id <= 3) { g_probe[id][0] = e0; g_probe[id][1] = e1; }
}
struct ProbeScope {
  int id; cudaStream_t s; bool on;
  ProbeScope(int id_, cudaStream_t s_) : id(id_), s(s_), on(g_probe[id_][0] != nullptr) {
    if (on) cudaEventRecord(g_probe[id][0], s);
  }
  ~ProbeScope() {
    if (on) { cudaEventRecord(g_probe[id][1], s); g_probe[id][0] = g_probe[id][1] = nullptr; }
  }
};

bool edge_tc_supported(const EdgeArgs& a) {
  return a.H0 == K0 && a.H1 == N1 && a.H2 == N2 && a.n_ef == 0 && a.N >= 2 &&
         (a.drop.p == 0.f || a.drop.p == 0.5f) && a.alpha > 0.f && a.alpha < 1.f;
}

size_t edge_tc_workspace_bytes(int B, int N, int H0, int H1, int H2) {
  if (H0 != K0 || H1 != N1 || H2 != N2) return 0;
  return W1_BYTES + W2_BYTES + 1024;
}

static int tc_prepare(const EdgeArgs& a, void* ws, TcArgs& t, int* grid, cudaStream_t stream) {
  MPG_CHECK(edge_tc_supported(a), "edge_tc: unsupported configuration");
  uint8_t* img = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  t.a = a;
  t.w1img = img;
  t.w2img = img + W1_BYTES;
  const float s = a.drop.p > 0.f ? 2.f : 1.f;  // dropout scale of the previous layer folded into the weights
  weight_image_kernel<<<cdiv(N1 * 128, 256), 256, 0, stream>>>(a.W1, a.b1, N1, K0, 128, s, img);
  MPG_LAUNCH_CHECK();
  weight_image_kernel<<<cdiv(N2 * 192, 256), 256, 0, stream>>>(a.W2, a.b2, N2, N1, 192, s, img + W1_BYTES);
  MPG_LAUNCH_CHECK();
  const long long BN = (long long)a.B * a.N;
  t.num_tiles = (int)((BN + TILE - 1) / TILE);
  t.total_steps = (long long)t.num_tiles * a.N;
  int dev = 0, sms = 148;
  MPG_CUDA(cudaGetDevice(&dev));
  MPG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  *grid = (int)(t.total_steps < sms ? t.total_steps : sms);
  return 0;
}

int launch_edge_tc_fwd(const EdgeArgs& a, void* ws, cudaStream_t stream) {
  TcArgs t;
  int grid = 1;
  if (tc_prepare(a, ws, t, &grid, stream)) return 1;
  MPG_CUDA(cudaMemsetAsync(a.agg, 0, (size_t)a.B * a.N * N2 * sizeof(float), stream));
  if (a.drop.p > 0.f) {
    MPG_CUDA(cudaFuncSetAttribute(edge_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ProbeScope probe(1, stream);
    edge_tc_fwd_kernel<true><<<grid, NTHREADS, SMEM_BYTES, stream>>>(t);
  } else {
    MPG_CUDA(cudaFuncSetAttribute(edge_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ProbeScope probe(1, stream);
    edge_tc_fwd_kernel<false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(t);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

template <int MODE, bool DROP>
static int launch_bwd_one(const TcArgs& t, int grid, uint32_t smem, cudaStream_t stream) {
  MPG_CUDA(cudaFuncSetAttribute(edge_tc_bwd_kernel<MODE, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProbeScope probe(MODE == BWD_CHAIN ? 2 : 3, stream);
    edge_tc_bwd_kernel<MODE, DROP><<<grid, NTHREADS, smem, stream>>>(t);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

// dP and dQ must be zeroed by the caller (both are accumulated with atomics here)
int launch_edge_tc_bwd(const EdgeArgs& a, void* ws, cudaStream_t stream) {
  TcArgs t;
  int grid = 1;
  if (tc_prepare(a, ws, t, &grid, stream)) return 1;
  if (a.drop.p > 0.f) {
    if (launch_bwd_one<BWD_CHAIN, true>(t, grid, BW_SMEM_CHAIN, stream)) return 1;
    if (launch_bwd_one<BWD_DW2, true>(t, grid, BW_SMEM_DW2, stream)) return 1;
  } else {
    if (launch_bwd_one<BWD_CHAIN, false>(t, grid, BW_SMEM_CHAIN, stream)) return 1;
    if (launch_bwd_one<BWD_DW2, false>(t, grid, BW_SMEM_DW2, stream)) return 1;
  }
  return 0;
}

}  // namespace mpg
