// tcgen05 / TMEM node network (fn): the three LinearNet layers of an MPLayer's node MLP
// (reference mpgan/model.py:70-85 applied at :268-279, cat(agg, x) -> 256 -> 256 -> out) as ONE kernel per
// 128-row tile, forward and input-gradient backward.  TF32 operands (kind::tf32), fp32 accumulate.
//
//   forward    h  = [a | b]                       (a = agg [M,Ka], b = x [M,Kb]: the cat is never materialised)
//              y0 = drop(lrelu(h  W0^T + b0))     -> HBM (saved for backward) and, TF32-rounded, back IN PLACE
//              y1 = drop(lrelu(y0 W1^T + b1))        into the TMEM columns it was read from: the next GEMM
//              out = drop(y1 W2^T + b2)              takes its A operand from TMEM
//   backward   dz2 = dout * drop'                 (input of the chain: built in shared memory)
//              dz1 = (dz2 W2) * lrelu'(y1) drop'  -> HBM (the weight-gradient GEMMs read it) and TMEM in place
//              dz0 = (dz1 W1) * lrelu'(y0) drop'  -> HBM and TMEM in place
//              [da | db] = dz0 W0                 -> HBM
//
// Both directions are the same chain "A0 (smem) x B0 -> epi -> (TMEM) x B1 -> epi -> (TMEM) x B2 -> store";
// the B operands are pre-swizzled TF32 weight images (fn_image_kernel; transposed ones for backward), streamed
// through a 3-stage ring of 32-wide K blocks by a loader warp with bulk-async copies.  16 epilogue warps
// (TMEM lane quarter x column quarter), one MMA-issuer warp, one loader warp.  TMEM: D0 [0,256) D1 [256,512),
// D2 over D0.  Weight gradients (dz^T y) stay on the TF32 mma.sync GEMM (gemm.cu), launched by the caller on a
// side stream.
#include "fn_tc.cuh"

namespace mpg {
namespace {

__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}


#define MPG_TC_WRAPPERS_ONLY
#include "edge_tc_common.cuh"   // PTX wrappers (mbarrier, bulk copies, tcgen05 alloc/commit/ld/st), umma_desc

constexpr int FN_EPI = 512;                       // epilogue threads (16 warps)
constexpr int FN_THREADS = FN_EPI + 64;           // + warp 16 (MMA issuer) + warp 17 (loader)
constexpr uint32_t FN_ABLK = 128 * 128;           // one 32-wide K block of the A tile: 128 rows x 128 B
constexpr uint32_t FN_STAGE = 256 * 128;          // one 32-wide K block of a weight image: <= 256 rows x 128 B
// Shared memory.  [0, 128 KB) is the staging tile through which every [128 x H] activation / gradient tile
// moves between the thread-per-row world of TMEM and coalesced global accesses (rows of H floats, 16-byte chunks
// XOR-swizzled with row & 7).  Forward: the same 128 KB first hold the A tile of GEMM 0 (up to 8 K blocks) and
// the ring has 3 stages; backward: the A tile is one K block (NO <= 32) behind the staging tile and the ring has
// 2 stages.
constexpr uint32_t FN_OFF_STG = 0;
constexpr uint32_t FN_STG_BYTES = 128 * 256 * 4;                    // 131072
constexpr uint32_t FN_OFF_BAR = FN_STG_BYTES + 3 * FN_STAGE;        // 229376
constexpr uint32_t FN_OFF_CSUM = FN_OFF_BAR + 256;                  // 256 floats: column sums of dz2 (backward)
constexpr uint32_t FN_SMEM = FN_OFF_CSUM + 1024 + 1024;             // 231680 <= 232448
constexpr int FN_MAX_STAGES = 3;

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// M = 128, K-major A and B, TF32 operands, fp32 accumulate
__device__ __forceinline__ uint32_t idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A = [128 x 8] TF32 slice held in TMEM (lane = row, one K element per 32-bit column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
        "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
        "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
        "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Weight images: img[n][k] = tf32(W[n][k]) (or W[k][n] when transposed), zero padded to `rows` x 32*kblocks,
// as 32-wide K blocks of `rows` x 128 B in the 128-byte-swizzled K-major layout tcgen05.mma reads.
// One thread per 16-byte chunk; one launch builds the three images of a direction.
// ---------------------------------------------------------------------------------------------------
__global__ void fn_image_kernel(FnImageJobs jobs) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const FnImageJob& jb = jobs.job[j];
    const int chunks = jb.rows * jb.kblocks * 8;
    if (idx < chunks) {
      const int n = idx / (jb.kblocks * 8), c = idx % (jb.kblocks * 8);
      const int k0 = c * 4;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e;
        float x = 0.f;
        if (!jb.transposed) {
          if (n < jb.R && k < jb.C) x = jb.W[(size_t)n * jb.ldw + k];
        } else {
          if (k < jb.R && n < jb.C) x = jb.W[(size_t)k * jb.ldw + n];
        }
        v[e] = tf32_rna(x);
      }
      const uint32_t off = (uint32_t)(c >> 3) * (uint32_t)jb.rows * 128u + (uint32_t)n * 128u +
                           ((((uint32_t)c & 7u) ^ ((uint32_t)n & 7u)) << 4);
      *reinterpret_cast<float4*>(jb.dst + off) = make_float4(v[0], v[1], v[2], v[3]);
      return;
    }
    idx -= chunks;
  }
}

// ---------------------------------------------------------------------------------------------------
// The chain kernel
// ---------------------------------------------------------------------------------------------------
// optional event trace (profiles/trace_fn.py): globaltimer stamps of CTA 0 -- slots 0..15 epilogue thread 0,
// 16..31 MMA issuer, 32..47 loader
__device__ long long* g_fn_trace = nullptr;
__device__ __forceinline__ void fn_stamp(int slot) {
  if (g_fn_trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    long long tnow;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
    g_fn_trace[slot] = tnow;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(FN_THREADS, 1) fn_tc_kernel(FnTcArgs t) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NSTG = BWD ? 2 : 3;
  constexpr uint32_t OFF_A = BWD ? FN_STG_BYTES : 0u;                       // A tile of GEMM 0
  constexpr uint32_t OFF_W = BWD ? FN_STG_BYTES + FN_ABLK : FN_STG_BYTES;   // weight ring
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SW128 tiles need 1024-byte alignment
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base + OFF_A, sW = base + OFF_W, sS = base + FN_OFF_STG, bar0 = base + FN_OFF_BAR;
  const uint32_t bar_a = bar0;                   // A tile built (FN_EPI arrivals)
  const uint32_t bar_e = bar0 + 8;               // [2] epilogue l done: TMEM A operand of layer l+1 ready
  const uint32_t bar_d = bar0 + 24;              // [3] accumulator of layer l complete
  const uint32_t bar_full = bar0 + 48;           // [FN_MAX_STAGES]
  const uint32_t bar_empty = bar0 + 48 + 8 * FN_MAX_STAGES;
  float* csum = reinterpret_cast<float*>(sm + FN_OFF_CSUM);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + FN_OFF_BAR + 192);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 128;
  if (warp == 1) fn_stamp(15);

  if (threadIdx.x == 0) {
    mbar_init(bar_a, FN_EPI);
    mbar_init(bar_e, FN_EPI);
    mbar_init(bar_e + 8, FN_EPI);
    for (int l = 0; l < 3; ++l) mbar_init(bar_d + 8 * l, 1);
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (BWD && threadIdx.x < 256) csum[threadIdx.x] = 0.f;
  if (warp == 16) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 17) {
    // =============================== loader: weight-image K blocks through the ring ===================
    uint32_t g = 0;
    for (int l = 0; l < 3; ++l) {
      const uint32_t bytes = (uint32_t)t.n[l] * 128u;
      const int kb = (t.kmma[l] + 3) >> 2;
      for (int blk = 0; blk < kb; ++blk, ++g) {
        const uint32_t st = g % NSTG;
        if (g >= NSTG) mbar_wait(bar_empty + 8 * st, (g / NSTG - 1) & 1);
        mbar_expect_tx_elect(bar_full + 8 * st, bytes);
        bulk_g2s_elect(sW + st * FN_STAGE, t.img[l] + (size_t)blk * bytes, bytes, bar_full + 8 * st);
      }
    }
  } else if (warp == 16) {
    // =============================== MMA issuer ========================================================
    uint32_t g = 0;
    for (int l = 0; l < 3; ++l) {
      mbar_wait(l == 0 ? bar_a : bar_e + 8 * (l - 1), 0);
      tc_fence_after();
      fn_stamp(16 + 2 * l);
      const uint32_t idesc = idesc_tf32(t.n[l]);
      const uint32_t dcol = tmem + (l == 1 ? 256u : 0u);
      const uint32_t acol = tmem + (l == 1 ? 0u : 256u);   // TMEM A operand (layers 1, 2)
      const int kb = (t.kmma[l] + 3) >> 2;
      for (int blk = 0; blk < kb; ++blk, ++g) {
        const uint32_t st = g % NSTG;
        mbar_wait(bar_full + 8 * st, (g / NSTG) & 1);
        tc_fence_after();
        const int nm = min(4, t.kmma[l] - 4 * blk);
        if (elect_one()) {
          const uint64_t bd = umma_desc(sW + st * FN_STAGE);
          if (l == 0) {
            const uint64_t ad = umma_desc(sA + (uint32_t)blk * FN_ABLK);
            for (int j = 0; j < nm; ++j) umma_tf32_ss(dcol, ad + (uint64_t)(j * 2), bd + (uint64_t)(j * 2), idesc, (uint32_t)(blk | j));
          } else {
            for (int j = 0; j < nm; ++j)
              umma_tf32_ts(dcol, acol + (uint32_t)(blk * 32 + j * 8), bd + (uint64_t)(j * 2), idesc, (uint32_t)(blk | j));
          }
          umma_commit(bar_empty + 8 * st);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar_d + 8 * l);
      __syncwarp();
      fn_stamp(17 + 2 * l);
    }
  } else {
    // =============================== epilogue warps ====================================================
    const int q = warp >> 2;                       // column quarter
    const int row = (warp & 3) * 32 + lane;        // tile row == TMEM lane
    const int grow = row0 + row;
    const bool valid = grow < t.M;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    DropCfg dc = t.drop;
    resolve_seed(dc);
    const bool drop = dc.p > 0.f;                  // p == 0.5 only (host-checked)
    if (warp == 0) fn_stamp(0);

    // Staging tile <-> global, coalesced: warp w moves rows w, w+16, ...; a lane moves the 16-byte chunks lane,
    // lane+32 of a row (512 contiguous bytes per warp instruction).  Chunk c of row r sits at chunk c ^ (r & 7).
    auto stage_to_global = [&](float* dst, int H) {
      const int cpr = H >> 2;                      // chunks per row
      for (int r = warp; r < 128; r += 16) {
        if (row0 + r >= t.M) break;
        float* o = dst + (size_t)(row0 + r) * H;
        for (int c = lane; c < cpr; c += 32) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(sS + (uint32_t)r * (uint32_t)(H * 4) + (((uint32_t)c ^ ((uint32_t)r & 7u)) << 4)));
          *reinterpret_cast<float4*>(o + 4 * c) = v;
        }
      }
    };
    auto global_to_stage = [&](const float* src, int H) {
      const int cpr = H >> 2;
      for (int rb = warp; rb < 128; rb += 64) {    // 4 rows per batch: all loads in flight before the first store
        float4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + 16 * u;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c = lane + 32 * i;
            v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < t.M && c < cpr) v[u][i] = *reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * H + 4 * c);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + 16 * u;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c = lane + 32 * i;
            if (c < cpr)
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sS + (uint32_t)r * (uint32_t)(H * 4) + (((uint32_t)c ^ ((uint32_t)r & 7u)) << 4)),
                           "f"(v[u][i].x), "f"(v[u][i].y), "f"(v[u][i].z), "f"(v[u][i].w) : "memory");
          }
        }
      }
    };

    // ---- A tile of layer 0: [a | b] (forward) or dout * drop' (backward), TF32-rounded, swizzled ------
    {
      const int kb = (t.kmma[0] + 3) >> 2;
      const int nch = kb * 8;                      // 16-byte chunks per row
      const int K = t.Ka + t.Kb;
      const int total = 128 * nch;
      auto load_item = [&](int idx, float* v) {
        const int r = idx / nch, c = idx % nch;
        const int gr = row0 + r, k0 = c * 4;
        if (gr < t.M && k0 < K) {
          if (k0 + 3 < t.Ka && t.vec_a) {
            const float4 x = *reinterpret_cast<const float4*>(t.a + (size_t)gr * t.lda + k0);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
          } else if (k0 >= t.Ka && k0 + 3 < K && t.vec_b) {
            const float4 x = *reinterpret_cast<const float4*>(t.b + (size_t)gr * t.ldb + (k0 - t.Ka));
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = k0 + e;
              if (k < t.Ka) v[e] = t.a[(size_t)gr * t.lda + k];
              else if (k < K) v[e] = t.b[(size_t)gr * t.ldb + (k - t.Ka)];
            }
          }
        }
      };
      auto store_item = [&](int idx, float* v) {
        const int r = idx / nch, c = idx % nch;
        const int gr = row0 + r, k0 = c * 4;
        if (BWD && gr < t.M && k0 < K) {
          if (drop) {   // dz2 = dout * keep * 2 (dropout after the final linear layer, stream 2)
            const uint32_t kw = drop_word32(dc, t.stream[2], (uint64_t)gr, (uint32_t)(k0 >> 5));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[e] = ((kw >> ((k0 + e) & 31)) & 1u) ? v[e] * dc.scale : 0.f;
              if (k0 + e < K) t.dz2[(size_t)gr * K + k0 + e] = v[e];
            }
          }
          if (t.dbias[2] != nullptr) {   // db2 = column sums of dz2
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k0 + e < K) atomicAdd(&csum[k0 + e], v[e]);
          }
        }
        const uint32_t off = (uint32_t)(c >> 3) * FN_ABLK + (uint32_t)r * 128u + ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sA + off), "f"(tf32_rna(v[0])), "f"(tf32_rna(v[1])),
                     "f"(tf32_rna(v[2])), "f"(tf32_rna(v[3])) : "memory");
      };
      if (!BWD && t.vec_a && (t.vec_b || t.Kb == 0) && (K & 3) == 0) {
        // fast path (16-byte aligned sources, whole chunks): short code, 8 loads in flight per thread
        constexpr int NB = 8;
        for (int i0 = threadIdx.x; i0 < total; i0 += FN_EPI * NB) {
          float4 v[NB];
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            const int idx = i0 + u * FN_EPI;
            const int r = idx / nch, c = idx - r * nch;
            const int gr = row0 + r, k0 = c * 4;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < total && gr < t.M && k0 < K) {
              const float* src = k0 < t.Ka ? t.a + (size_t)gr * t.lda + k0 : t.b + (size_t)gr * t.ldb + (k0 - t.Ka);
              v[u] = *reinterpret_cast<const float4*>(src);
            }
          }
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            const int idx = i0 + u * FN_EPI;
            const int r = idx / nch, c = idx - r * nch;
            if (idx < total) {
              const uint32_t off = (uint32_t)(c >> 3) * FN_ABLK + (uint32_t)r * 128u + ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sA + off), "f"(tf32_rna(v[u].x)), "f"(tf32_rna(v[u].y)),
                           "f"(tf32_rna(v[u].z)), "f"(tf32_rna(v[u].w)) : "memory");
            }
          }
        }
      } else {
#pragma unroll 1
        for (int idx = threadIdx.x; idx < total; idx += FN_EPI) {
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          load_item(idx, v);
          store_item(idx, v);
        }
      }
      fence_async_smem();
      mbar_arrive(bar_a);
      if (warp == 0) fn_stamp(1);
      if (BWD && t.dbias[2] != nullptr) {
        named_bar_sync(1, FN_EPI);   // every epilogue thread's shared-memory atomics have landed
        if ((int)threadIdx.x < K) atomicAdd(t.dbias[2] + threadIdx.x, csum[threadIdx.x]);
      }
    }

    // ---- hidden layers: accumulator -> activation (or its derivative) -> staging tile + TMEM in place ------
#pragma unroll 1
    for (int l = 0; l < 2; ++l) {
      const int H = t.n[l], cw = H >> 2;           // this thread's columns: [q*cw, (q+1)*cw)
      const uint32_t dcol = tl + (l == 1 ? 256u : 0u);
      const uint32_t stream = t.stream[BWD ? 1 - l : l];
      float* dst = t.out01[l];
      const uint32_t srow = sS + (uint32_t)row * (uint32_t)(H * 4);
      if (BWD) {
        // the saved layer output whose derivative this epilogue applies: coalesced into the staging tile while the
        // GEMM runs (the tile is free: the previous layer's copy-out ended at the barrier below)
        global_to_stage(t.ysave[l], H);
        named_bar_sync(2, FN_EPI);
      }
      mbar_wait(bar_d + 8 * l, 0);
      tc_fence_after();
      if (warp == 0) fn_stamp(2 + 2 * l);
#pragma unroll 1
      for (int c0 = q * cw; c0 < (q + 1) * cw; c0 += 32) {
        float v[32];
        tmem_ld32(dcol + (uint32_t)c0, v);
        uint32_t kw = 0xFFFFFFFFu;
        if (drop) kw = drop_word32(dc, stream, (uint64_t)grow, (uint32_t)(c0 >> 5));
        if (!BWD) {
          const float* bias = t.bias[l];
#pragma unroll
          for (int j = 0; j < 32; ++j) {   // bias: warp-uniform address (parameters need not be 16-byte aligned)
            float x = lrelu(v[j] + __ldg(bias + c0 + j), t.alpha);
            if (drop) x = ((kw >> j) & 1u) ? x * dc.scale : 0.f;
            v[j] = x;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float y4[4];
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(y4[0]), "=f"(y4[1]), "=f"(y4[2]), "=f"(y4[3])
                         : "r"(srow + ((((uint32_t)(c0 + j) >> 2) ^ ((uint32_t)row & 7u)) << 4)));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float gfac = lrelu_grad_from_out(y4[e], t.alpha);
              if (drop) gfac = ((kw >> (j + e)) & 1u) ? gfac * dc.scale : 0.f;
              v[j + e] *= gfac;
            }
          }
        }
        // rounded to TF32 once: the saved copy only ever feeds TF32 GEMMs (weight gradients) and the sign test
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = tf32_rna(v[j]);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(srow + ((((uint32_t)(c0 + j) >> 2) ^ ((uint32_t)row & 7u)) << 4)),
                       "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
        tmem_st32(dcol + (uint32_t)c0, v);
        if (BWD && t.dbias[1 - l] != nullptr) {
          // db = column sums of dz: butterfly transpose-reduce over the warp's 32 rows (31 shuffles for 32
          // columns; rows past M hold zeros), lane j ends with column c0 + j, one coalesced reduction per warp
#pragma unroll
          for (int ofs = 16; ofs >= 1; ofs >>= 1) {
            const bool up = (lane & ofs) != 0;
#pragma unroll
            for (int j = 0; j < ofs; ++j) {
              const float send = up ? v[j] : v[j + ofs];
              const float keep = up ? v[j + ofs] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
            }
          }
          atomicAdd(t.dbias[1 - l] + c0 + lane, v[0]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_e + 8 * l);                  // the next GEMM starts; the copy-out below runs under it
      if (warp == 0) fn_stamp(3 + 2 * l);
      named_bar_sync(1, FN_EPI);                   // staging tile complete
      if (dst != nullptr) stage_to_global(dst, H);
      named_bar_sync(2, FN_EPI);                   // staging tile free again
    }

    // ---- last layer: 8-column chunks q, q+4, ... of the accumulator at column 0 -----------------------
    mbar_wait(bar_d + 16, 0);
    tc_fence_after();
    if (warp == 0) fn_stamp(6);
    const int NC = t.Na + t.Nb;
    for (int c0 = q * 8; c0 < t.n[2]; c0 += 32) {
      float v[8];
      tmem_ld8(tl + (uint32_t)c0, v);
      if (!valid || c0 >= NC) continue;
      if (!BWD) {
        uint32_t kw = 0xFFFFFFFFu;
        if (drop) kw = drop_word32(dc, t.stream[2], (uint64_t)grow, (uint32_t)(c0 >> 5));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = c0 + e;
          if (c < NC) {
            float x = v[e] + __ldg(t.bias[2] + c);
            if (drop) x = ((kw >> (c & 31)) & 1u) ? x * dc.scale : 0.f;
            v[e] = x;
          }
        }
      }
      // a thread owns a row: every warp-wide store touches 32 lines, so one 32-byte store where alignment allows
      if (c0 + 7 < t.Na && t.vec_oa) {
        float* o = t.outa + (size_t)grow * t.ldoa + c0;
        if (t.vec_oa == 2) {
          stg256(o, v);
        } else {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else if (c0 >= t.Na && c0 + 7 < NC && t.vec_ob) {
        float* o = t.outb + (size_t)grow * t.ldob + (c0 - t.Na);
        if (t.vec_ob == 2) {
          stg256(o, v);
        } else {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = c0 + e;
          if (c < t.Na) t.outa[(size_t)grow * t.ldoa + c] = v[e];
          else if (c < NC) t.outb[(size_t)grow * t.ldob + (c - t.Na)] = v[e];
        }
      }
    }
  }

  if (warp == 0) fn_stamp(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) fn_stamp(8);
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
// Weight gradients of the node network: dW_l += dz_l^T in_l  (l = 0, 1, 2; reduction over the M rows).
// dz_l [M, na] and in_l [M, k] are row-major, i.e. MN-major operands of a GEMM whose K dimension is the row
// index: 32-row sub-tiles are staged in shared memory as 32-column blocks of [32 rows x 128 B] (128-byte
// swizzle with 32-byte atomicity, the only MN-major TF32 layout) by cp.async and read by tcgen05.mma through
// MN-major descriptors (32-element MN blocks 4 KB apart, 8 rows of K per MMA).  The [na x k] fp32 accumulator of one layer fills TMEM (two M = 128
// halves x 256 columns); a CTA owns a contiguous range of (layer, 128-row tile) items, layer-major, and
// flushes with 16-byte reductions when the layer changes.
// ---------------------------------------------------------------------------------------------------
constexpr int DW_PROD = 256;                      // producer / flush threads (8 warps)
constexpr int DW_THREADS = DW_PROD + 32;          // + MMA issuer warp
constexpr int DW_STAGES = 3;
constexpr uint32_t DW_BLK = 32 * 128;             // one 32-column block of a 32-row sub-tile
constexpr uint32_t DW_OPER = 8 * DW_BLK;          // one operand: up to 256 columns
constexpr uint32_t DW_STAGE = 2 * DW_OPER;        // 64 KB
constexpr uint32_t DW_OFF_BAR = DW_STAGES * DW_STAGE;
constexpr uint32_t DW_SMEM = DW_OFF_BAR + 256 + 1024;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// MN-major TF32 operands exist in ONE shared-memory layout: 128-byte swizzle with 32-byte atomicity (layout type
// 1): rows of 32 MN elements (128 B), the four 32-byte chunks of a row XOR-ed with (row & 3), 4-row K groups SBO
// apart (512 B: rows are contiguous), 32-element MN blocks LBO apart.
__device__ __forceinline__ uint64_t desc_mn_tf32(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(DW_BLK >> 4) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}

__global__ void __launch_bounds__(DW_THREADS, 1) fn_dw_kernel(FnDwArgs t) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + DW_OFF_BAR;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * DW_STAGES, bar_acc = bar0 + 16 * DW_STAGES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + DW_OFF_BAR + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, DW_PROD);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int items = t.nslots * t.T;
  const int i0 = (int)((long long)items * blockIdx.x / gridDim.x);
  const int i1 = (int)((long long)items * (blockIdx.x + 1) / gridDim.x);
  uint32_t g = 0;       // sub-tiles streamed so far (ring position), same sequence in every role
  int run = 0;
  for (int ib = i0; ib < i1; ++run) {
    const int l = ib / t.T;                               // layer slot of this run
    const int ie = min(i1, (l + 1) * t.T);                // items [ib, ie) share the layer
    const int na = t.na[l], K = t.ka[l] + t.kb[l];
    const int kpad = (K + 15) & ~15;
    const int nblkA = (na + 31) >> 5, nblkB = (kpad + 31) >> 5;
    const int nmh = (na + 127) >> 7;
    const int nsub = (ie - ib) * 4;
    if (warp < 8) {
      // ============================ producers ===========================================================
      const int chA = nblkA * 8, chB = nblkB * 8;         // 16-byte chunks per row
      const int per_sub = 32 * (chA + chB);
      // The chunk -> (source, destination) map of a thread is the same for every sub-tile of the run: computed once
      // (the per-chunk divisions made eight producer warps instruction-latency bound: ~17 us per 128-row item).
      constexpr int MAXC = 16;                            // 32 * (64 + 64) chunks / 256 threads
      const bool fast = t.vec_dz[l] && t.vec_a[l] && (t.kb[l] == 0 || t.vec_b[l]) && (na & 31) == 0 && (K & 31) == 0;
      uint32_t dsto[MAXC], srco[MAXC], sel = 0;           // sel: 2 bits per chunk (0 dz, 1 ina, 2 inb, 3 none)
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int idx = threadIdx.x + i * DW_PROD;
        dsto[i] = srco[i] = 0;
        uint32_t sl = 3;
        if (idx < per_sub) {
          const bool isB = idx >= 32 * chA;
          const int j = isB ? idx - 32 * chA : idx;
          const int ch = isB ? chB : chA;
          const int r = j / ch, cc = j % ch, col = cc * 4;
          const uint32_t c16 = (uint32_t)cc & 7u;
          dsto[i] = (isB ? DW_OPER : 0u) + (uint32_t)(cc >> 3) * DW_BLK + (uint32_t)r * 128u +
                    ((((c16 >> 1) ^ ((uint32_t)r & 3u)) << 5) | ((c16 & 1u) << 4));
          if (!isB) { sl = 0; srco[i] = (uint32_t)(r * na + col); }
          else if (col < t.ka[l]) { sl = 1; srco[i] = (uint32_t)(r * t.lda[l] + col); }
          else { sl = 2; srco[i] = (uint32_t)(r * t.ldb[l] + (col - t.ka[l])); }
        }
        sel |= sl << (2 * i);
      }
      auto issue = [&](int sidx) {
        const uint32_t gi = g + (uint32_t)sidx, st = gi % DW_STAGES;
        if (gi >= DW_STAGES) mbar_wait(bar_empty + 8 * st, (gi / DW_STAGES - 1) & 1);
        const int rbase = ((ib - l * t.T) + (sidx >> 2)) * 128 + (sidx & 3) * 32;
        const uint32_t sbase = base + st * DW_STAGE;
        if (fast && rbase + 32 <= t.M) {
          const float* s0 = t.dz[l] + (size_t)rbase * na;
          const float* s1 = t.ina[l] + (size_t)rbase * t.lda[l];
          const float* s2 = t.kb[l] ? t.inb[l] + (size_t)rbase * t.ldb[l] : s1;
#pragma unroll
          for (int i = 0; i < MAXC; ++i) {
            const uint32_t sl = (sel >> (2 * i)) & 3u;
            if (sl != 3u) cp_async16(sbase + dsto[i], (sl == 0 ? s0 : (sl == 1 ? s1 : s2)) + srco[i]);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          return;
        }
        for (int idx = threadIdx.x; idx < per_sub; idx += DW_PROD) {
          const bool isB = idx >= 32 * chA;
          const int j = isB ? idx - 32 * chA : idx;
          const int ch = isB ? chB : chA;
          const int r = j / ch, cc = j % ch;
          const int col = cc * 4, gr = rbase + r;
          const uint32_t c16 = (uint32_t)cc & 7u;   // 16-byte chunk of the 128-byte row; its 32-byte pair is swizzled
          const uint32_t dst = sbase + (isB ? DW_OPER : 0u) + (uint32_t)(cc >> 3) * DW_BLK + (uint32_t)r * 128u +
                               ((((c16 >> 1) ^ ((uint32_t)r & 3u)) << 5) | ((c16 & 1u) << 4));
          const float* src = nullptr;   // 16-byte source, or element-wise below
          float v[4] = {0.f, 0.f, 0.f, 0.f};
          bool direct = false;
          if (gr < t.M) {
            if (!isB) {
              if (col + 3 < na && t.vec_dz[l]) { src = t.dz[l] + (size_t)gr * na + col; direct = true; }
              else
                for (int e = 0; e < 4; ++e) if (col + e < na) v[e] = t.dz[l][(size_t)gr * na + col + e];
            } else {
              const int ka = t.ka[l];
              if (col + 3 < ka && t.vec_a[l]) { src = t.ina[l] + (size_t)gr * t.lda[l] + col; direct = true; }
              else if (col >= ka && col + 3 < K && t.vec_b[l]) { src = t.inb[l] + (size_t)gr * t.ldb[l] + (col - ka); direct = true; }
              else
                for (int e = 0; e < 4; ++e) {
                  const int k = col + e;
                  if (k < ka) v[e] = t.ina[l][(size_t)gr * t.lda[l] + k];
                  else if (k < K) v[e] = t.inb[l][(size_t)gr * t.ldb[l] + (k - ka)];
                }
            }
          }
          if (direct) cp_async16(dst, src);
          else asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      auto publish = [&](int sidx) {   // the thread's copies of sub-tile sidx have landed
        fence_async_smem();
        mbar_arrive(bar_full + 8 * ((g + (uint32_t)sidx) % DW_STAGES));
      };
      // two sub-tiles in flight behind the one being published
      for (int s = 0; s < nsub + 2; ++s) {
        if (s < nsub) issue(s);
        else asm volatile("cp.async.commit_group;" ::: "memory");   // empty group keeps the wait distance fixed
        if (s >= 2) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          publish(s - 2);
        }
      }
      // ============================ flush ================================================================
      mbar_wait(bar_acc, run & 1);
      tc_fence_after();
      {
        const int lq = warp & 3, half = warp >> 2;
        const int cbeg = half * (kpad >> 1), cend = cbeg + (kpad >> 1);
        float* dw = t.dw[l];
        const int ldw = t.lddw[l];
        for (int mh = 0; mh < nmh; ++mh) {
          const int n = mh * 128 + lq * 32 + lane;
          const uint32_t ta = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)mh * 256u;
          for (int c0 = cbeg; c0 < cend; c0 += 8) {
            float v[8];
            tmem_ld8(ta + (uint32_t)c0, v);
            if (n < na) {
              float* o = dw + (size_t)n * ldw + c0;
              if (c0 + 7 < K && t.vec_dw[l]) {
                red_add_v4(o, v[0], v[1], v[2], v[3]);
                red_add_v4(o + 4, v[4], v[5], v[6], v[7]);
              } else {
                for (int e = 0; e < 8; ++e) if (c0 + e < K) atomicAdd(o + e, v[e]);
              }
            }
          }
        }
      }
      tc_fence_before();
    } else {
      // ============================ MMA issuer ===========================================================
      const uint32_t idesc = idesc_tf32(kpad) | (1u << 15) | (1u << 16);   // A and B MN-major
      for (int s = 0; s < nsub; ++s) {
        const uint32_t gi = g + (uint32_t)s, st = gi % DW_STAGES;
        mbar_wait(bar_full + 8 * st, (gi / DW_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + st * DW_STAGE, sb = sa + DW_OPER;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            for (int mh = 0; mh < nmh; ++mh)
              umma_tf32_ss(tmem + (uint32_t)mh * 256u, desc_mn_tf32(sa + (uint32_t)mh * 4u * DW_BLK + (uint32_t)ks * 1024u),
                           desc_mn_tf32(sb + (uint32_t)ks * 1024u), idesc, (uint32_t)(s | ks));
          umma_commit(bar_empty + 8 * st);
          if (s == nsub - 1) umma_commit(bar_acc);
        }
        __syncwarp();
      }
    }
    g += (uint32_t)nsub;
    ib = ie;
    __syncthreads();   // accumulators flushed before the next layer's first MMA overwrites them
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int mpg_debug_set_fn_trace(void* p) { return (int)cudaMemcpyToSymbol(g_fn_trace, &p, sizeof(p)); }

bool fn_tc_supported(int Ka, int Kb, int H1, int H2, int NO, float p) {
  // NO <= 32: the backward kernel's first A tile (dz2) is one 32-wide K block
  return Ka > 0 && Kb >= 0 && Ka + Kb <= 256 && (H1 == 128 || H1 == 256) && (H2 == 128 || H2 == 256) && NO >= 1 &&
         NO <= 32 && (p == 0.f || p == 0.5f);
}

static int r_up(int v, int m) { return (v + m - 1) / m * m; }

size_t fn_tc_workspace_bytes(int Ka, int Kb, int H1, int H2, int NO) {
  const int K0 = Ka + Kb;
  // forward images: [H1 x K0] [H2 x H1] [NO x H2]; backward: [H2 x NO] [H1 x H2] [K0 x H1] (rows padded to 16,
  // K to 32) -- the workspace holds one direction at a time
  const size_t fwd = (size_t)H1 * r_up(K0, 32) + (size_t)H2 * H1 + (size_t)r_up(NO, 16) * H2;
  const size_t bwd = (size_t)H2 * r_up(NO, 32) + (size_t)H1 * H2 + (size_t)r_up(K0, 16) * H1;
  return (fwd > bwd ? fwd : bwd) * sizeof(float) + 3 * 1024 + 1024;
}

int launch_fn_tc(FnTcArgs t, bool bwd, const float* w0, const float* w1, const float* w2, int H1, int H2, int NO,
                 void* ws, cudaStream_t stream) {
  const int K0 = bwd ? (t.Na + t.Nb) : (t.Ka + t.Kb);   // width of the network input [a | b]
  MPG_CHECK(t.M > 0, "fn_tc: empty batch");
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  FnImageJobs jobs;
  auto set_job = [&](int j, const float* W, int R, int C, bool tr, int rows, int kblocks) {
    FnImageJob& jb = jobs.job[j];
    jb.W = W; jb.ldw = C; jb.R = R; jb.C = C; jb.transposed = tr ? 1 : 0; jb.rows = rows; jb.kblocks = kblocks;
    jb.dst = p;
    t.img[j] = p;
    p += ((size_t)rows * kblocks * 128 + 1023) & ~(size_t)1023;
    return rows * kblocks * 8;
  };
  int chunks = 0;
  if (!bwd) {
    t.n[0] = H1; t.kmma[0] = r_up(K0, 8) / 8;
    t.n[1] = H2; t.kmma[1] = H1 / 8;
    t.n[2] = r_up(NO, 16); t.kmma[2] = H2 / 8;
    chunks += set_job(0, w0, H1, K0, false, H1, (t.kmma[0] + 3) / 4);
    chunks += set_job(1, w1, H2, H1, false, H2, H1 / 32);
    chunks += set_job(2, w2, NO, H2, false, t.n[2], H2 / 32);
  } else {
    t.n[0] = H2; t.kmma[0] = r_up(NO, 8) / 8;
    t.n[1] = H1; t.kmma[1] = H2 / 8;
    t.n[2] = r_up(K0, 16); t.kmma[2] = H1 / 8;
    chunks += set_job(0, w2, NO, H2, true, H2, (t.kmma[0] + 3) / 4);    // B[n = h2][k = o] = W2[o][h2]
    chunks += set_job(1, w1, H2, H1, true, H1, H2 / 32);                // B[n = h1][k = h2] = W1[h2][h1]
    chunks += set_job(2, w0, H1, K0, true, t.n[2], H1 / 32);            // B[n = k0][k = h1] = W0[h1][k0]
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  t.vec_a = (t.a != nullptr) && (t.lda % 4 == 0) && (t.Ka % 4 == 0) && al16(t.a);
  t.vec_b = (t.b != nullptr) && (t.ldb % 4 == 0) && (t.Ka % 4 == 0) && (t.Kb % 4 == 0) && al16(t.b);
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  t.vec_oa = (t.outa != nullptr) && (t.ldoa % 4 == 0) && al16(t.outa);               // 2: 32-byte stores allowed
  t.vec_ob = (t.outb != nullptr) && (t.ldob % 4 == 0) && (t.Na % 8 == 0) && al16(t.outb);
  if (t.vec_oa && t.ldoa % 8 == 0 && al32(t.outa)) t.vec_oa = 2;
  if (t.vec_ob && t.ldob % 8 == 0 && al32(t.outb)) t.vec_ob = 2;
  fn_image_kernel<<<cdiv(chunks, 256), 256, 0, stream>>>(jobs);
  MPG_LAUNCH_CHECK();
  const int grid = cdiv(t.M, 128);
  if (bwd) {
    MPG_CUDA(cudaFuncSetAttribute(fn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FN_SMEM));
    fn_tc_kernel<true><<<grid, FN_THREADS, FN_SMEM, stream>>>(t);
  } else {
    MPG_CUDA(cudaFuncSetAttribute(fn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FN_SMEM));
    fn_tc_kernel<false><<<grid, FN_THREADS, FN_SMEM, stream>>>(t);
  }
  MPG_LAUNCH_CHECK();
  return 0;
}

int launch_fn_dw(FnDwArgs t, cudaStream_t stream) {
  if (t.M <= 0) return 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  t.T = cdiv(t.M, 128);
  MPG_CHECK(t.nslots >= 1 && t.nslots <= 3, "fn_dw: 1..3 products per launch");
  for (int l = 0; l < t.nslots; ++l) {
    const int K = t.ka[l] + t.kb[l];
    MPG_CHECK(t.na[l] >= 1 && t.na[l] <= 256 && K >= 1 && K <= 256, "fn_dw: product %d shape [%d x %d] unsupported", l,
              t.na[l], K);
    t.vec_dz[l] = (t.na[l] % 4 == 0) && al16(t.dz[l]);
    t.vec_a[l] = (t.lda[l] % 4 == 0) && (t.ka[l] % 4 == 0) && al16(t.ina[l]);
    t.vec_b[l] = t.inb[l] != nullptr && (t.ldb[l] % 4 == 0) && (t.ka[l] % 4 == 0) && (t.kb[l] % 4 == 0) && al16(t.inb[l]);
    t.vec_dw[l] = (t.lddw[l] % 4 == 0) && al16(t.dw[l]);
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // The kernel is bound by the per-SM rate at which rows stream in (measured: time ~ 1 / CTAs), so the wide node
  // network products take every SM; items_per_cta > 1 trades that against the per-CTA atomic flush for narrow
  // products.
  const int items = t.nslots * t.T;
  int grid = items / (t.items_per_cta > 0 ? t.items_per_cta : 1);
  if (grid < 1) grid = 1;
  if (grid > sms) grid = sms;
  if (grid > items) grid = items;
  MPG_CUDA(cudaFuncSetAttribute(fn_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SMEM));
  fn_dw_kernel<<<grid, DW_THREADS, DW_SMEM, stream>>>(t);
  MPG_LAUNCH_CHECK();
  return 0;
}

}  // namespace mpg
