// tcgen05 forward of the edge network (included by edge_tc.cu inside its anonymous namespace).
//
// Work unit ("step"): 128 receivers (rows r = b*N + i of the flattened node list) x one sender
// index s of each row's own jet.  A CTA owns a contiguous range of (tile, s) steps and keeps the
// running neighbour sum of its 128 receivers in registers, so the reduction over senders is a
// plain in-thread add and the [B*N*N, hidden] tensors exist only as bf16 tiles in shared memory:
//
//   H0'[r,:] = g(P[r,:] + Q[jet(r)*N + s,:])                    g(v) = v + c|v| = lrelu(v) / sl
//   D1       = H0' * (sl W1)^T  (tcgen05.mma, M=128 N=160 K=96+16)   fp32, TMEM, double buffered
//   H1'      = g(D1)                                             bf16, written back IN PLACE into TMEM
//   D2       = H1' * (sl W2)^T  (A operand from TMEM; two N=96 halves, K=160+16)   fp32, TMEM
//   acc[r,:] += mask[jet(r), s] * g(D2)                          fp32 registers; x sl at the flush
//
// c = (1-alpha)/(1+alpha), sl = (1+alpha)/2: leaky-relu costs ONE fma per element (|v| is a free
// operand modifier) and its scale rides in the next layer's weight image.  The biases ride in an
// extra K=16 step (constant-1 columns of the A tiles x bias_hi/bias_lo rows of the weight images).
//
// Pipeline.  16 epilogue warps (thread <-> TMEM lane = tile row, column quarter) plus one control
// warpgroup: one lane issues every tcgen05.mma, another the bulk-async (TMA) copies of the weight
// images and of the Q rows (a 4-stage shared-memory ring, so no thread waits on global memory per
// step).  Registers are re-split with setmaxnreg (epilogue 120, control 32 per thread).  The tensor pipe executes   ... M2hi(s-1) | M1(s+1) | M2lo(s) | M2hi(s) | M1(s+2) ...  :
// layer 1 runs one step ahead into a double-buffered D1.  H1' never touches shared memory: the
// epilogue warps pack it to bf16 and store it over the first 88 columns of the D1 buffer they read
// (the four warps sharing a TMEM lane quarter meet at a named barrier between the loads and the
// stores), and M2 takes its A operand from there.  Shared-memory bandwidth (128 B/clk) is what bounds
// an SS-mode MMA of this shape: with H1' in TMEM a step moves ~150 KB instead of ~320 KB through it.
// Every epilogue phase then has ~1000 cycles of slack: E1(s) needs only D1(s) (ready one step early),
// E2lo/E2hi(s-1) and H0'(s+2) are produced under M2(s) / M1(s+1).
// TMEM: D1a [0,160) D1b [160,320) D2lo [320,416) D2hi [416,512).
//
// Dropout (p = 0.5): Philox bits as laid out in common.cuh (edge_drop_*: one draw per pair and column quarter,
// made when H0' is built and carried with the step record), applied to the packed bf16 words (layers 0/1)
// through PRMT sign-replication masks or as the predicate of the accumulating FFMA (layer 2); the 2x
// scale of layers 0/1 is folded into the next weight image, the last one into the flush.
// Tiles: 128 consecutive rows of [B*N], or the rows the receiver-compaction map lists (tile_row / tile_jets,
// edge_tc_common.cuh).  Per-row global traffic (P on entering a tile, the aggregate flush) is organised so that
// a warp -- one row per lane -- touches contiguous memory: tile-major P, flush through shared memory + bulk
// reduce-add.

constexpr int F_NEPI = 512;                  // epilogue threads
constexpr int F_NTHR = F_NEPI + 128;         // + control warpgroup (warp 16: MMA issuer, warp 17: TMA loader)
constexpr int F_REGS_EPI = 112, F_REGS_CTL = 32;   // setmaxnreg split of the 64K register file
constexpr int F_QS = 4;                      // Q ring stages
constexpr int F_QJ = 10;                     // jets a 128-row tile can span (N >= 15)
constexpr uint32_t F_QROW = K0 * 4;          // bytes of one Q row
constexpr uint32_t F_QHDR = F_QJ * F_QROW;   // stage header: int tile, int sender, float mask[F_QJ] (written by the loader warp)
constexpr uint32_t F_QSTAGE = F_QHDR + 64;
constexpr int F_NEPIW = 16;                  // epilogue warps = arrivals per barrier phase
constexpr uint32_t F_OFF_Q = OFF_H1;                       // 147456 (no H1 tile in shared memory)
constexpr uint32_t F_OFF_BAR = F_OFF_Q + F_QS * F_QSTAGE;  // 163072
// staging for the accumulator flush: per TMEM lane quarter (4 warps, 32 rows) one D2 half of every row, padded rows
// (conflict-free 16-byte accesses both ways).  Threads own rows, global memory wants lines: see flush()
constexpr uint32_t F_STG_ROW = N2 * 4 + 16;                // 784 bytes: one agg row + pad
constexpr uint32_t F_OFF_STG = F_OFF_BAR + 256;
constexpr uint32_t F_SMEM = F_OFF_STG + 4 * 16 * F_STG_ROW + 1024;   // 214528: 16 rows per lane quarter and round
constexpr uint32_t F_D1_COL = 0, F_D2LO_COL = 320, F_D2HI_COL = 416;
constexpr int NH2 = N2 / 2;                  // 96: columns of one D2 half
constexpr int QH = NH2 / NQ;                 // 24: columns of a D2 half owned by one thread

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// keep mask (0xFFFF per kept bf16 lane) of the element pair (e, e+1), e even, of a 32-element slice
__device__ __forceinline__ uint32_t keep_pair(uint32_t word, int e) {
  return prmt(word << (e >> 2), 0u, ((e >> 1) & 1) ? 0xbbaau : 0x9988u);
}
// keep mask (all ones / zero) of element e of a 32-element slice
__device__ __forceinline__ uint32_t keep_one(uint32_t word, int e) {
  const int byte = 2 * ((e >> 1) & 1) + (e & 1);
  return prmt(word << (e >> 2), 0u, 0x8888u + 0x1111u * (uint32_t)byte);
}
// g(v) = v + c|v|
__device__ __forceinline__ float lrelu_g(float v, float c) { return fmaf(fabsf(v), c, v); }

// optional event trace (build with -DMPG_TRACE): clock64 stamps of block 0, lane 0 of warps 0 and 16
#ifdef MPG_TRACE
__device__ long long* g_trace = nullptr;
#define MPG_TR(it, slot)                                                                        \
  do {                                                                                           \
    if (g_trace && blockIdx.x == 0 && lane == 0 && (it) < 256) g_trace[(it) * 16 + (slot)] = clock64(); \
  } while (0)
#define MPG_TRW(it, slot)                                                                       \
  do {                                                                                           \
    if (g_trace && blockIdx.x == 0 && threadIdx.x == 0 && (it) < 256) g_trace[(it) * 16 + (slot)] = clock64(); \
  } while (0)
#define MPG_TP(slot)                                                                              \
  do {                                                                                            \
    if (g_trace && blockIdx.x == 0 && threadIdx.x == 0) g_trace[4080 + (slot)] = clock64();       \
  } while (0)
#else
#define MPG_TR(it, slot) do { } while (0)
#define MPG_TRW(it, slot) do { } while (0)
#define MPG_TP(slot) do { } while (0)
#endif

template <bool DROP>
__global__ void __launch_bounds__(F_NTHR, 1) edge_tc_fwd_kernel(TcArgs t) {
  extern __shared__ uint8_t smem_raw[];
  const EdgeArgs& a = t.a;
  uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SW128 tiles need 1024-byte alignment
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  opaque(base);   // keep it in a register: every barrier / tile address below is base + constant
  const uint32_t sW1 = base + OFF_W1, sW2 = base + OFF_W2, sH0 = base + OFF_H0, sQ = base + F_OFF_Q;
  const uint32_t bar0 = base + F_OFF_BAR;
  const uint32_t bar_w = bar0, bar_h0 = bar0 + 8, bar_d1 = bar0 + 16 /* [2] */, bar_h1 = bar0 + 32,
                 bar_d2lo = bar0 + 40, bar_d2hi = bar0 + 48, bar_f2lo = bar0 + 56, bar_f2hi = bar0 + 64,
                 bar_q = bar0 + 72 /* [F_QS] */, bar_qe = bar0 + 72 + 8 * F_QS /* [F_QS] */;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + F_OFF_BAR + 192);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_steps = *t.total_steps;
  const long long g0 = total_steps * blockIdx.x / gridDim.x;
  const long long g1 = total_steps * (blockIdx.x + 1) / gridDim.x;
  const int nsteps = (int)(g1 - g0);
  const int2* steps = t.steps + g0;   // this CTA's (tile, sender) list
  const int N = a.N, BN = a.B * a.N;
  // every thread fetches the first record itself: the load is in flight under the barrier / TMEM set-up, and the
  // epilogue threads can start their first tile's row loads without waiting for the loader's first stage
  int2 first = make_int2(0, 0);
  if (nsteps > 0) first = __ldg(steps);
  MPG_TP(0);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_h0, F_NEPIW);
    mbar_init(bar_d1, 1);
    mbar_init(bar_d1 + 8, 1);
    mbar_init(bar_h1, F_NEPIW);
    mbar_init(bar_d2lo, 1);
    mbar_init(bar_d2hi, 1);
    mbar_init(bar_f2lo, F_NEPIW);
    mbar_init(bar_f2hi, F_NEPIW);
    for (int i = 0; i < F_QS; ++i) {
      mbar_init(bar_q + 8 * i, 1);
      mbar_init(bar_qe + 8 * i, F_NEPIW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F_REGS_CTL));
    // both roles run with all 32 lanes converged; single-thread instructions are elect-predicated
    if (warp == 17 && nsteps > 0) {
      // =============================== TMA loader ================================================
      mbar_expect_tx_elect(bar_w, W1_BYTES + W2_BYTES);
      bulk_g2s_elect(sW1, t.w1img, W1_BYTES, bar_w);
      bulk_g2s_elect(sW2, t.w2img, W2_BYTES, bar_w);
      int2 ts = first;
      int l_tile = -1, j0 = 0, nj = 1;   // jets of the tile whose steps are being loaded
      for (int it = 0; it < nsteps; ++it) {
        const int2 ts_next = steps[it + 1 < nsteps ? it + 1 : it];   // in flight while this step's copies are issued
        // stage it % F_QS is free once every builder thread has read the rows of step it - F_QS
        if (it >= F_QS) mbar_wait(bar_qe + 8 * (it & (F_QS - 1)), (it / F_QS - 1) & 1);
        const int q_tile = ts.x, q_s = ts.y;
        if (q_tile != l_tile) { tile_jets(a, q_tile, j0, nj); l_tile = q_tile; }
        const uint32_t bar = bar_q + 8 * (it & (F_QS - 1));
        const uint32_t dst = sQ + (uint32_t)(it & (F_QS - 1)) * F_QSTAGE;
        // header: the step and the sender's mask in each jet of the tile, so that no epilogue thread touches global
        // memory per step (the loader runs F_QS steps ahead: its own loads are off everybody's critical path)
        // the mask load and the row copies are in flight together; the header is stored and the barrier's own
        // arrival (with the byte count) comes last, so the phase cannot complete before the header is visible
        float mv = 1.f;
        if (lane < nj && a.mask) mv = __ldg(a.mask + (size_t)(j0 + lane) * N + q_s);
        for (int j = 0; j < nj; ++j)
          bulk_g2s_elect(dst + (uint32_t)j * F_QROW, a.Q + ((size_t)(j0 + j) * N + q_s) * K0, F_QROW, bar);
        if (lane < nj) asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + F_QHDR + 8 + 4 * (uint32_t)lane), "f"(mv) : "memory");
        if (lane == 0) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst + F_QHDR), "r"(q_tile), "r"(q_s) : "memory");
        __syncwarp();
        mbar_expect_tx_elect(bar, (uint32_t)nj * F_QROW);
        ts = ts_next;
      }
    } else if (warp == 16 && nsteps > 0) {
      // =============================== MMA issuer ================================================
      // operand descriptors = tile base (low word) + compile-time offset.  Everything here derives from
      // warp-uniform values (no opaque()/per-thread registers), so ptxas keeps it in uniform registers
      constexpr uint32_t idesc1 = umma_idesc(N1), idesc2h = umma_idesc(NH2);
      const uint32_t ub = (smem_u32(smem_raw) + 1023u) & ~1023u;
      uint64_t dH0 = umma_desc(ub + OFF_H0), dW1 = umma_desc(ub + OFF_W1), dW2 = umma_desc(ub + OFF_W2);
      auto issue_m1 = [&](int it) {   // D1[it&1] = H0' * W1'^T
        const uint32_t d = tmem + F_D1_COL + (uint32_t)(it & 1) * N1;
        if (elect_one()) {
          opaque(dH0); opaque(dW1);   // no hoisting of the 14 descriptors out of the step loop (32 registers)
#pragma unroll
          for (uint32_t ks = 0; ks < KSTEPS1; ++ks) {
            const uint32_t blk = ks >> 2, j = ks & 3;
            umma_bf16(d, dH0 + ((blk * A_BLK + j * 32) >> 4), dW1 + ((blk * W1_BLK + j * 32) >> 4), idesc1, ks);
          }
          umma_commit(bar_d1 + 8 * (it & 1));
        }
        __syncwarp();
      };
      // one N=96 half of D2; A = H1' (bf16) in the first 88 columns of D1[it&1]
      auto issue_m2 = [&](int it, uint32_t dcol, uint32_t wrow_off, uint32_t bar) {
        const uint32_t at = tmem + F_D1_COL + (uint32_t)(it & 1) * N1;
        if (elect_one()) {
          opaque(dW2);
#pragma unroll
          for (uint32_t ks = 0; ks < KSTEPS2; ++ks) {
            const uint32_t blk = ks >> 2, j = ks & 3;
            umma_bf16_ts(tmem + dcol, at + ks * 8, dW2 + ((blk * W2_BLK + wrow_off + j * 32) >> 4), idesc2h, ks);
          }
          umma_commit(bar);
        }
        __syncwarp();
      };

      mbar_wait(bar_w, 0);
      mbar_wait(bar_h0, 0);
      tc_fence_after();
      issue_m1(0);
      for (int it = 0; it < nsteps; ++it) {
        if (it + 1 < nsteps) {
          mbar_wait(bar_h0, (it + 1) & 1);
          tc_fence_after();
          MPG_TR(it, 8);
          issue_m1(it + 1);
          MPG_TR(it, 9);
        }
        mbar_wait(bar_h1, it & 1);
        MPG_TR(it, 10);
        if (it >= 1) mbar_wait(bar_f2lo, (it - 1) & 1);
#ifdef MPG_M2_UNSPLIT   // experiment: one N=192 MMA per K step
        if (it >= 1) mbar_wait(bar_f2hi, (it - 1) & 1);
        tc_fence_after();
        {
          const uint32_t at = tmem + F_D1_COL + (uint32_t)(it & 1) * N1;
          if (elect_one()) {
            opaque(dW2);
#pragma unroll
            for (uint32_t ks = 0; ks < KSTEPS2; ++ks) {
              const uint32_t blk = ks >> 2, j = ks & 3;
              umma_bf16_ts(tmem + F_D2LO_COL, at + ks * 8, dW2 + ((blk * W2_BLK + j * 32) >> 4), umma_idesc(N2), ks);
            }
            umma_commit(bar_d2lo);
            umma_commit(bar_d2hi);
          }
          __syncwarp();
        }
#else
        tc_fence_after();
        MPG_TR(it, 11);
        issue_m2(it, F_D2LO_COL, 0, bar_d2lo);
        MPG_TR(it, 12);
        if (it >= 1) {
          mbar_wait(bar_f2hi, (it - 1) & 1);
          tc_fence_after();
        }
        MPG_TR(it, 13);
        issue_m2(it, F_D2HI_COL, NH2 * 128, bar_d2hi);
        MPG_TR(it, 14);
#endif
      }
    }
  } else {
    // =============================== epilogue warps ==============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F_REGS_EPI));
    MPG_TP(1);
    if (nsteps > 0) {
    const int q = warp >> 2;                        // column quarter: chunks 4c + q (edge_tc_common.cuh)
    const int row = (warp & 3) * 32 + lane;         // tile row == TMEM lane
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)q * 8;
    const float cg = (1.f - a.alpha) / (1.f + a.alpha);
    DropCfg drop = a.drop;
    if (DROP) resolve_seed(drop);
    // this thread's two 16-byte positions inside a swizzled 128-byte tile row: chunk 4c + q of a tile
    // lives at xs[c & 1] + (c >> 1) * A_BLK + <tile offset>
    uint32_t xs[2];
    xs[0] = base + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4);
    xs[1] = base + (uint32_t)row * 128u + ((uint32_t)((4 + q) ^ (row & 7)) << 4);
    opaque(xs[0]);
    opaque(xs[1]);

    // constant 1.0 columns of the bias K-step (cols 96,97 of H0'; 160,161 of H1'), zeros after them
    if (q == 0) {
      st_ones_chunk(sH0 + swz_chunk(row, 96, A_BLK));
      st_zero_chunk(sH0 + swz_chunk(row, 104, A_BLK));
    }

    float acc[2 * QH];
#pragma unroll
    for (int c = 0; c < 2 * QH; ++c) acc[c] = 0.f;
    float Preg[Q0];

    // ---- state of the H0' builder (step it+2) -----------------------------------------------------
    int h_loaded = -1, h_r = 0;
    uint32_t h_qoff = 0, h_moff = 0;
    bool h_valid = false;
    // step records (tile, sender, this row's mask multiplier) of the steps whose H0' has been built: step it+2 is
    // built during iteration it, so three are live
    int s_tile[3] = {0, 0, 0}, s_snd[3] = {0, 0, 0}, s_row[3] = {0, 0, 0};
    float s_m[3] = {0.f, 0.f, 0.f};
    u4 s_kb[3] = {};   // DROP: the step's Philox draw (layer 0 = y << 2, consumed at once; the rest two steps later)
    auto tile_rows = [&](int tile) {   // this thread's row of `tile`: index, stage offsets, its P values
      h_loaded = tile;
      int tj0, tnj;
      tile_jets(a, tile, tj0, tnj);
      const int r = tile_row(a, tile, row);                 // padded row of this lane, -1: none
      h_valid = r >= 0;
      // lanes without a row compute on a row that exists (their mask multiplier is 0): the last one / the tile's first jet
      const int rc = h_valid ? r : (a.cmap ? tj0 * N : BN - 1);
      h_r = rc;
      const uint32_t jl = (uint32_t)(rc / N - tj0);
      h_moff = F_QHDR + 8 + 4 * jl;
      h_qoff = jl * F_QROW + (uint32_t)q * 32u;
      // row-major: 8 floats at column 32c + 8q; tiled (EdgeArgs::p_tiled): column groups 8c + 2q (+1), the warp's 32
      // rows of one group contiguous, indexed by the lane's position in the tile (lanes without a row: any lane's)
      const int prow = (a.cmap || h_valid) ? row : rc - tile * TILE;
      const float* p = a.p_tiled ? a.P + (size_t)tile * TILE * K0 + ((size_t)(2 * q) * TILE + prow) * 4
                                 : a.P + (size_t)rc * K0 + q * 8;
      const int cs = a.p_tiled ? 8 * TILE * 4 : 32, hs = a.p_tiled ? TILE * 4 : 4;
#pragma unroll
      for (int c = 0; c < Q0 / 8; ++c) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + cs * c));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + cs * c + hs));
        Preg[8 * c] = v0.x; Preg[8 * c + 1] = v0.y; Preg[8 * c + 2] = v0.z; Preg[8 * c + 3] = v0.w;
        Preg[8 * c + 4] = v1.x; Preg[8 * c + 5] = v1.y; Preg[8 * c + 6] = v1.z; Preg[8 * c + 7] = v1.w;
      }
    };
    auto build_h0 = [&](int it) {
      mbar_wait(bar_q + 8 * (it & (F_QS - 1)), (it / F_QS) & 1);
      const uint32_t stage = sQ + (uint32_t)(it & (F_QS - 1)) * F_QSTAGE;
      int h_tile, h_s;
      asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(h_tile), "=r"(h_s) : "r"(stage + F_QHDR));
      if (h_tile != h_loaded) tile_rows(h_tile);
      {
        float mv;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mv) : "r"(stage + h_moff));
        s_tile[2] = h_tile; s_snd[2] = h_s; s_row[2] = h_valid ? h_r : -1; s_m[2] = h_valid ? mv : 0.f;
      }
      uint32_t kw = 0;
      if (DROP) {
        s_kb[2] = edge_drop_bits(drop.seed, (uint64_t)h_r * N + h_s, q, 1);
        kw = s_kb[2].y << 2;
      }
      const uint32_t qa = stage + h_qoff;
#pragma unroll
      for (int c = 0; c < Q0 / 8; ++c) {
        float v[8];
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(qa + c * 128));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(qa + c * 128 + 16));
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float x0 = lrelu_g(v[e] + Preg[8 * c + e], cg), x1 = lrelu_g(v[e + 1] + Preg[8 * c + e + 1], cg);
          w[e >> 1] = pack_bf16(x0, x1);
          if (DROP) w[e >> 1] &= keep_pair(kw, 8 * c + e);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(xs[c & 1] + OFF_H0 + (c >> 1) * A_BLK),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_qe + 8 * (it & (F_QS - 1)));   // Q rows consumed (generic-proxy reads are complete)
        mbar_arrive(bar_h0);
      }
    };
    auto rotate = [&]() {   // records of steps it+1, it+2 become those of it, it+1
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        s_tile[i] = s_tile[i + 1]; s_snd[i] = s_snd[i + 1]; s_row[i] = s_row[i + 1]; s_m[i] = s_m[i + 1];
        if (DROP) s_kb[i] = s_kb[i + 1];
      }
    };

    // ---- tile state: `cur` = tile of step `it` (mask / dropout rows), `acc` = tile the accumulators belong to
    int acc_tile = 0, acc_R = -1;   // tile the accumulators belong to; this lane's padded row in it (-1: none)
    float e_m = 0.f;      // mask multiplier of the step whose E2 is pending
    const float fl_scale = a.out_scale * (DROP ? 2.f : 1.f) * 0.5f * (1.f + a.alpha);
    // A thread holds one row's 8-column chunks: 12 reductions of 16 bytes per thread, 6144 per CTA and flush, all
    // CTAs at once at the end of the kernel (~6.5 k clk per flush, bound by the L2 reduction rate of 16-byte packets).
    // Instead the rows go to shared memory (padded: conflict-free) and one bulk reduce-add per row (768 contiguous
    // bytes, cp.reduce.async.bulk) adds them to agg asynchronously: 16 rows per lane quarter and round, each of the
    // quarter's four warps issues four of them.
    auto flush = [&]() {
      const int g = warp & 3;
      const uint32_t stg = base + F_OFF_STG + (uint32_t)g * (16u * F_STG_ROW);
#pragma unroll 1
      for (int rd = 0; rd < 2; ++rd) {
        if ((lane >> 4) == rd) {
          const uint32_t dst = stg + (uint32_t)(lane & 15) * F_STG_ROW + (uint32_t)(8 * q) * 4u;
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < QH / 8; ++c)
#pragma unroll
              for (int e = 0; e < 8; e += 4) {
                const float* v = acc + h * QH + 8 * c + e;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst + (uint32_t)(h * NH2 + 32 * c + e) * 4u),
                             "f"(v[0] * fl_scale), "f"(v[1] * fl_scale), "f"(v[2] * fl_scale), "f"(v[3] * fl_scale) : "memory");
              }
        }
        fence_async_smem();                 // generic-proxy writes -> visible to the bulk (async proxy) reads
        named_bar_sync(1 + g, 128);
        int grow[4];   // padded rows of the four staged rows this warp issues (each lane knows its own: acc_R)
#pragma unroll
        for (int i = 0; i < 4; ++i) grow[i] = __shfl_sync(0xffffffffu, acc_R, 16 * rd + 4 * q + i);
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 4 * q + i;
            if (grow[i] >= 0) bulk_reduce_add_f32(a.agg + (size_t)grow[i] * N2, stg + (uint32_t)rr * F_STG_ROW, N2 * 4);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // rows read: the staging may be rewritten
        }
        named_bar_sync(1 + g, 128);
      }
#pragma unroll
      for (int c = 0; c < 2 * QH; ++c) acc[c] = 0.f;
    };
    // one D2 half (this thread's 3 chunks): acc += m * keep * g(D2)
    auto e2_half = [&](uint32_t dcol, float* ac, uint32_t kw) {
      float v[QH];
      tmem_ld8x3(tl + dcol, v);
#pragma unroll
      for (int e = 0; e < QH; ++e) {
        const float g = lrelu_g(v[e], cg);
        if (DROP) {   // a compile-time bit of the keep word: one predicate-setting LOP3 + a predicated FFMA
          if (kw & (1u << edge_drop_bitpos(e))) ac[e] = fmaf(g, e_m, ac[e]);
        } else {
          ac[e] = fmaf(g, e_m, ac[e]);
        }
      }
    };
    uint32_t kz = 0, kwd = 0;   // layer-2 keep words of the step whose E1 ran last (consumed one iteration later)

    MPG_TP(2);
    tile_rows(first.x);   // the first tile's P rows are requested before the loader's first stage is waited for
    build_h0(0);
    rotate();
    rotate();              // record of step 0 -> slot 0
    acc_tile = s_tile[0];
    acc_R = s_row[0];
    MPG_TP(3);
    if (nsteps > 1) {
      mbar_wait(bar_d1, 0);   // M1(0) done: the H0' tile may be overwritten
      build_h0(1);
      s_tile[1] = s_tile[2]; s_snd[1] = s_snd[2]; s_row[1] = s_row[2]; s_m[1] = s_m[2];
      if (DROP) s_kb[1] = s_kb[2];
    }

    for (int it = 0; it < nsteps; ++it) {
      const int cur_tile = s_tile[0], cur_s = s_snd[0], cur_row = s_row[0];
      const float m_cur = s_m[0];   // mask multiplier of step `it`
      // ---- E2lo(it-1) ----------------------------------------------------------------------------
      if (it >= 1) {
        mbar_wait(bar_d2lo, (it - 1) & 1);
        MPG_TR(it, 0);
        tc_fence_after();
        e2_half(F_D2LO_COL, acc, kz);
        tc_fence_before();
        warp_arrive(bar_f2lo);
        MPG_TR(it, 1);
      }
      // ---- E1(it): D1 -> H1' ---------------------------------------------------------------------
      uint32_t kx = 0, ky = 0, kz_n = 0, kw_n = 0;
      if (DROP) { kx = s_kb[0].x; ky = s_kb[0].y; kz_n = s_kb[0].z; kw_n = s_kb[0].w; }   // drawn when H0'(it) was built
      mbar_wait(bar_d1 + 8 * (it & 1), (it >> 1) & 1);
      MPG_TR(it, 3);
      tc_fence_after();
      {
        const uint32_t d1 = tl + F_D1_COL + (uint32_t)(it & 1) * N1;   // column of chunk q of D1[it&1]
        float v[Q1];
        tmem_ld8x5(d1, v);
        uint32_t w[Q1 / 2];
#pragma unroll
        for (int i = 0; i < Q1; i += 2) {
          w[i >> 1] = pack_bf16(lrelu_g(v[i], cg), lrelu_g(v[i + 1], cg));
          if (DROP) w[i >> 1] &= keep_pair(i < 32 ? kx : ky, i & 31);
        }
        // H1' goes back into D1[it&1]: element column 32c + 8q + e -> packed column 16c + 4q + e/2.  Those
        // columns hold fp32 values other warps of this lane quarter are still reading: meet them first
        named_bar_sync(1 + (warp & 3), 128);
        const uint32_t h1 = d1 - (uint32_t)q * 4;                      // packed column 4q
#pragma unroll
        for (int c = 0; c < Q1 / 8; ++c) tmem_st4(h1 + 16 * c, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        if (q == 0) tmem_st8(d1 + N1 / 2, 0x3F803F80u, 0u);            // bias K-step: columns 160,161 = 1.0
        tmem_st_wait();
      }
      tc_fence_before();
      warp_arrive(bar_h1);
      MPG_TR(it, 4);
      // ---- E2hi(it-1) ----------------------------------------------------------------------------
      if (it >= 1) {
        mbar_wait(bar_d2hi, (it - 1) & 1);
        MPG_TR(it, 2);
        tc_fence_after();
        e2_half(F_D2HI_COL, acc + QH, kwd);
        tc_fence_before();
        warp_arrive(bar_f2hi);
        MPG_TR(it, 5);
        if (cur_tile != acc_tile) {   // step it-1 was the last one of its tile
          flush();
          acc_tile = cur_tile;
          acc_R = cur_row;
        }
      }
      e_m = m_cur;          // multiplier of step `it` (its E2 runs in the next iteration)
      kz = kz_n;
      kwd = kw_n;
      // ---- H0'(it+2) -----------------------------------------------------------------------------
      if (it + 2 < nsteps) {
        mbar_wait(bar_d1 + 8 * ((it + 1) & 1), ((it + 1) >> 1) & 1);   // M1(it+1) done: H0' tile free
        MPG_TR(it, 6);
        build_h0(it + 2);
        MPG_TR(it, 7);
      }
      rotate();
    }
    MPG_TP(4);
    // ---- drain: E2 of the last step ------------------------------------------------------------------
    mbar_wait(bar_d2lo, (nsteps - 1) & 1);
    mbar_wait(bar_d2hi, (nsteps - 1) & 1);
    tc_fence_after();
    e2_half(F_D2LO_COL, acc, kz);
    e2_half(F_D2HI_COL, acc + QH, kwd);
    MPG_TP(5);
    flush();
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's reductions have been performed
    MPG_TP(6);
    }
  }

  tc_fence_before();
  __syncthreads();
  MPG_TP(7);
  if (warp == 16) tmem_dealloc(tmem, TMEM_COLS);
}
